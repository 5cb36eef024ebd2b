#!/bin/bash
# ncu --set full of the stand-alone tracker / zone kernels on the dense-crowd step (configs[4])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/probe_dense.py > gpurun_out/dense_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/dense_plain.log; exit 1; }
tail -4 gpurun_out/dense_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'track_step_kernel|zone_step_kernel|track_assoc' -s 20 -c 4 -f \
  -o gpurun_out/prof_r2_dense python tools/probe_dense.py > gpurun_out/ncu_dense.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dense.log | cut -c1-200
