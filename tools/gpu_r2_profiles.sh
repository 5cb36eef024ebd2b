#!/bin/bash
# round 2: GPU tests, then the ncu evidence for profiles/ (launch list + --set full of the step kernel and of the
# kernels beside the step).  tools/ncu_summary.py <tag> condenses gpurun_out/ into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r2a}
mkdir -p gpurun_out
SHORT="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras --no-parity"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
timeout 300 $SHORT > gpurun_out/plain_${TAG}.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launches_${TAG}.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'step_kernel' -s 10 -c 3 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
fi
timeout 300 python tools/probe_kernels.py > gpurun_out/probe_kernels_${TAG}.log 2>&1; rc=$?; echo "probe_kernels rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'letterbox|track_step_kernel|zone_step_kernel' -s 8 -c 8 -f -o gpurun_out/prof_${TAG}_side python tools/probe_kernels.py > gpurun_out/ncu_side_${TAG}.log 2>&1
  echo "ncu side rc=$?"; tail -2 gpurun_out/ncu_side_${TAG}.log | cut -c1-200
fi
