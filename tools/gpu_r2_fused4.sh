#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused4.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=120 run python __graft_entry__.py --smoke
if ! grep -q "smoke ok" $L; then tail -20 $L; exit 1; fi
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
run python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_zones.py -x -q -m gpu
run python tools/probe_dense.py
grep -E "^diag|passed|failed|rc=|^dense|per kernel" $L | cut -c1-300
