#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused3.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tools/tmabench
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
run env RTM_STEP_FUSED=0 python tools/diag_dist.py --tag twolaunch_k400 --steps 400 --reps 3
run python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu
grep -E "^diag|passed|failed|rc=|^w=" $L | cut -c1-300
