#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_l2.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for a in 0 600 1184 2400 4800; do
run env RTM_STEP_L2_AHEAD=$a python tools/diag_dist.py --tag l2ahead$a --steps 400 --reps 3
done
run python tools/diag_dist.py --tag default_k20 --reps 5
run python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu
grep -E "^diag|passed|failed|rc=" $L | cut -c1-300
