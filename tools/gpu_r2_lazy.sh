#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_lazy.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=900 run python -m pytest tests -x -q -m gpu
run python tools/diag_dist.py --tag lazy --reps 3
run python tools/diag_dist.py --tag lazy_k400 --steps 400 --reps 3
RTM_STEP_LAZY=0 run python tools/diag_dist.py --tag eager_k400 --steps 400 --reps 3
RTM_STEP_GRID=96 run python tools/diag_dist.py --tag lazy_k400_g96 --steps 400 --reps 3
RTM_STEP_GRID=128 run python tools/diag_dist.py --tag lazy_k400_g128 --steps 400 --reps 3
run python tools/step_timeline.py async
grep -E "^diag|passed|failed|rc=|Error|error|team 0|step period|scan phase" $L | sed -E 's/rank=0\/1 dist=0 sampler=0 side=0 //; s/maxconn=- //' | cut -c1-330
