#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_lazy2.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-200} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
D="python tools/diag_dist.py --steps 400 --reps 2"
run $D --tag base
RTM_STEP_STATIC=100 run $D --tag static100
RTM_STEP_STATIC=3 run $D --tag static3
RTM_STEP_STAGES=4 run $D --tag stages4
RTM_STEP_STAGES=2 run $D --tag stages2
RTM_STEP_GRID=148 run $D --tag grid148
RTM_STEP_GRID=96 run $D --tag grid96
RTM_STEP_GRID=80 run $D --tag grid80
RTM_TMA_L2PROMO=2 run $D --tag promo128
RTM_TMA_L2PROMO=0 run $D --tag promo0
RTM_TMA_EVICT_FIRST=0 run $D --tag noevict
RTM_STEP_POST_CTAS=32 run $D --tag workers32
grep -E "^diag|rc=[1-9]|Error|error" $L | sed -E 's/rank=0\/1 dist=0 sampler=0 side=0 //; s/maxconn=- //' | cut -c1-200
