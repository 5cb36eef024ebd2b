#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused6.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for g in 232 248 264 280; do
run env RTM_STEP_GRID=$g python tools/diag_dist.py --tag grid$g --steps 400 --reps 3
done
run env RTM_STEP_GRID=232 RTM_STEP_STATIC=1 python tools/diag_dist.py --tag grid232_static --steps 400 --reps 3
run env RTM_STEP_STATIC=1 python tools/diag_dist.py --tag grid296_static --steps 400 --reps 3
run env RTM_STEP_GRID=232 RTM_TMA_EVICT_FIRST=0 python tools/diag_dist.py --tag grid232_noevict --steps 400 --reps 3
grep -E "^diag|rc=" $L | cut -c1-300
