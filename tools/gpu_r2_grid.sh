#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_grid2.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for g in 64 80 96 116 132; do
run env RTM_STEP_GRID=$g python tools/diag_dist.py --tag grid$g --steps 400 --reps 3
run env RTM_STEP_GRID=$g python tools/diag_dist.py --tag grid${g}_k20 --reps 5
done
grep -E "^diag|rc=" $L | cut -c1-220
