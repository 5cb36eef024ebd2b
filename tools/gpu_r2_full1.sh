#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_full1.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
TMO=900 run python -m pytest tests -x -q -m gpu
TMO=600 run python bench.py --steps 20 --warmup 5
grep -E "^diag|passed|failed|rc=|Error|error" $L | cut -c1-300
tail -c 6000 $L
