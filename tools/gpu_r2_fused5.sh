#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused5.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=120 run python __graft_entry__.py --smoke
if ! grep -q "smoke ok" $L; then tail -20 $L; exit 1; fi
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
run env RTM_STEP_STAGES=2 python tools/diag_dist.py --tag fused_k400_st2 --steps 400 --reps 3
run python tools/step_timeline.py async
run python tools/step_timeline.py sync
run python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu
grep -E "^diag|passed|failed|rc=" $L | cut -c1-300
grep -A21 "mode=async" $L | cut -c1-250
grep -A21 "mode=sync" $L | tail -5 | cut -c1-250
