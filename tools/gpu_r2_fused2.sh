#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused2.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
run env RTM_STEP_STAGES=6 python tools/diag_dist.py --tag fused_k400_st6 --steps 400 --reps 3
run env RTM_STEP_POST_CTAS=32 python tools/diag_dist.py --tag fused_k400_p32 --steps 400 --reps 3
run env RTM_TMA_STATIC_ROUNDS=0 python tools/diag_dist.py --tag fused_k400_s0 --steps 400 --reps 3
run env RTM_STEP_FUSED=0 python tools/diag_dist.py --tag twolaunch --reps 3
TMO=900 run python -m pytest tests -x -q -m gpu
run python bench.py --steps 20 --warmup 5 --no-cpu
grep -E "^diag|passed|failed|rc=" $L | cut -c1-300
bash tools/gpu_r2_dense_ncu.sh
