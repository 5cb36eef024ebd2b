#!/bin/bash
# round 2, GPU call 2: first run of the one-launch step kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_fused1.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python __graft_entry__.py --smoke
TMO=600 run python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu
run python tools/diag_dist.py --tag fused --reps 3
run python tools/diag_dist.py --tag fused_k400 --steps 400 --reps 3
run env RTM_STEP_FUSED=0 python tools/diag_dist.py --tag twolaunch --reps 3
TMO=900 run python -m pytest tests -x -q -m gpu
run python bench.py --steps 20 --warmup 5 --no-cpu
tail -c 3000 $L
