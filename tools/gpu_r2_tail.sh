#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_tail.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=600 run python -m pytest tests/test_gpu_detect.py tests/test_gpu_nms_soak.py tests/test_gpu_pipeline.py -x -q -m gpu
run python tools/diag_dist.py --tag f4 --reps 5
run python tools/diag_dist.py --tag f4_k400 --steps 400 --reps 3
run python tools/post_timeline.py --fused 48
grep -E "^diag|passed|failed|rc=" $L | cut -c1-300
grep -A18 "post kernel timeline" $L | cut -c1-200
