#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_lb.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=600 run python -m pytest tests/test_gpu_detect.py -x -q -m gpu
run python tools/probe_letterbox.py
run env RTM_LETTERBOX_IMPL=unstaged python tools/probe_letterbox.py
grep -E "^letterbox|passed|failed|rc=" $L | cut -c1-300
