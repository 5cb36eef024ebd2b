#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_lb.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=600 run python -m pytest tests/test_gpu_detect.py tests/test_gpu_tracker.py -x -q -m gpu
run python tools/probe_letterbox.py
RTM_LETTERBOX_IMPL=narrow run python tools/probe_letterbox.py
run python tools/probe_letterbox.py
grep -E "^letterbox|passed|failed|rc=|Error|error" $L | cut -c1-300
tail -c 3000 $L
