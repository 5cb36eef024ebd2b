#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_lb.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=600 run python -m pytest tests/test_gpu_detect.py tests/test_gpu_zones.py -x -q -m gpu
run python tools/probe_letterbox.py
RTM_LETTERBOX_IMPL=pixels run python tools/probe_letterbox.py
run python tools/probe_dense.py
RTM_ZONE_THREADS=256 run python tools/probe_dense.py
grep -E "^letterbox|passed|failed|rc=|Error|error|track|zone" $L | cut -c1-300
tail -c 2500 $L
