#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_dense2.log
: > $L
run() { echo "== $*" >> $L; timeout ${TMO:-300} "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
TMO=600 run python -m pytest tests/test_gpu_tracker.py tests/test_gpu_zones.py tests/test_gpu_pipeline.py -x -q -m gpu
run python tools/probe_dense.py
run python tools/post_timeline.py --fused 48
grep -E "passed|failed|rc=|^dense|per kernel|live tracks" $L | cut -c1-300
grep -A40 "post kernel timeline" $L | cut -c1-200
