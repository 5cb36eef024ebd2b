#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tracker.py -x -q -m gpu -k "size_limits or lapjv or optimal" > gpurun_out/pytest_lap.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_lap.log | cut -c1-300
