#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "variants" > gpurun_out/pytest_var.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_var.log | cut -c1-400
