#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r2.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2.log | cut -c1-300
