#!/bin/bash
# round 2, GPU call 1: diagnosis probes (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_diag1.log
: > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/diag_dist.py --tag plain
run python tools/diag_dist.py --dist --tag dist
run env CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/diag_dist.py --dist --tag dist_conn32
run python tools/diag_dist.py --dist --side-stream --tag dist_side
run python tools/diag_dist.py --sampler --tag sampler
run python tools/probe_occupy.py 0 64 84
run env RTM_SCAN_TRIGGER=1 python tools/probe_occupy.py 0 64 84
run python tools/probe_dense.py
grep -E "^diag|^occupy|^dense|per kernel|live tracks|rc=" $L
