#!/usr/bin/env python
"""Timing probe: what the orderings around the head scan cost in scan_async mode.

    python tools/probe_overlap.py --build     # here (no GPU): tools/librtmodt_b200_probe.so, -DRTM_PROBES
    RTM_PROBE_BITS=<bits> [RTM_SCAN_TRIGGER=1] python tools/probe_overlap.py [steps]

bits: 1 = no 'scanned' event between scan and post kernel (the post kernel races: results invalid),
2 = no 'consumed' wait before a slot is refilled, 4 = no post kernel at all.  The product library
has no trace of these switches.
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PLIB = os.path.join(ROOT, "tools", "librtmodt_b200_probe.so")
if "--build" in sys.argv:
    b = importlib.import_module("real-time-multi-object-detection---tracking-system_b200.build")
    print(b.build(force=True, out=PLIB, defines=("RTM_PROBES",)))
    sys.exit(0)
os.environ["RTM_LIB_PATH"] = PLIB
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
S, F = 64, 16
dev = torch.device("cuda", 0)
wl = PostBackboneWorkload(S, F, first_stream=0, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
f = 0
def run(n):
    global f
    for _ in range(n):
        sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f, heads_ready=True)
        f += 1
run(20)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()  # wall clock around a device-wide synchronise: with bit 1 no stream sees every kernel
    run(steps)
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / steps * 1e6)
print("bits=%s trigger=%s us_per_step=%.2f" % (os.environ.get("RTM_PROBE_BITS", "0"), os.environ.get("RTM_SCAN_TRIGGER", "0"), best))
