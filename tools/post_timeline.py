#!/usr/bin/env python
"""Where the post kernel's time goes: %globaltimer stamps at its stage boundaries.

    python tools/post_timeline.py --build      # here (no GPU): tools/librtmodt_b200_tl.so, -DRTM_TIMELINE
    python tools/post_timeline.py [steps]      # on the GPU box: bench workload, mean ns per stage

The diagnosis library is the product source compiled with -DRTM_TIMELINE; the product library has
no trace of the instrumentation.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TL_LIB = os.path.join(ROOT, "tools", "librtmodt_b200_tl.so")
MARKS = {0: "start", 1: "nms: count", 2: "nms: expand ranks", 3: "nms: stage + keys", 4: "nms: sort", 5: "nms: sorted boxes",
         7: "nms: scan, warp segments", 8: "nms: scan, long segments", 6: "nms: scan, order survivors", 10: "nms: write detections", 11: "trk: stage dets + split", 12: "trk: stage 1",
         13: "trk: stage 2", 14: "trk: births", 20: "trk: update + compact", 21: "zone: stage tables", 30: "zone: tests + events"}

if "--build" in sys.argv:
    b = importlib.import_module("real-time-multi-object-detection---tracking-system_b200.build")
    print(b.build(force=True, out=TL_LIB, defines=("RTM_TIMELINE",)))
    sys.exit(0)

os.environ["RTM_LIB_PATH"] = TL_LIB
if "--fused" not in sys.argv:
    os.environ["RTM_STEP_FUSED"] = "0"  # the two-launch post kernel (stamps indexed by blockIdx = stream); --fused: the step kernel's workers
sys.argv = [a for a in sys.argv if a != "--fused"]
import ctypes as C
import numpy as np
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 48
S, F = 64, 16
dev = torch.device("cuda", 0)
wl = PostBackboneWorkload(S, F, first_stream=0, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev,
                     assignment=os.environ.get("TL_ASSIGNMENT", "greedy"))   # TL_ASSIGNMENT=lapjv: the optimal-assignment mode
lib = sb.lib
lib.rtm_debug_timeline.restype, lib.rtm_debug_timeline.argtypes = C.c_int, [C.c_void_p]
buf = torch.zeros((16384 + 16 * 512 * 8,), dtype=torch.int64, device=dev)   # stage rows first, the step kernel's CTA rows behind
for f in range(16):
    sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f)
torch.cuda.synchronize()
pkg._lib.check(lib.rtm_debug_timeline(buf.data_ptr()))
acc = []
for f in range(16, 16 + steps):
    buf.zero_()
    sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f)
    torch.cuda.synchronize()
    acc.append(buf[:S * 32].view(S, 32).cpu().numpy().copy())
pkg._lib.check(lib.rtm_debug_timeline(None))
t = np.stack(acc).astype(np.float64)          # (steps, S, 32)
ids = [0, 1, 2, 3, 4, 5, 7, 8, 6, 10, 11, 12, 13, 14, 20, 21, 30]
print(f"post kernel timeline, {steps} steps x {S} streams (mean ns per CTA; max = slowest stream of a step, mean over steps)")
prev = ids[0]
for i in ids[1:]:
    d = t[:, :, i] - t[:, :, prev]
    ok = (t[:, :, i] > 0) & (t[:, :, prev] > 0)
    if ok.any():
        print(f"  {MARKS[i]:28s} mean {d[ok].mean():8.0f}   max {np.where(ok, d, 0).max(axis=1).mean():8.0f}")
        prev = i
tot = t[:, :, 30] - t[:, :, 0]
print(f"  {'whole CTA':28s} mean {tot.mean():8.0f}   max {tot.max(axis=1).mean():8.0f}")
span = t[:, :, 30].max(axis=1) - t[:, :, 0].min(axis=1)
print(f"  first CTA start -> last CTA end: {span.mean():.0f} ns")
# the slowest CTA of each step: where its time went (mean over steps) - that CTA sets the kernel's duration
slow = tot.argmax(axis=1)
print("  slowest CTA of a step, per stage:")
prev = ids[0]
for i in ids[1:]:
    a, b = t[np.arange(len(slow)), slow, i], t[np.arange(len(slow)), slow, prev]
    ok = (a > 0) & (b > 0)
    if ok.any():
        print(f"    {MARKS[i]:28s} {np.mean((a - b)[ok]):8.0f}   (in {int(ok.sum())} of {len(slow)} steps)")
        prev = i
print("  slowest streams:", np.bincount(slow, minlength=S).argsort()[::-1][:6].tolist())
# tail counters (RTM_TIMELINE builds): per CTA and step
names = {22: "tail chunks with work", 23: "alive candidates in them", 24: "tail survivors", 25: "cycles vs earlier tail survivors",
         26: "cycles chunk resolve", 27: "longest class segment"}
for k, nm in names.items():
    v = t[:, :, k]
    print(f"  {nm:34s} mean {v.mean():9.1f}   slowest CTA {v[np.arange(len(slow)), slow].mean():9.1f}   max {v.max():9.0f}")
