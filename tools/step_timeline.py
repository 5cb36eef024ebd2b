#!/usr/bin/env python
"""When do the CTAs of the one-launch step kernel run?  %globaltimer stamps by thread 0 of every CTA
(library built with -DRTM_TIMELINE: tools/post_timeline.py --build), 16 consecutive launches kept.

    python tools/step_timeline.py [async|sync] [streams]

Prints, per launch (relative to the first CTA start of the first launch kept): when its scan CTAs start and end,
when its post CTAs become resident, get past their waits and end - i.e. how consecutive steps interleave.
"""
import ctypes as C
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTM_LIB_PATH"] = os.path.join(ROOT, "tools", "librtmodt_b200_tl.so")
import numpy as np
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

mode = sys.argv[1] if len(sys.argv) > 1 else "async"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
F = 8
dev = torch.device("cuda", 0)
wl = PostBackboneWorkload(S, F, first_stream=0, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
lib = sb.lib
lib.rtm_debug_timeline.restype, lib.rtm_debug_timeline.argtypes = C.c_int, [C.c_void_p]
WORDS = 16384 + 16 * 512 * 8
buf = torch.zeros(WORDS, dtype=torch.int64, device=dev)
ready = True if mode == "async" else None
f = 0
def run(n):
    global f
    for _ in range(n):
        sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f, heads_ready=ready)
        f += 1
run(32)
torch.cuda.synchronize()
pkg._lib.check(lib.rtm_debug_timeline(buf.data_ptr()))
run(48)                      # the last 16 launches stay in the buffer
torch.cuda.synchronize()
pkg._lib.check(lib.rtm_debug_timeline(None))
t = buf[16384:].cpu().numpy().reshape(16, 512, 8).astype(np.float64)
first_seq = f - 16
rows = []
for k in range(16):
    seq = first_seq + k                      # library sequence numbers run with f here (one workspace, fused steps only)
    r = t[seq & 15]
    live = r[:, 0] > 0
    rows.append((seq, r[live]))
t0 = min(r[:, 0].min() for _, r in rows)
us = lambda a: (a - t0) / 1e3
print(f"step kernel timeline, mode={mode}, {S} streams: us relative to the first CTA start of the window")
print("  seq | CTAs  start(first, last) | scan over(first, median, last) | post workers: n  streams  first post start(first, last)  end(first, last) | SMs")
for seq, r in rows:
    w = r[r[:, 3] > 0]
    print(f"  {seq:4d} | {len(r):3d} {us(r[:,0].min()):8.2f} {us(r[:,0].max()):8.2f} | {us(r[:,1].min()):8.2f} {us(np.median(r[:,1])):8.2f} {us(r[:,1].max()):8.2f} | "
          f"{len(w):3d} {int(w[:,3].sum()):3d} {us(w[:,5].min()):8.2f} {us(w[:,5].max()):8.2f} {us(w[:,2].min()):8.2f} {us(w[:,2].max()):8.2f} | {len(set(r[:,4].astype(int)))}")
raw = buf[16384:].cpu().numpy().reshape(16, 512, 8)
cw, cl, pw, nt = [], [], [], []
for seq, _ in rows:
    r = raw[seq & 15]
    live = r[:, 0] > 0
    cw.append((r[live, 6] & 0xffffffff).astype(np.float64)); cl.append((r[live, 6] >> 32).astype(np.float64))
    pw.append((r[live, 7] & 0xffffffff).astype(np.float64)); nt.append((r[live, 7] >> 32).astype(np.float64))
cw, cl, pw, nt = (np.concatenate(x) for x in (cw, cl, pw, nt))
print("  team 0 of every CTA: tiles %.1f, consumer loop %.2f us of which waiting for a full stage %.2f us (%.0f %%); per tile: %.2f us, "
      "of which work %.2f us; producer waiting for an empty stage %.2f us" %
      (nt.mean(), cl.mean() / 1e3, cw.mean() / 1e3, 100 * cw.sum() / max(cl.sum(), 1), cl.sum() / max(nt.sum(), 1) / 1e3,
       (cl.sum() - cw.sum()) / max(nt.sum(), 1) / 1e3, pw.mean() / 1e3))
ends = [r[:, 2].max() for _, r in rows]
print("  step period (last end to last end): mean %.2f us" % (np.diff(ends).mean() / 1e3))
scan_len = [np.median(r[:, 1]) - r[:, 0].min() for _, r in rows]
post_len = [((r[:, 2] - r[:, 5]) / np.maximum(r[:, 3], 1))[r[:, 3] > 0].mean() for _, r in rows]
print("  scan phase (first CTA start -> median scan over): mean %.2f us;  post stage per stream (first post start -> end): mean %.2f us" %
      (np.mean(scan_len) / 1e3, np.mean(post_len) / 1e3))
