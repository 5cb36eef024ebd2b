#!/usr/bin/env python
"""Where the host-fed step loses its 4 % against the bare copy: K steps through HostFeeder.step_pinned with CUDA events
around the host->device copy of every step (recorded on the slot's stream through copy_wait / copy_done events)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload
dev = torch.device("cuda", 0)
S, F, K = 64, 4, 40
wl = PostBackboneWorkload(S, F, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
feeder = pkg.HostFeeder(sb, wl.dtype)
pinned = []
for f in range(F):
    host = feeder.alloc_pinned_heads()
    for d, s in zip(host, wl.heads[f]):
        d.copy_(s)
    pinned.append(host)
torch.cuda.synchronize()
def go(n, f0=0):
    res, t_done = [], []
    t0 = time.perf_counter()
    for f in range(f0, f0 + n):
        res.append(feeder.step_pinned(pinned[f % F], now=1.7e9 + f / 30, frame_id=f))
        if len(res) > 1:
            res.pop(0).wait(); t_done.append(time.perf_counter() - t0)
    while res:
        res.pop(0).wait(); t_done.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, np.diff(np.asarray(t_done)) * 1e3
go(6)
dt, iv = go(K, 6)
gbs = feeder.h2d_bytes * K / dt / 1e9
print(f"K={K}: {1e3 * dt / K:.3f} ms per step, {gbs:.2f} GB/s; intervals between completed steps: median {np.median(iv):.3f} ms, min {iv.min():.3f}, max {iv.max():.3f}")
print("first intervals", np.round(iv[:6], 3), "last", np.round(iv[-4:], 3))
# the copy alone, same buffers, one stream
st = torch.cuda.Stream(device=dev)
devb = feeder.slots[0]["dev_heads"][0]._base
with torch.cuda.stream(st):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    e[0].record()
    for i in range(K):
        devb.copy_(pinned[i % F][0]._base, non_blocking=True)
        e[i + 1].record()
st.synchronize()
ms = np.asarray([e[i].elapsed_time(e[i + 1]) for i in range(K)])
print(f"bare copies: median {np.median(ms):.3f} ms = {feeder.h2d_bytes / np.median(ms) / 1e6:.2f} GB/s")
