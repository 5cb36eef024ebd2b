#!/usr/bin/env python
"""Dense-crowd step (BASELINE.json configs[4]) split by kernel: tracker vs zones, with table sizes.

    python tools/probe_dense.py [streams] [objects]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200 import _lib

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 128
objects = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
zones, distinct, frames, steps, slots = 16, 4, 8, 48, 1024
dev = torch.device("cuda", 0)
T0, FPS = 1_700_000_000.0, 30.0
xyxy, conf, cls, count = pkg.synth.scripted_batch(distinct, frames, slots, seed=900, **pkg.synth.dense_crowd_kwargs(objects))
rep = streams // distinct
tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * rep, axis=1))).to(dev)
d_xyxy, d_conf, d_cls, d_count = tile(xyxy), tile(conf), tile(cls), tile(count)
zcfg = [pkg.synth.make_zones(seed=b % distinct, num_zones=zones, width=1920, height=1080, kmin=4, kmax=12) for b in range(streams)]
order = list(range(frames)) + list(range(frames - 2, 0, -1))
sb = pkg.StreamBatch(streams, zcfg, src_hw=(1080, 1920), max_det=slots, max_tracks=4096, max_events=4096, device=dev)
lib = sb.lib
k = 0


def go(n):
    global k
    for _ in range(n):
        f = order[k % len(order)]
        sb.track_only(d_xyxy[f], d_conf[f], d_cls[f], d_count[f], now=T0 + k / FPS, frame_id=k)
        k += 1


go(8)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
go(steps)
b.record()
b.synchronize()
print(f"dense: {streams} streams x {objects} objects: {a.elapsed_time(b) / steps * 1e3:.1f} us per step")
lib.rtm_profile_enable(1)
_lib.profile_read()
ev = []
for _ in range(steps):
    go(1)
    ev.append(int(sb.zones.event_count.sum().item()))
prof = _lib.profile_read()
lib.rtm_profile_enable(0)
tracks, _ = sb.read_tracks()
print("per kernel (alone):", {n: round(1e3 * v[0] / v[1], 1) for n, v in prof.items()}, "us")
print("live tracks per stream:", float(np.mean([len(t) for t in tracks])), "events per step (all streams): mean", np.mean(ev), "max", max(ev),
      "high dets:", float((conf[0, 0, :int(count[0, 0])] >= 0.5).mean()), "dets", int(count[0, 0]))
sb.check_status()
