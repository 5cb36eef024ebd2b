#!/usr/bin/env python
"""A few launches of the kernels that are not part of the step: letterbox (1080p -> 640, and a general ratio),
the stand-alone tracker and zone kernels on the dense-crowd clip.  For ncu captures (tools/gpu_r2_profiles.sh)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200 import _lib

dev = torch.device("cuda", 0)
lib = _lib.lib()
S = 64
for (h, w) in ((1080, 1920), (720, 1280)):
    frames = torch.randint(0, 256, (S, h, w, 3), dtype=torch.uint8, device=dev)
    out = torch.empty((S, 3, 640, 640), dtype=torch.bfloat16, device=dev)
    for _ in range(6):
        _lib.check(lib.rtm_letterbox(frames.data_ptr(), S, h, w, w * 3, h * w * 3, out.data_ptr(), _lib.RTM_BF16, 640, 640, _lib.cuda_stream()))
torch.cuda.synchronize()
streams, objects, zones, distinct, frames_n, slots = 128, 1000, 16, 4, 8, 1024
xyxy, conf, cls, count = pkg.synth.scripted_batch(distinct, frames_n, slots, seed=900, **pkg.synth.dense_crowd_kwargs(objects))
rep = streams // distinct
tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * rep, axis=1))).to(dev)
d = [tile(a) for a in (xyxy, conf, cls, count)]
zcfg = [pkg.synth.make_zones(seed=b % distinct, num_zones=zones, width=1920, height=1080, kmin=4, kmax=12) for b in range(streams)]
sb = pkg.StreamBatch(streams, zcfg, src_hw=(1080, 1920), max_det=slots, max_tracks=4096, max_events=4096, device=dev)
order = list(range(frames_n)) + list(range(frames_n - 2, 0, -1))
for k in range(24):
    f = order[k % len(order)]
    sb.track_only(d[0][f], d[1][f], d[2][f], d[3][f], now=1.7e9 + k / 30, frame_id=k)
torch.cuda.synchronize()
sb.check_status()
print("probe_kernels done")
