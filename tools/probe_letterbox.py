#!/usr/bin/env python
"""Letterbox timing: 64 x 1080p -> 640 x 640 bf16 (the 3 : 1 decimation path) and 720p (general path).
RTM_LETTERBOX_IMPL=narrow (16 pixels per thread, table lookup) | direct (general kernel) select the other kernels."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200 import _lib
dev = torch.device("cuda", 0)
lib = _lib.lib()
S = 64
for (h, w) in ((1080, 1920), (720, 1280)):
    frames = [torch.randint(0, 256, (S, h, w, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
    out = torch.empty((S, 3, 640, 640), dtype=torch.bfloat16, device=dev)
    call = lambda i: _lib.check(lib.rtm_letterbox(frames[i % 3].data_ptr(), S, h, w, w * 3, h * w * 3, out.data_ptr(), _lib.RTM_BF16, 640, 640, _lib.cuda_stream()))
    for i in range(5):
        call(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(30):
        call(i)
    b.record(); b.synchronize()
    us = 1e3 * a.elapsed_time(b) / 30
    print(f"letterbox impl={os.environ.get('RTM_LETTERBOX_IMPL', 'default')} {h}x{w}: {us:.1f} us per {S} frames")
