#!/usr/bin/env python
"""Where the host side of a step goes: StreamBatch.step() split into its parts, 2000 calls each (the GPU is left to
drain between parts so that nothing blocks on a full launch queue)."""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

dev = torch.device("cuda", 0)
S, F, N = 64, 4, 2000
wl = PostBackboneWorkload(S, F, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
for f in range(50):
    sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f, heads_ready=True)
torch.cuda.synchronize()


def clock(fn, n=N):
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    return (time.perf_counter() - t0) / n * 1e6


print("whole step()            %.2f us" % clock(lambda i: sb.step(wl.heads[i % F], now=1.7e9 + i / 30, frame_id=i, heads_ready=True)))
torch.cuda.synchronize()
print("_io()                   %.2f us" % clock(lambda i: sb._io(wl.heads[i % F], 1.7e9 + i / 30, i)))
print("current_stream handle   %.2f us" % clock(lambda i: torch.cuda.current_stream().cuda_stream))
print("current_device          %.2f us" % clock(lambda i: torch.cuda.current_device()))
st = torch.cuda.current_stream().cuda_stream


def call(i):
    io = sb._io(wl.heads[i % F], 1.7e9 + i / 30, i)
    io.scan_async, io.results_alternate, io.heads_ready_event = 1, 1, None
    sb.lib.rtm_post_backbone_step(C.byref(io), C.byref(sb.params), st)
    sb._advance()


print("_io + C call + advance  %.2f us" % clock(call))
torch.cuda.synchronize()
for chunk in (20, 20, 20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(chunk):
        sb.step(wl.heads[i % F], now=1.7e9 + i / 30, frame_id=i, heads_ready=True)
    print("first %d steps after a synchronize: %.2f us per step" % (chunk, (time.perf_counter() - t0) / chunk * 1e6))
torch.cuda.synchronize()
sb.check_status()
