#!/usr/bin/env python
"""Diagnosis: why does the scan_async step sequence fall back to the single-stream time under
torch.distributed?  Replays bench.py's timed region (K steps between two
CUDA events on the current stream) under switchable conditions:

    python tools/diag_dist.py [--dist] [--sampler] [--side-stream] [--steps 20] [--reps 5]

--dist         init_process_group("nccl") for a world of this one rank + a barrier before every region
--sampler      bench.py's NVML sampler thread (50 Hz) running meanwhile
--side-stream  step on a torch side stream instead of the legacy default stream
Under torchrun (WORLD_SIZE > 1) --dist uses the real world.  Prints one line per mode.
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--dist", action="store_true")
ap.add_argument("--sampler", action="store_true")
ap.add_argument("--side-stream", action="store_true")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--streams", type=int, default=64)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--tag", default="")
a = ap.parse_args()

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if a.dist:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

S, F, K = a.streams, a.frames, a.steps
wl = PostBackboneWorkload(S, F, first_stream=rank * S, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
f = 0


def run(n, ready):
    global f
    for _ in range(n):
        sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f, heads_ready=ready)
        f += 1


def barrier():
    if a.dist:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(ready):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run(K, ready)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) * 1e3 / K


def host_cost(ready, n=300):
    """wall time of the issuing loop alone (the GPU is still busy when it returns unless the host is the slower side)"""
    import time
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    run(n, ready)
    t1 = time.perf_counter()
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


def body():
    out = {}
    for name, ready in (("async", True), ("single", None)):
        run(5, ready)
        out[name] = sorted(timed(ready) for _ in range(a.reps))
    return out


sampler = None
if a.sampler:
    bench = importlib.import_module("bench")
    sampler = bench.ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.__enter__()
if a.side_stream:
    with torch.cuda.stream(torch.cuda.Stream(device=dev)):
        res = body()
else:
    res = body()
if sampler:
    sampler.__exit__()
sb.check_status()
print(f"diag rank={rank}/{world} dist={int(a.dist)} sampler={int(a.sampler)} side={int(a.side_stream)} K={K} "
      f"maxconn={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS', '-')} {a.tag} "
      + " ".join(f"{k}: min {v[0]:.2f} med {v[len(v) // 2]:.2f} max {v[-1]:.2f} us/step" for k, v in res.items())
      + " | issue loop %.1f us/step (loop + drain %.1f)" % host_cost(True), flush=True)
if a.dist:
    dist.barrier()
    dist.destroy_process_group()
