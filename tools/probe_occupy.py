#!/usr/bin/env python
"""Timing probe for the one-launch step design: how fast do head scans run while another kernel holds
64 of the 148 SMs (the post CTAs of the step before)?

    python tools/probe_overlap.py --build          # builds tools/librtmodt_b200_probe.so (-DRTM_PROBES)
    python tools/probe_occupy.py [held_sms ...]    # on the GPU box

Scans alone (probe bits 4 + 2: no post kernel, no slot waits), back to back, 400 per measurement, while
`rtm_debug_occupy` keeps `held` SMs busy with one 200 KB CTA each.  Results are timing only.
"""
import ctypes as C
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTM_LIB_PATH"] = os.path.join(ROOT, "tools", "librtmodt_b200_probe.so")
os.environ["RTM_PROBE_BITS"] = "6"
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload

held_list = [int(v) for v in sys.argv[1:]] or [0, 32, 64, 84, 100]
S, F, N = 64, 8, 400
dev = torch.device("cuda", 0)
wl = PostBackboneWorkload(S, F, first_stream=0, device=dev, dtype=torch.bfloat16)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=512, device=dev)
lib = sb.lib
lib.rtm_debug_occupy.restype = C.c_int
lib.rtm_debug_occupy.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p]
side = torch.cuda.Stream(device=dev)
f = 0


def run(n):
    global f
    for _ in range(n):
        sb.step(wl.heads[f % F], now=1.7e9 + f / 30, frame_id=f, heads_ready=True)
        f += 1


run(20)
torch.cuda.synchronize()
for held in held_list:
    best = 1e9
    for rep in range(3):
        if held:
            pkg._lib.check(lib.rtm_debug_occupy(held, 200 * 1024, 40_000_000, side.cuda_stream))
            time.sleep(0.002)                      # the holders are resident and spinning by now
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(N)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / N)
        torch.cuda.synchronize()
    print(f"occupy probe: held_sms={held} trigger={os.environ.get('RTM_SCAN_TRIGGER', '0')} "
          f"ctas={os.environ.get('RTM_TMA_CTAS', '-')} stages={os.environ.get('RTM_TMA_STAGES', '-')} scan_us={best:.2f}", flush=True)
