#!/usr/bin/env python
"""H2D throughput of 155 MB pinned buffers: one stream back to back vs two streams whose copies overlap (what the
host-fed step did before the copies were chained)."""
import time
import torch
dev = torch.device("cuda", 0)
n = 154828800
host = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
devb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
def run(mode, K=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev = None
    for i in range(K):
        s = streams[i % 2] if mode != "one" else streams[0]
        with torch.cuda.stream(s):
            if mode == "chained" and ev is not None:
                s.wait_event(ev)
            devb[i % 2].copy_(host[i % 4], non_blocking=True)
            if mode == "chained":
                ev = torch.cuda.Event()
                ev.record(s)
    torch.cuda.synchronize()
    return n * K / (time.perf_counter() - t0) / 1e9
for mode in ("one", "two", "chained", "one", "two", "chained"):
    run(mode, 4)
    print(f"{mode:8s} {run(mode):.2f} GB/s")
