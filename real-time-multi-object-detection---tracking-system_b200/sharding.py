"""Multi-GPU plumbing: static stream -> rank sharding and the end-of-run reduction.

Streams are independent (every stream owns its tracker and zone state - tracker.py:55-56,
zone_engine.py:72-75), so the per-frame path has NO collective: each rank (one process per GPU)
steps its own contiguous block of streams.  Only the run summary crosses ranks, once:
``all_reduce(SUM)`` of an int64 counter vector and an ``all_gather`` of per-rank event counts
(SURVEY.md section 8e).  Works with the ``nccl`` backend on GPUs and ``gloo`` on the CPU.
"""

from __future__ import annotations

from typing import Sequence

COUNTERS = ("frames", "detections_last_step", "births", "events_last_step")


def shard_streams(total_streams: int, world_size: int, rank: int) -> range:
    """Contiguous block of stream ids owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(total_streams, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(stream: int, total_streams: int, world_size: int) -> int:
    for r in range(world_size):
        if stream in shard_streams(total_streams, world_size, r):
            return r
    raise ValueError(f"stream {stream} outside 0..{total_streams - 1}")


def reduce_summary(counters: Sequence[int], events, device="cpu", first_stream: int = 0):
    """Sum ``counters`` over ranks and gather every rank's zone events of the closing step.

    ``events``: this rank's 64-byte ``rtm_zone_event`` records (a NumPy structured array of
    ``_lib.EVENT_DTYPE``, e.g. :meth:`StreamBatch.event_records`), streams numbered within the rank;
    ``first_stream`` is the global id of the rank's stream 0.  Returns ``(totals list[int], records)``:
    all ranks' records in global stream order with the ``stream`` field made global.  The exchange is
    the one SURVEY section 8e specifies: ``all_reduce(SUM)`` of the counter vector, an ``all_gather``
    of the record counts, and an ``all_gather`` of the records padded to the largest count - a few KB,
    once per run.  A plain sequence of integers is accepted in place of records and gathered the same
    way (returned as a flat list).  With no process group initialised it is the identity.
    """
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import _lib
    c = torch.tensor(list(counters), dtype=torch.int64, device=device)
    is_records = isinstance(events, np.ndarray) and events.dtype.names is not None
    if is_records:
        rec = np.ascontiguousarray(events).copy()
        rec["stream"] += np.int32(first_stream)
        payload = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(device)
        unit = rec.dtype.itemsize
    else:
        payload = torch.as_tensor(list(events), dtype=torch.int64).reshape(-1).to(device)
        unit = 1
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed:
        world = dist.get_world_size()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([payload.numel()], dtype=torch.int64, device=device))
        width = max(int(max(int(s.item()) for s in sizes)), 1)
        padded = torch.zeros(width, dtype=payload.dtype, device=device)
        padded[: payload.numel()] = payload
        parts = [torch.zeros(width, dtype=payload.dtype, device=device) for _ in range(world)]
        dist.all_gather(parts, padded)
        payload = torch.cat([g[: int(s.item())] for g, s in zip(parts, sizes)])
    if is_records:
        flat = payload.cpu().numpy().reshape(-1, unit).view(np.dtype(_lib.EVENT_DTYPE)).reshape(-1)
        return c.tolist(), flat
    return c.tolist(), payload.tolist()


def bind_process_to_gpu(index: int, uuid=None):
    """One process per GPU: run this process on the CPUs NVML names as closest to its GPU.  Call it before anything is
    allocated, so that pinned staging buffers (HostFeeder) land on the GPU's NUMA node (first touch).  Matters on boxes
    whose GPUs hang off different sockets; on the B200 pool this was developed on NVML names the same CPUs for every GPU
    and an A/B showed no difference.  Returns the number of CPUs bound to, or None when NVML has no affinity to offer."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}".encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None
