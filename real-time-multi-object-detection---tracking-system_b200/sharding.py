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


def reduce_summary(counters: Sequence[int], events_per_stream, device="cpu"):
    """Sum ``counters`` over ranks and gather every rank's per-stream event counts.

    Returns ``(totals list[int], per_stream_events list[int] in global stream order)``.  With no
    process group initialised it is the identity (single-GPU runs).
    """
    import torch
    import torch.distributed as dist
    c = torch.tensor(list(counters), dtype=torch.int64, device=device)
    ev = torch.as_tensor(events_per_stream, dtype=torch.int64).to(device).reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return c.tolist(), ev.tolist()
    world = dist.get_world_size()
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([ev.numel()], dtype=torch.int64, device=device))
    width = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(width, dtype=torch.int64, device=device)
    padded[: ev.numel()] = ev
    gathered = [torch.zeros(width, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(gathered, padded)
    flat = []
    for g, s in zip(gathered, sizes):
        flat += g[: int(s.item())].tolist()
    return c.tolist(), flat
