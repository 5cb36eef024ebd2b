"""Multi-GPU plumbing on the CPU: static stream sharding and the end-of-run reduction over a
world_size-2 gloo process group (the only collective of the design; the per-frame path has none)."""

import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shards_partition_the_streams(pkg):
    sh = pkg.sharding
    for total, world in [(512, 8), (512, 4), (512, 2), (64, 1), (10, 3), (5, 8)]:
        blocks = [sh.shard_streams(total, world, r) for r in range(world)]
        flat = [s for b in blocks for s in b]
        assert flat == list(range(total))                               # contiguous, ordered, complete
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1
        for s in (0, total // 2, total - 1):
            assert s in blocks[sh.owner_of(s, total, world)]
    assert list(sh.shard_streams(512, 8, 3)) == list(range(192, 256))   # BASELINE config 4: 64 streams / GPU
    with pytest.raises(ValueError):
        sh.shard_streams(8, 2, 2)


def test_reduce_summary_without_process_group(pkg):
    totals, ev = pkg.sharding.reduce_summary([64, 10, 3, 2], [1, 0, 1])
    assert totals == [64, 10, 3, 2] and ev == [1, 0, 1]


def _worker(rank, world, port, total_streams, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import importlib
    sharding = importlib.import_module("rtmodt_b200").sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_streams(total_streams, world, rank)
    # every rank "processes" its own streams: stream s produces s % 4 events and 10 + s detections
    counters = [len(mine) * 7, sum(10 + s for s in mine), len(mine), sum(s % 4 for s in mine)]
    totals, per_stream = sharding.reduce_summary(counters, [s % 4 for s in mine])
    out.put((rank, totals, per_stream))
    dist.barrier()
    dist.destroy_process_group()


def test_reduction_over_gloo_world_size_2(pkg):
    """The result of the run summary does not depend on how streams were split over ranks."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 11                                                        # uneven split: 6 + 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect_totals = [total * 7, sum(10 + s for s in range(total)), total, sum(s % 4 for s in range(total))]
    for rank, totals, per_stream in res:
        assert totals == expect_totals
        assert per_stream == [s % 4 for s in range(total)]            # global stream order


def test_workload_is_identical_wherever_a_stream_is_generated(pkg):
    """Stream s gets the same planted cells and zones whether it is stream 0 of rank 1 or stream
    s of a single-rank run (the seed depends on the global stream id only)."""
    import torch
    from rtmodt_b200.workload import PostBackboneWorkload
    whole = PostBackboneWorkload(3, 2, first_stream=4, device="cpu", dtype=torch.float32)
    part = PostBackboneWorkload(1, 2, first_stream=6, device="cpu", dtype=torch.float32)
    assert whole.zones[2] == part.zones[0]
    for k in ("boxes", "cls", "logit", "present"):
        np.testing.assert_array_equal(whole.objects[2][k], part.objects[0][k])


def _record_worker(rank, world, port, total_streams, out):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    import importlib
    pkg = importlib.import_module("rtmodt_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = pkg.sharding.shard_streams(total_streams, world, rank)
    # stream s (global) emits s % 3 events; a rank numbers its streams from 0
    rec = np.zeros(sum(s % 3 for s in mine), np.dtype(pkg._lib.EVENT_DTYPE))
    k = 0
    for local, s in enumerate(mine):
        for e in range(s % 3):
            rec[k]["stream"], rec[k]["track_id"], rec[k]["zone"], rec[k]["dwell"] = local, 100 * s + e, e, 0.25 * s
            rec[k]["xyxy"] = (s, e, s + 1, e + 1)
            k += 1
    totals, gathered = pkg.sharding.reduce_summary([len(mine), len(rec)], rec, first_stream=mine.start)
    out.put((rank, totals, gathered.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_event_records_are_gathered_over_gloo_world_size_2(pkg):
    """the padded 64-byte event records of every rank, in global stream order on every rank."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 9                                                         # 5 + 4 streams, different record counts per rank
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_record_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, totals, raw in res:
        rec = np.frombuffer(raw, np.dtype(pkg._lib.EVENT_DTYPE))
        assert totals == [total, sum(s % 3 for s in range(total))]
        assert rec["stream"].tolist() == [s for s in range(total) for _ in range(s % 3)]      # global ids, stream order
        assert rec["track_id"].tolist() == [100 * s + e for s in range(total) for e in range(s % 3)]
        assert rec["xyxy"][:, 2].tolist() == [s + 1 for s in range(total) for _ in range(s % 3)]
