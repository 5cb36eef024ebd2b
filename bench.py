#!/usr/bin/env python
"""bench.py - tracked frames/s of the post-backbone hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

Workload (config.workload).  N = 1: BASELINE.json configs[2] - 64 synthetic 1080p streams, YOLOv8 head
tensors of 8400 anchors x (64 + 80) channels per stream-frame (planted objects, bf16 like the
reference's half=True network output), class-aware NMS, ByteTrack-style association and 4 polygon
zones per stream.  N > 1: configs[3] - 512 such streams in contiguous shards of 512 / N per GPU
(256 / 128 / 64), no per-frame collective; one NCCL reduction of the run summary and one gather of the
event records happen after the timed region.  One step = one frame of every stream of the rank:
decode + NMS -> tracker -> zones, ONE kernel launch.  Every stream's tensors depend on its global id
only, so both arms and every sharding see identical inputs.

The JSON line carries: `value` (device-resident inputs, CUDA-event timed, max over ranks), `e2e` (same
step through rtm_post_backbone_step_host: pinned host head tensors -> H2D -> kernel -> D2H of
detections / events, copies inside the timed region, with the bare-copy ceiling of the same run beside
it), `roofline` (algorithmic head bytes / the step kernel's average duration in the timed region vs the
measured HBM peak), `cpu_baseline` (the CPU chain on one host core over a bounded sample), latency
percentiles, clocks, and `parity`: every rank checks its first 64 streams x 16 frames against the oracle
chain, and rank 0 re-runs streams of the other ranks' shards to show N ranks = 1 rank bit for bit.
"""

from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tracked_frames_per_sec_post_backbone"
UNIT = "frames/s"
WANTED = [0, 1, 2, 3, 5, 7]
STREAMS_1GPU = 64          # configs[2]
TOTAL_STREAMS_MULTI = 512  # configs[3]
CYCLE_FRAMES = 16
FPS = 30.0
T0 = 1_700_000_000.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default: 64 on one GPU, 512 / N on N)")
    ap.add_argument("--frames", type=int, default=CYCLE_FRAMES, help="distinct frames in the replayed cycle")
    ap.add_argument("--head-dtype", default="bf16", choices=["bf16", "f16", "f32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip letterbox / latency / dense-crowd / opt-in mode side measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity leg (experiments only)")
    a = ap.parse_args()
    if a.streams <= 0:
        a.streams = STREAMS_1GPU if a.gpus <= 1 else max(1, TOTAL_STREAMS_MULTI // a.gpus)
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu
    --set full capture (profiles/traffic.json, written by tools/ncu_summary.py), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[kernel]
        return t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]
    except Exception:
        return None


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, uuid=None) -> None:
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}".encode() if uuid else b"")
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------
# the oracle chain on the CPU (cpu_baseline leg and the reference arm)
# ---------------------------------------------------------------------------
def load_reference_modules():
    """The UNMODIFIED reference tracker / zone engine from baseline/_ref (pip --no-deps install of
    /root/reference), or None.  Its detector cannot be imported anywhere: it needs
    ultralytics, which is not installable offline."""
    path = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(path, "src", "tracking", "tracker.py")):
        return None
    try:
        if path not in sys.path:
            sys.path.insert(0, path)
        from loguru import logger
        logger.remove()
        rt = importlib.import_module("src.tracking.tracker")
        rz = importlib.import_module("src.events.zone_engine")
        return rt, rz
    except Exception:
        return None


class CpuStream:
    """One stream of the CPU arm: oracle port of ultralytics' decode + NMS + rescale (torch CPU,
    torchvision.ops.nms), then the reference's own MultiObjectTracker.update and
    ZoneEventEngine.process when baseline/_ref is present, else their oracle restatements."""

    def __init__(self, zones, ref=None, tmpdir=None):
        import types
        from oracle import tracker_ref, zone_ref
        self.ref, self.types = ref, types
        if ref is not None:
            rt, rz = ref
            self.trk = rt.MultiObjectTracker("bytetrack", bytetrack=dict(track_thresh=0.5, track_buffer=30, match_thresh=0.8))
            self.eng = rz.ZoneEventEngine(zones, log_path=os.path.join(tmpdir or "/tmp", f"rtm_events_{os.getpid()}_{id(self)}.jsonl"))
            self.clock = [0.0]
            rz.time.time = lambda c=self.clock: c[0]         # zone_engine.py:84 reads time.time()
        else:
            self.trk = tracker_ref.TrackerOracle()
            self.eng = zone_ref.ZoneOracle(zones, pip=zone_ref.cv2_pip)

    def step(self, heads_f32, frame_id, now):
        from oracle import detect_ref
        det = detect_ref.detect_post(heads_f32, (1080, 1920), classes=WANTED)[0]
        if self.ref is not None:
            self.ref[1].time.time = lambda c=self.clock: c[0]
            self.clock[0] = now
            out = self.trk.update(self.types.SimpleNamespace(xyxy=det["xyxy"], confidence=det["conf"], class_id=det["cls"]))
            assert out == []
            active = [self.types.SimpleNamespace(track_id=t["track_id"], xyxy=t["xyxy"], class_id=t["class_id"])
                      for t in self.trk._core._tracks if t["time_since_update"] == 1]
            ev = self.eng.process(active, frame_id)
        else:
            self.trk.step(det["xyxy"], det["conf"], det["cls"])
            act = self.trk.active_rows()
            ev = self.eng.process(zip(self.trk.track_id[act], self.trk.xyxy[act], self.trk.cls[act]), frame_id, now)
        return len(det["conf"]), len(ev)


def cpu_chain_description(ref):
    tail = ("the reference's unmodified MultiObjectTracker.update + ZoneEventEngine.process (baseline/_ref)" if ref is not None
            else "oracle restatements of the reference tracker / zone engine (baseline/_ref absent)")
    return "oracle port of ultralytics decode + non_max_suppression + scale_boxes (torch CPU, torchvision.ops.nms; ultralytics is not installable) -> " + tail


def cpu_port_throughput(host_frames, zones, seconds):
    """frames/s of the CPU chain on ONE thread over `host_frames[f][level] (S,144,h,w)` replayed."""
    import tempfile
    import torch
    torch.set_num_threads(1)
    ref = load_reference_modules()
    S = host_frames[0][0].shape[0]
    tmp = tempfile.mkdtemp(prefix="rtm_cpu_")
    streams = [CpuStream(zones[s], ref, tmp) for s in range(S)]
    done, f = 0, 0
    t0 = time.perf_counter()
    while True:
        heads = host_frames[f % len(host_frames)]
        for s in range(S):
            streams[s].step([h[s:s + 1] for h in heads], f, T0 + f / FPS)
        done += S
        f += 1
        if time.perf_counter() - t0 >= seconds and f >= 2:
            break
    return done / (time.perf_counter() - t0), done, ref is not None


def _reference_worker(args):
    """Worker process of the reference arm: owns `streams`, steps them W + K times."""
    (streams, frames, steps, warmup, threads, barrier_path, head_dtype) = args
    import torch
    torch.set_num_threads(threads)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    importlib.import_module("rtmodt_b200")
    from rtmodt_b200.workload import PostBackboneWorkload
    ref = load_reference_modules()
    # the very tensors the CUDA arm reads: generated per global stream id, rounded to the head dtype, widened to f32
    tdt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[head_dtype]
    wl = [PostBackboneWorkload(1, frames, first_stream=s, device="cpu", dtype=tdt) for s in streams]
    for w in wl:
        w.heads = [[t.float() for t in fr] for fr in w.heads]
    cpu = [CpuStream(w.zones[0], ref, os.path.dirname(barrier_path)) for w in wl]
    dets = evs = 0

    def one(f):
        nonlocal dets, evs
        for w, c in zip(wl, cpu):
            d, e = c.step(w.heads[f % frames], f, T0 + f / FPS)
            dets += d
            evs += e

    for f in range(warmup):
        one(f)
    # file-based barrier: all workers start the timed region together
    # NB: the zone engine's clock is injected by replacing time.time (zone_engine.py:84 reads it), which is the
    # process-wide function - wall time is therefore taken from the monotonic clock (system-wide on Linux, so
    # the workers' readings are comparable)
    open(f"{barrier_path}.{streams[0]}", "w").close()
    deadline = time.monotonic() + 600
    want = int(open(barrier_path).read())
    while len([n for n in os.listdir(os.path.dirname(barrier_path)) if n.startswith(os.path.basename(barrier_path) + ".")]) < want:
        if time.monotonic() > deadline:
            raise RuntimeError("reference arm: workers did not reach the barrier")
        time.sleep(0.005)
    t0 = time.monotonic()
    for f in range(warmup, warmup + steps):
        one(f)
    t1 = time.monotonic()
    return t0, t1, len(streams) * steps, dets, evs, ref is not None


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import tempfile
    cores = os.cpu_count() or 1
    total_streams = args.streams * args.gpus
    workers = max(1, min(cores, total_streams))
    threads = max(1, cores // workers)
    shards = [list(range(total_streams))[w::workers] for w in range(workers)]
    tmp = tempfile.mkdtemp(prefix="rtm_ref_")
    barrier = os.path.join(tmp, "barrier")
    with open(barrier, "w") as f:
        f.write(str(workers))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        res = pool.map(_reference_worker, [(s, args.frames, args.steps, args.warmup, threads, barrier, args.head_dtype) for s in shards])
    t0 = min(r[0] for r in res)
    t1 = max(r[1] for r in res)
    frames = sum(r[2] for r in res)
    value = frames / (t1 - t0)
    used_ref = all(r[5] for r in res)
    sample = (f"{total_streams} streams x {args.steps} frames (after {args.warmup} warm-up frames) of the same seeded workload, "
              f"{workers} worker processes x {threads} torch threads on {cores} host cores; "
              + cpu_chain_description(True if used_ref else None))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (t1 - t0) / args.steps, "higher_is_better": True, "scaling": scaling_label(args.gpus),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, total_streams),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers * threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "detections": sum(r[3] for r in res), "events": sum(r[4] for r in res),
    }
    emit(line)


def scaling_label(n_gpus):
    # N = 1 is configs[2] (64 streams); N = 2 / 4 / 8 all run configs[3], a fixed total of 512 streams
    return "weak" if n_gpus <= 1 else "strong"


def workload_config(args, total_streams):
    if args.gpus <= 1:
        name = ("BASELINE.json configs[2]: %d synthetic 1080p streams on one GPU, 8400 anchors x 80 classes, ByteTrack + 4 zones"
                % total_streams)
    else:
        name = ("BASELINE.json configs[3]: %d synthetic 1080p streams sharded over %d GPUs (%d per GPU, contiguous), full "
                "post-backbone pipeline: 8400 anchors x 80 classes, ByteTrack + 4 zones" % (total_streams, args.gpus, args.streams))
    return {"workload": name,
            "streams_per_gpu": args.streams, "total_streams": total_streams, "head_dtype": args.head_dtype,
            "anchors": 8400, "classes": 80, "zones_per_stream": 4, "objects_per_stream": 30,
            "cycle_frames": args.frames, "class_filter": WANTED, "conf": 0.35, "iou": 0.45, "max_det": 100,
            "track_thresh": 0.5, "match_thresh": 0.8, "track_buffer": 30,
            "scaling_note": "N = 1 runs configs[2] (64 streams); N = 2 / 4 / 8 run configs[3] (512 streams in all, 512 / N per GPU); "
                            "the 64-streams-per-GPU (weak) figure of the same run is in `weak_scaling`",
            "step_mode": "heads declared complete (resident): one step kernel per step on the library's stream, each a programmatic "
                         "dependent of the one before (its head scan overlaps the previous step's NMS / tracker / zone stage)",
            "l2_policy": "inputs larger than L2: every step reads a different frame of the cycle "
                         "(streams x 8400 x 144 head elements per step, cycle of frames resident in HBM)"}


# ---------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ge = importlib.import_module("__graft_entry__")
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    pkg = importlib.import_module("rtmodt_b200")
    from rtmodt_b200.workload import PostBackboneWorkload
    from rtmodt_b200 import sharding, _lib
    affinity = None
    if world > 1 and os.environ.get("RTM_BENCH_AFFINITY", "1") != "0":
        # one process per GPU: run on the CPUs closest to it before the buffers are allocated (pinned staging buffers then
        # land on the GPU's NUMA node).  On this pool it changes nothing - NVML names the same 32 CPUs for every GPU and an
        # A/B at N = 4 gave 53.4 vs 53.5 GB/s per GPU; what differs is the box: 28.7, 40.7 and 53.4 GB/s per GPU in three
        # 4-GPU runs of the same code
        affinity = sharding.bind_process_to_gpu(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None))

    tdt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[args.head_dtype]
    S, F, K, W = args.streams, args.frames, args.steps, max(args.warmup, 3)
    total_streams = S * world
    my_streams = sharding.shard_streams(total_streams, world, rank)
    wl = PostBackboneWorkload(S, F, first_stream=my_streams.start, device=dev, dtype=tdt)
    lib = _lib.lib()
    hbm_peak, peak_src = peaks()

    def new_batch(n, zones):
        return pkg.StreamBatch(n, zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    with sampler:
        # ---- parity: this rank's first 64 streams, every frame of the cycle, against the oracle chain; then the
        #      equivalence of shardings: rank 0 re-runs streams that belong to the other ranks ----
        parity = None
        if not args.no_parity:
            parity = parity_leg(pkg, wl, new_batch, args, rank, world, my_streams, total_streams, dev, tdt)

        sb = new_batch(S, wl.zones)
        state = {"f": 0}

        def run_steps(batch, heads, n, heads_ready=None):
            f = state["f"]
            for i in range(f, f + n):
                batch.step(heads[i % F], now=T0 + i / FPS, frame_id=i, heads_ready=heads_ready)
            state["f"] = f + n

        # ---- timed region: K steps, device-resident inputs, CUDA events ----
        # The head tensors are resident and complete before the region starts, and the step is told so
        # (heads_ready=True -> rtm_step_io.scan_async): every step is one kernel on the library's own stream, a
        # programmatic dependent of the one before; the current stream is made to wait for each of them, so the
        # closing event (current stream) fires after the last step.  The same K steps in the default mode (ordinary
        # launches on the current stream, no overlap between steps) are timed right after, for comparison.
        def timed(batch, heads, heads_ready):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record()
            run_steps(batch, heads, K, heads_ready)
            ev1.record()
            barrier()
            ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            batch.check_status()
            return float(ms.item())

        # warm-up: the W steps asked for, then one untimed rehearsal of the K-step region itself (the first region after
        # other kernels have run is 1 - 1.5 us per step slower than every later one: instruction cache, TLB)
        run_steps(sb, wl.heads, W, True)
        barrier()
        run_steps(sb, wl.heads, K, True)
        ms_total = timed(sb, wl.heads, True)
        value = total_streams * K / (ms_total / 1e3)
        run_steps(sb, wl.heads, W)
        ms_single = timed(sb, wl.heads, None)

        # ---- the other scaling curve: 64 streams per GPU whatever N is (N > 1 only; N = 1 is that already) ----
        weak = None
        if world > 1 and S > STREAMS_1GPU:
            sb64 = new_batch(STREAMS_1GPU, wl.zones[:STREAMS_1GPU])
            heads64 = [[t[:STREAMS_1GPU] for t in fr] for fr in wl.heads]
            run_steps(sb64, heads64, W, True)
            timed(sb64, heads64, True)                           # rehearsal, as for the main region (first sight of every head buffer)
            ms64 = timed(sb64, heads64, True)
            weak = {"streams_per_gpu": STREAMS_1GPU, "total_streams": STREAMS_1GPU * world, "scaling": "weak",
                    "value": STREAMS_1GPU * world * K / (ms64 / 1e3), "ms_per_step": ms64 / K}
            sb64.close()
        elif world > 1:
            weak = {"streams_per_gpu": S, "total_streams": total_streams, "scaling": "weak", "value": value,
                    "ms_per_step": ms_total / K, "note": "this run is the 64-streams-per-GPU point itself"}

        # ---- per-kernel pass: same steps with CUDA events around every launch (rtm_profile_*), each kernel alone ----
        lib.rtm_profile_enable(1)
        _lib.profile_read()
        barrier()
        run_steps(sb, wl.heads, K)
        torch.cuda.synchronize(dev)
        prof = _lib.profile_read()
        lib.rtm_profile_enable(0)
        kernels = {k: {"avg_us": 1e3 * v[0] / v[1], "launches": v[1]} for k, v in prof.items()}
        # The step is ONE kernel (step_kernel: head scan + NMS + tracker + zones; `decode` + `post` on the two-launch
        # fallback).  Timed alone (profiling pass: no overlap between steps) a launch lasts scan + post tail; inside
        # the timed region consecutive launches overlap, so the kernel's average duration THERE is the region's
        # length / launches - that is what `achieved` is computed from.
        top = "step" if "step" in kernels else "decode"
        alone_us = kernels[top]["avg_us"]
        alg_bytes = S * wl.bytes_per_stream_frame
        in_region_us = 1e3 * ms_total / K if top == "step" else alone_us
        achieved = alg_bytes / (in_region_us * 1e-6) / 1e9
        kname = "step_kernel" if top == "step" else "decode_tma_kernel"
        traffic = ncu_traffic(kname)
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": in_region_us, "launch_alone_us": alone_us,
                    "share_of_step": alone_us / sum(v["avg_us"] for v in kernels.values()),
                    "whole_step_frac": (alg_bytes * K / (ms_total / 1e3) / 1e9) / hbm_peak}
        if traffic and top == "step":
            # the step kernel reads the class rows of every anchor but the 64 DFL values only of candidates:
            # what crosses HBM (`traffic`, ncu) is less than the algorithmic bytes (the whole head tensors,
            # which is what the reference reads).  `achieved` / `frac` follow the contract (algorithmic bytes / time);
            # `dram_*` is the same launch on the bytes it really moves - the number to hold against the hardware.
            moved = traffic * S / STREAMS_1GPU                     # the capture is a 64-stream launch
            roofline["dram_achieved"] = moved / (in_region_us * 1e-6) / 1e9
            roofline["dram_frac"] = roofline["dram_achieved"] / hbm_peak
            roofline["note"] = ("lazy box rows: DFL channels are read for candidates only, so traffic < algorithmic bytes; "
                                "frac = algorithmic bytes / time / peak (contract; it can pass 1.0 in long regions for that reason), "
                                "dram_frac = bytes moved / time / peak")

        extras = {}
        if not args.no_extras:
            # ---- latency mode: one step at a time, p50 / p99 of the per-step device time ----
            # (the reference's profiler convention, latency_profiler.py / default.yaml:88: synchronise around the
            # stage, 50 warm-up frames, then percentiles; here over 1000 steps of S stream-frames each)
            lat, ev_seen = [], 0
            for i in range(50 + 1000):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run_steps(sb, wl.heads, 1)
                b.record()
                b.synchronize()
                if i >= 50:
                    lat.append(a.elapsed_time(b))
                    ev_seen += int(sb.zones.event_count.sum().item())
            lat.sort()
            extras["latency_ms_per_step"] = {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                                             "max": lat[-1], "steps": len(lat), "warmup_steps": 50, "streams_per_step": S,
                                             "zone_events_emitted": ev_seen}
            if rank == 0 and world == 1:
                extras["config1_single_stream"] = single_stream_bench(pkg, dev)
                extras["letterbox"] = letterbox_bench(pkg, lib, dev, S, hbm_peak)
                extras["dense_crowd"] = dense_crowd_bench(pkg, dev, hbm_peak)
                extras["opt_in_modes"] = modes_bench(pkg, wl, dev, S, F)

        # ---- e2e: the same step fed from pinned HOST head tensors through the C ABI ----
        e2e = None
        if not args.no_e2e:
            e2e = e2e_bench(pkg, wl, sb, state["f"], K, W, total_streams, world, dev)
            e2e["cpu_affinity"] = (f"rank bound to the {affinity} CPUs NVML names as closest to its GPU (pinned buffers on that node)"
                                   if affinity else "not bound")
            state["f"] += K + W
        sb.check_status()

    # ---- run summary: the only collectives of the run (NCCL all_reduce + all_gather), after the timed regions ----
    tracks, next_id = sb.read_tracks()
    ev_records = sb.event_records()
    totals, gathered = sharding.reduce_summary(
        [S * state["f"], int(sb.det_count.sum().item()), int((next_id - 1).sum()), len(ev_records)], ev_records, device=dev,
        first_stream=my_streams.start)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_streams = list(range(min(4, S)))
        host_frames = [wl.host_frame(ff, sample_streams) for ff in range(F)]
        v, n, used_ref = cpu_port_throughput(host_frames, [wl.zones[s] for s in sample_streams], args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{len(sample_streams)} streams of the same workload, {F}-frame cycle replayed for {n} stream-frames "
                         f"(>= {args.cpu_seconds:.0f} s), 1 thread; " + cpu_chain_description(True if used_ref else None),
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "rehearsal_steps": K,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling_label(world), "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, total_streams),
            "single_stream": {"value": total_streams * K / (ms_single / 1e3), "ms_per_step": ms_single / K,
                              "note": "the same K steps as ordinary launches on the caller's stream (rtm_step_io.scan_async = 0): "
                                      "head tensors ordered on the stream like any input, no overlap between steps"},
            "weak_scaling": weak,
            "clocks": sampler.summary(), "e2e": e2e,
            # kernels of the library inside the timed region: one step kernel per step (the per-kernel pass above
            # counted exactly these launches for the same K steps)
            "gpu_launches": sum(v["launches"] for v in kernels.values()), "roofline": roofline,
            "cpu_baseline": cpu, "kernels": kernels, "parity": parity,
            "summary": dict(zip(sharding.COUNTERS, totals), live_tracks_rank0=int(sum(len(t) for t in tracks)),
                            event_records_gathered=len(gathered)),
        }
        line.update(extras)
        emit(line)
    sb.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_leg(pkg, wl, new_batch, args, rank, world, my_streams, total_streams, dev, tdt, per_rank=64, frames=16):
    """Every rank: its first `per_rank` streams x `frames` frames against the oracle chain (oracle/chain.py).  With
    N > 1 ranks: the digests of what every rank's CUDA path produced are gathered, and rank 0 re-runs two streams of
    every other rank's shard on its own GPU - the same global streams in another batch on another device must give the
    same bits (streams are independent, tracker.py:55-56 / zone_engine.py:72-75; SURVEY section 4 tier 5)."""
    import torch.distributed as dist
    from oracle import chain
    from rtmodt_b200.workload import PostBackboneWorkload
    n = min(per_rank, wl.S)
    frames = min(frames, wl.F)
    heads = [[t[:n] for t in fr] for fr in wl.heads]
    sb = new_batch(n, wl.zones[:n])
    res = chain.run_chain_parity(sb, lambda f: heads[f % wl.F], lambda f: [t.float().cpu() for t in heads[f % wl.F]],
                                 wl.zones[:n], frames, classes=WANTED, t0=T0, fps=FPS)
    sb.close()
    digests = res.pop("digests")
    if not res["ok"]:
        print(f"bench.py: PARITY FAILURE on rank {rank}: {res}", file=sys.stderr)
    out = dict(res, ranks_checked=1, streams_per_rank=n)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {"rank": rank, "first": my_streams.start, "res": res, "digests": digests})
        if rank == 0:
            for k in ("detections_checked", "events_checked", "nms_index_flips", "box_mismatch", "track_id_mismatch",
                      "track_table_mismatch", "event_mismatch"):
                out[k] = sum(g["res"][k] for g in gathered)
            out["streams"] = sum(g["res"]["streams"] for g in gathered)
            out["ok"] = all(g["res"]["ok"] for g in gathered)
            out["ranks_checked"] = world
            # sharding equivalence: two streams of every other rank, re-run here as batches of their own
            picks = [(g["first"] + k, g["digests"][k]) for g in gathered[1:] for k in (0, n - 1)]
            same = 0
            for gid, want in picks:
                w1 = PostBackboneWorkload(1, wl.F, first_stream=gid, device=dev, dtype=tdt)
                sb1 = new_batch(1, w1.zones)
                r1 = chain.run_chain_parity(sb1, lambda f: w1.heads[f % wl.F], lambda f: [t.float().cpu() for t in w1.heads[f % wl.F]],
                                            w1.zones, frames, classes=WANTED, t0=T0, fps=FPS)
                sb1.close()
                same += int(r1["digests"][0] == want)
            out["sharding_equivalence"] = {"streams_rerun_on_rank0": len(picks), "bit_identical": same, "ok": same == len(picks),
                                           "what": "global streams owned by other ranks, re-run on rank 0 as single-stream batches: "
                                                   "sha1 over detections, track ids, track tables and events of every frame"}
            out["ok"] = out["ok"] and same == len(picks)
    return out


def e2e_bench(pkg, wl, sb, f0, K, W, total_streams, world, dev, pinned_frames=4):
    """Same metric through rtm_post_backbone_step_host: pinned host heads -> H2D -> kernel -> D2H.  The step is bound by
    the host -> device copy of its inputs, so the bare copy ceiling of the same buffers is measured in the same run,
    all ranks copying at once like they do in the timed region (`h2d_ceiling_gbs`)."""
    import torch
    import torch.distributed as dist
    feeder = pkg.HostFeeder(sb, wl.dtype)
    n_pin = min(pinned_frames, wl.F)
    order = list(range(n_pin)) + list(range(n_pin - 2, 0, -1))     # forward, then back: every step continues the last
    pinned = []
    for f in range(n_pin):                                    # frames of the cycle in page-locked host memory
        host = feeder.alloc_pinned_heads()
        for dst, src in zip(host, wl.heads[f]):
            dst.copy_(src)
        pinned.append(host)
    torch.cuda.synchronize(dev)
    results = []

    def go(first, n):
        for f in range(first, first + n):
            res = feeder.step_pinned(pinned[order[f % len(order)]], now=T0 + f / FPS, frame_id=f)
            results.append(res)
            if len(results) > 1:
                results.pop(0).wait()                     # consume the results of step k-1 on the host (two slots: step
                                                          # k-2's buffers already belong to step k)
        while results:
            results.pop(0).wait()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    go(f0, W)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    go(f0 + W, K)
    torch.cuda.synchronize(dev)
    dt = max_over_ranks(time.perf_counter() - t0)
    # the bare copies: K times the same pinned buffers into the same device buffers, nothing else
    dev_heads = feeder.slots[0]["dev_heads"]
    copy_stream = torch.cuda.Stream(device=dev)
    def copies(n):
        with torch.cuda.stream(copy_stream):
            for i in range(n):
                host = pinned[i % n_pin]
                if dev_heads[0]._base is not None and host[0]._base is not None:   # the three levels are one buffer on both sides
                    dev_heads[0]._base.copy_(host[0]._base, non_blocking=True)
                else:
                    for d, h in zip(dev_heads, host):
                        d.copy_(h, non_blocking=True)
        copy_stream.synchronize()
    copies(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    copies(K)
    dc = max_over_ranks(time.perf_counter() - t0)
    gbs, ceil_gbs = feeder.h2d_bytes * K / dt / 1e9, feeder.h2d_bytes * K / dc / 1e9
    return {"value": total_streams * K / dt, "unit": UNIT, "h2d_bytes_per_step": feeder.h2d_bytes,
            "d2h_bytes_per_step": feeder.d2h_bytes, "ms_per_step": 1e3 * dt / K,
            "h2d_gbs": gbs, "h2d_ceiling_gbs": ceil_gbs, "frac_of_copy_ceiling": gbs / ceil_gbs,
            "ceiling_note": "bare cudaMemcpyAsync of the same pinned head buffers, same count, every rank at once (max over ranks): "
                            "what the host memory system and the PCIe links of this box give the run",
            "pinned_frames": n_pin,
            "api": "HostFeeder.step_pinned -> rtm_post_backbone_step_host (2 CUDA streams, H2D of step k+1 overlaps step k)"}


def modes_bench(pkg, wl, dev, S, F, steps=100):
    """The same workload with the tracker's opt-in modes (neither is the reference's default behaviour
    here): use_kalman=True and assignment="lapjv", both inside the step kernel, steps back to back (heads_ready)."""
    import torch
    out = {}
    for name, kw in (("kalman", dict(use_kalman=True)), ("lapjv", dict(assignment="lapjv"))):
        sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev, **kw)
        for f in range(20):
            sb.step(wl.heads[f % F], now=T0 + f / FPS, frame_id=f, heads_ready=True)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for f in range(20, 20 + steps):
            sb.step(wl.heads[f % F], now=T0 + f / FPS, frame_id=f, heads_ready=True)
        b.record()
        b.synchronize()
        sb.check_status()
        ms = a.elapsed_time(b) / steps
        out[name] = {"frames_per_s": S / (ms * 1e-3), "ms_per_step": ms}
    return out


def dense_crowd_bench(pkg, dev, hbm_peak, streams=128, objects=1000, zones=16, distinct=8, frames=16, steps=48):
    """BASELINE.json configs[4] as a side measurement: 128 streams x 1000 scripted objects (MOT20-like
    motion) x 16 zone polygons (K in 4..12), tracker + zones only (the detector cannot emit 1000 boxes:
    max_det = 100).  `distinct` different streams are generated and repeated to fill the batch (streams
    are independent, so repetition changes nothing about the work); the clip runs forward then back.
    Every distinct stream is checked against the oracle on every frame of the clip; the two kernels are
    timed alone as well and held against what bounds them (algorithmic bytes; the pair tests)."""
    import numpy as np
    import torch
    from oracle import tracker_ref, zone_ref
    from rtmodt_b200 import _lib
    slots = 1024
    xyxy, conf, cls, count = pkg.synth.scripted_batch(distinct, frames, slots, seed=900, **pkg.synth.dense_crowd_kwargs(objects))
    rep = streams // distinct
    tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.concatenate([a] * rep, axis=1))).to(dev)
    d_xyxy, d_conf, d_cls, d_count = tile(xyxy), tile(conf), tile(cls), tile(count)
    zcfg = [pkg.synth.make_zones(seed=b % distinct, num_zones=zones, width=1920, height=1080, kmin=4, kmax=12) for b in range(streams)]
    order = list(range(frames)) + list(range(frames - 2, 0, -1))                 # forward, then back

    def fresh():
        return pkg.StreamBatch(streams, zcfg, src_hw=(1080, 1920), max_det=slots, max_tracks=4096, max_events=4096, device=dev)

    # parity: every distinct stream, every frame of the clip, against the oracle (tracker pinned to the reference)
    sb = fresh()
    trk = [tracker_ref.TrackerOracle() for _ in range(distinct)]
    zon = [zone_ref.ZoneOracle(zcfg[b]) for b in range(distinct)]
    bad = checked = 0
    for f in range(frames):
        now = T0 + f / FPS
        sb.track_only(d_xyxy[f], d_conf[f], d_cls[f], d_count[f], now=now, frame_id=f)
        tracks, next_id = sb.read_tracks()
        got = sb.read_events()
        for b in range(distinct):
            n = int(count[f, b])
            trk[b].step(xyxy[f, b, :n], conf[f, b, :n], cls[f, b, :n])
            ok = [t["track_id"] for t in tracks[b]] == trk[b].track_id.tolist() and int(next_id[b]) == trk[b].next_id
            ok = ok and [t["time_since_update"] for t in tracks[b]] == trk[b].tsu.tolist()
            ok = ok and np.array_equal(np.asarray([t["xyxy"] for t in tracks[b]], np.float32).reshape(-1, 4), trk[b].xyxy)
            act = trk[b].active_rows()
            exp = zon[b].process(zip(trk[b].track_id[act], trk[b].xyxy[act], trk[b].cls[act]), f, now)
            ok = ok and [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec) for e in got[b]] == \
                [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec) for e in exp]
            bad += int(not ok)
            checked += 1
    sb.check_status()
    sb.close()
    # timing
    sb = fresh()
    k = 0
    def go(n):
        nonlocal k
        for _ in range(n):
            f = order[k % len(order)]
            sb.track_only(d_xyxy[f], d_conf[f], d_cls[f], d_count[f], now=T0 + k / FPS, frame_id=k)
            k += 1
    go(len(order))
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    go(steps)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / steps
    sb.lib.rtm_profile_enable(1)
    _lib.profile_read()
    go(steps)
    torch.cuda.synchronize(dev)
    prof = _lib.profile_read()
    sb.lib.rtm_profile_enable(0)
    sb.check_status()
    tracks, _ = sb.read_tracks()
    T = float(np.mean([len(t) for t in tracks]))
    N = float(count.mean())
    kern = {n: 1e3 * v[0] / v[1] for n, v in prof.items()}
    # track step 36 (T_in + T_out) + 24 N bytes, zone step 24 T + 8 Z K + 2 * 16 T Z bytes per stream-frame
    alg_track = streams * (36 * 2 * T + 24 * N)
    alg_zone = streams * (24 * T + 8 * zones * 8 + 32 * T * zones)
    pair_tests = streams * T * N                           # IoU pairs of stage 1 (stage 2 adds the unmatched rest)
    out = {"config": f"BASELINE.json configs[4]: {streams} streams x {objects} scripted objects, {zones} zones (K 4..12), tracker + zones "
                     f"({distinct} distinct streams repeated)", "frames_per_s": streams / (ms * 1e-3), "ms_per_step": ms,
           "live_tracks_per_stream": T, "detections_per_stream": N,
           "parity_ok": bad == 0, "parity": {"stream_frames_checked": checked, "mismatches": bad, "frames": frames, "streams": distinct},
           "steps": steps, "kernels_alone_us": kern,
           "roofline": {"bound": "hbm (algorithmic bytes, SURVEY 8d) - the kernels are nowhere near it: the step is pair tests and per-row "
                                 "serial logic, not bytes",
                        "algorithmic_bytes_per_step": alg_track + alg_zone,
                        "achieved_gbs": (alg_track + alg_zone) / (ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": (alg_track + alg_zone) / (ms * 1e-3) / 1e9 / hbm_peak,
                        "iou_pair_tests_per_step": pair_tests,
                        "pair_tests_per_s": pair_tests / (kern.get("track", ms * 1e3) * 1e-6)}}
    sb.close()
    return out


def single_stream_bench(pkg, dev, steps=300):
    """BASELINE.json configs[1]: ONE stream, batch 1 - the YOLOv8s stand-in's conv forward on a 640 x 640 bf16 frame
    (random init, not a parity subject: its heads score nothing above 0.35), then decode + NMS + tracker + zones on a
    planted head tensor of the same shape.  What the reference's loop does per frame
    (tools/run_pipeline.py:131-146), one frame at a time, latency of each part by CUDA events."""
    import numpy as np
    import torch
    from rtmodt_b200.detection.yolov8s import YOLOv8s
    from rtmodt_b200.workload import PostBackboneWorkload
    torch.manual_seed(0)
    net = YOLOv8s().to(dev).to(torch.bfloat16).eval()
    frame = torch.rand(1, 3, 640, 640, device=dev).to(torch.bfloat16)
    wl = PostBackboneWorkload(1, 8, first_stream=0, device=dev, dtype=torch.bfloat16)
    sb = pkg.StreamBatch(1, wl.zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)
    t_net, t_post = [], []
    with torch.no_grad():
        for i in range(20 + steps):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            heads = net(frame)
            b.record()
            sb.step(wl.heads[i % 8], now=T0 + i / FPS, frame_id=i)
            c.record()
            c.synchronize()
            if i >= 20:
                t_net.append(a.elapsed_time(b))
                t_post.append(b.elapsed_time(c))
    sb.check_status()
    dets = int(sb.det_count.sum().item())
    sb.close()
    assert [tuple(h.shape) for h in heads] == [(1, 144, 80, 80), (1, 144, 40, 40), (1, 144, 20, 20)]
    pct = lambda v, q: float(np.percentile(np.asarray(v), q))
    return {"config": "BASELINE.json configs[1]: YOLOv8s stand-in (random init) 640x640 batch 1, one stream: conv forward (PyTorch, "
                      "not timed against anything), then decode + NMS + ByteTrack + 4 zones on a planted head tensor",
            "post_backbone_ms": {"p50": pct(t_post, 50), "p99": pct(t_post, 99)},
            "conv_forward_ms": {"p50": pct(t_net, 50), "p99": pct(t_net, 99)},
            "frames_per_s_post_backbone": 1e3 / pct(t_post, 50), "steps": steps, "detections_last_frame": dets}


def letterbox_bench(pkg, lib, dev, S, hbm_peak, iters=20):
    """P1 on its own: S 1080p u8 frames -> (S, 3, 640, 640) bf16 (reported beside the main metric)."""
    import torch
    from rtmodt_b200 import _lib
    frames = torch.randint(0, 256, (S, 1080, 1920, 3), dtype=torch.uint8, device=dev)
    out = torch.empty((S, 3, 640, 640), dtype=torch.bfloat16, device=dev)
    call = lambda: _lib.check(lib.rtm_letterbox(frames.data_ptr(), S, 1080, 1920, 1920 * 3, 1080 * 1920 * 3, out.data_ptr(),
                                                _lib.RTM_BF16, 640, 640, _lib.cuda_stream()))
    for _ in range(3):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(iters):
        call()
    b.record()
    b.synchronize()
    us = 1e3 * a.elapsed_time(b) / iters
    alg = S * (640 * 360 * 3 + 3 * 640 * 640 * 2)          # contributing source pixels + bf16 output
    gbs = alg / (us * 1e-6) / 1e9
    return {"frames_per_s": S / (us * 1e-6), "avg_launch_us": us, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak,
            "algorithmic_bytes_per_launch": alg, "config": f"{S} x 1080p u8 BGR -> 640x640 bf16 CHW"}


def emit(line) -> None:
    """The one JSON line of the run, on the process's original stdout."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


_STDOUT_FD = 1

if __name__ == "__main__":
    a = parse_args()
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner under torchrun,
    # for one) goes to stderr instead
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
