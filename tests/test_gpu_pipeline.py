"""The fused post-backbone step (rtm_post_backbone_step) against the oracle chain."""

import numpy as np
import pytest

from oracle import detect_ref, tracker_ref, zone_ref

pytestmark = pytest.mark.gpu

WANTED = [0, 1, 2, 3, 5, 7]


def moving_heads(pkg, B, F, seed, n_obj=12):
    """F frames of planted head tensors whose objects move slowly (letterbox coordinates)."""
    rng = np.random.default_rng(seed)
    frames = []
    wh = np.stack([rng.uniform(60, 150, (B, n_obj)), rng.uniform(80, 200, (B, n_obj))], -1)
    c = np.stack([rng.uniform(100, 540, (B, n_obj)), rng.uniform(200, 440, (B, n_obj))], -1)
    v = rng.uniform(-1.5, 1.5, (B, n_obj, 2))
    cls = rng.choice(np.asarray(WANTED), (B, n_obj))
    for f in range(F):
        c = c + v
        levels = [[], [], []]
        for b in range(B):
            boxes = np.concatenate([c[b] - wh[b] / 2, c[b] + wh[b] / 2], 1)
            for l, t in enumerate(pkg.synth.plant_head(rng, boxes, cls[b], distractor_frac=0.0)):
                levels[l].append(t)
        frames.append([np.stack(l) for l in levels])
    return frames


@pytest.mark.parametrize("dtype,kalman,assignment", [("f32", False, "greedy"), ("bf16", False, "greedy"), ("bf16", True, "greedy"),
                                                     ("bf16", False, "lapjv"), ("bf16", True, "lapjv")])
def test_fused_step_matches_oracle_chain(pkg, dtype, kalman, assignment):
    import torch
    B, F = 4, 25
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    zones = [pkg.synth.make_zones(seed=b, num_zones=4, width=1920, height=1080, dwell_time_sec=0.2, cooldown_sec=0.4)
             for b in range(B)]
    sb = pkg.StreamBatch(B, zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=256, use_kalman=kalman, assignment=assignment)
    assign = tracker_ref.assign_lapjv_emulated if assignment == "lapjv" else tracker_ref.assign_rowloop
    trk = [tracker_ref.TrackerOracle(use_kalman=kalman, assign=assign) for _ in range(B)]
    zon = [zone_ref.ZoneOracle(z) for z in zones]
    n_events = 0
    for f, heads in enumerate(moving_heads(pkg, B, F, seed=4)):
        ht = [torch.from_numpy(h).to(tdt) for h in heads]
        now = 50.0 + f / 30.0
        sb.step([h.to(sb.device).contiguous() for h in ht], now=now, frame_id=f)
        got_det = sb.read_detections()
        got_trk, got_next = sb.read_tracks()
        got_ev = sb.read_events()
        ref_det = detect_ref.detect_post([h.float() for h in ht], (1080, 1920), classes=WANTED)
        for b in range(B):
            r = ref_det[b]
            assert len(r["conf"]) == len(got_det[b]["confidence"]) > 0
            np.testing.assert_array_equal(got_det[b]["anchor"], r["anchor"])
            np.testing.assert_array_equal(got_det[b]["class_id"], r["cls"])
            np.testing.assert_allclose(got_det[b]["xyxy"], r["xyxy"], rtol=1e-4, atol=1e-2)
            # the tracker / zone oracles consume the DEVICE detections, so that the stages after the
            # (tolerance-checked) float decode are compared bit for bit
            tid, kind = trk[b].step(got_det[b]["xyxy"], got_det[b]["confidence"], got_det[b]["class_id"])
            np.testing.assert_array_equal(got_det[b]["track_id"], tid)
            np.testing.assert_array_equal(got_det[b]["kind"], kind)
            o = trk[b]
            assert int(got_next[b]) == o.next_id
            assert [t["track_id"] for t in got_trk[b]] == o.track_id.tolist()
            np.testing.assert_array_equal(np.array([t["xyxy"] for t in got_trk[b]], np.float32).reshape(-1, 4), o.xyxy)
            assert [t["time_since_update"] for t in got_trk[b]] == o.tsu.tolist()
            assert [t["age"] for t in got_trk[b]] == o.age.tolist()
            act = o.active_rows()
            exp_ev = zon[b].process(zip(o.track_id[act], o.xyxy[act], o.cls[act]), f, now)
            assert [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec, e.bbox_xyxy, e.frame_id) for e in got_ev[b]] == \
                   [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec, e.bbox_xyxy, e.frame_id) for e in exp_ev]
            n_events += len(exp_ev)
            if kalman:
                host = sb.table.to_host()
                np.testing.assert_array_equal(host["kf_mean"][b, :len(o)], o.kf_mean)
                np.testing.assert_array_equal(host["kf_cov"][b, :len(o)], o.kf_cov)
    assert n_events > 0


def test_host_fed_step_equals_device_step(pkg):
    """rtm_post_backbone_step_host (pinned host heads in, events out) == the device-resident step."""
    import torch
    B, F = 3, 6
    zones = [pkg.synth.make_zones(seed=b, num_zones=3, width=1920, height=1080, dwell_time_sec=0.0, cooldown_sec=0.1)
             for b in range(B)]
    a = pkg.StreamBatch(B, zones, classes=WANTED, max_tracks=128)
    h = pkg.StreamBatch(B, zones, classes=WANTED, max_tracks=128)
    feeder = pkg.HostFeeder(h, torch.bfloat16)
    for f, heads in enumerate(moving_heads(pkg, B, F, seed=9)):
        ht = [torch.from_numpy(x).to(torch.bfloat16) for x in heads]
        a.step([x.to(a.device).contiguous() for x in ht], now=10.0 + f, frame_id=f)
        res = feeder.step(ht, now=10.0 + f, frame_id=f)
        ev_a = a.read_events()
        ev_h = res.events()
        assert [[(e.track_id, e.zone_name, e.centroid, e.bbox_xyxy) for e in s] for s in ev_a] == \
               [[(e.track_id, e.zone_name, e.centroid, e.bbox_xyxy) for e in s] for s in ev_h]
        np.testing.assert_array_equal(res.det_count, a.det_count.cpu().numpy())
    assert sum(len(s) for s in ev_a) >= 0


@pytest.mark.parametrize("S", [6, 24])          # 6 streams: the scan of step k+1 is over long before the post kernel of step k
def test_back_to_back_steps_equal_synchronised_steps(pkg, S):
    """Steps enqueued without synchronisation overlap on the device (the head scan of a step runs
    beside the post kernel of the step before: programmatic dependent launch over a ring of
    candidate-list slots).  The state they leave must be that of the same steps run one at a time:
    every frame's detections feed the tracker, so the final tables pin the whole sequence."""
    import torch
    from rtmodt_b200.workload import PostBackboneWorkload
    F = 8
    dev = torch.device("cuda", 0)
    wl = PostBackboneWorkload(S, F, first_stream=0, device=dev, dtype=torch.bfloat16)

    def run(sync, steps):
        sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)
        for f in range(steps):
            sb.step(wl.heads[f % F], now=1.7e9 + f / 30.0, frame_id=f)      # nothing else goes on the stream
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        sb.check_status()
        state = sb.table.to_host()
        n = state["count"]
        rows = {k: [v[b, :n[b]].copy() for b in range(S)] for k, v in state.items() if k not in ("count", "next_id")}
        cnt = sb.det_count.cpu().numpy()
        dets = [(sb.det_xyxy[b, :cnt[b]].cpu().numpy(), sb.det_track_id[b, :cnt[b]].cpu().numpy()) for b in range(S)]
        ec = sb.zones.event_count.cpu().numpy()
        evs = [sb.zones.events[b, :ec[b]].cpu().numpy() for b in range(S)]
        zstate = [t.cpu().numpy() for t in sb.zones.state_in()[:2]]
        return dict(count=n, next_id=state["next_id"], rows=rows, det_count=cnt, dets=dets, ev_count=ec, evs=evs, zstate=zstate)

    def same(a, b):
        assert np.array_equal(a["count"], b["count"]) and np.array_equal(a["next_id"], b["next_id"])
        assert np.array_equal(a["det_count"], b["det_count"]) and np.array_equal(a["ev_count"], b["ev_count"])
        for k in a["rows"]:
            for x, y in zip(a["rows"][k], b["rows"][k]):
                np.testing.assert_array_equal(x, y, err_msg=k)
        for (x0, t0), (x1, t1) in zip(a["dets"], b["dets"]):
            np.testing.assert_array_equal(x0, x1)
            np.testing.assert_array_equal(t0, t1)
        for x, y in zip(a["evs"], b["evs"]):
            np.testing.assert_array_equal(x, y)

    for steps in (37, 96):
        ref = run(True, steps)
        assert ref["det_count"].sum() > 0 and ref["count"].sum() > 0
        for attempt in range(3):                   # overlap is timing dependent: look more than once
            same(ref, run(False, steps))


def test_interleaved_stream_batches_do_not_disturb_each_other(pkg):
    """Three stream batches (own workspaces) stepped round-robin without synchronisation end in the
    state each reaches when it runs alone: the candidate-list slots rotate per workspace."""
    import torch
    from rtmodt_b200.workload import PostBackboneWorkload
    dev = torch.device("cuda", 0)
    S, F, steps = 8, 6, 40
    wls = [PostBackboneWorkload(S, F, first_stream=100 * k, device=dev, dtype=torch.bfloat16) for k in range(3)]

    def make(k):
        return pkg.StreamBatch(S, wls[k].zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)

    def final(sb):
        torch.cuda.synchronize()
        sb.check_status()
        h = sb.table.to_host()
        return h["count"].copy(), h["next_id"].copy(), [h["track_id"][b, :h["count"][b]].copy() for b in range(S)], \
            [h["xyxy"][b, :h["count"][b]].copy() for b in range(S)]

    alone = []
    for k in range(3):
        sb = make(k)
        for f in range(steps):
            sb.step(wls[k].heads[f % F], now=1.7e9 + f / 30.0, frame_id=f)
            torch.cuda.synchronize()
        alone.append(final(sb))
    batches = [make(k) for k in range(3)]
    for f in range(steps):
        for k in range(3):
            batches[k].step(wls[k].heads[f % F], now=1.7e9 + f / 30.0, frame_id=f)
    for k in range(3):
        got = final(batches[k])
        assert np.array_equal(got[0], alone[k][0]) and np.array_equal(got[1], alone[k][1])
        for a, b in zip(got[2] + got[3], alone[k][2] + alone[k][3]):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("mode", ["resident", "event", "mixed"])
def test_async_scan_steps_equal_synchronised_steps(pkg, mode):
    """rtm_step_io.scan_async: the head scan runs on a stream of the library's own, behind a 'heads
    ready' event, so that consecutive scans are back to back.  Same final state as one step at a time -
    with resident heads, with heads copied in on a side stream (event per frame), and when the two
    modes are mixed on one batch."""
    import torch
    from rtmodt_b200.workload import PostBackboneWorkload
    S, F, steps = 12, 8, 60
    dev = torch.device("cuda", 0)
    wl = PostBackboneWorkload(S, F, first_stream=40, device=dev, dtype=torch.bfloat16)

    def final(sb):
        torch.cuda.synchronize()
        sb.check_status()
        h = sb.table.to_host()
        n = h["count"]
        ec = sb.zones.event_count.cpu().numpy()
        return (n.copy(), h["next_id"].copy(), [h["track_id"][b, :n[b]].copy() for b in range(S)],
                [h["xyxy"][b, :n[b]].copy() for b in range(S)], ec.copy(), [sb.zones.events[b, :ec[b]].cpu().numpy() for b in range(S)])

    ref = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)
    for f in range(steps):
        ref.step(wl.heads[f % F], now=1.7e9 + f / 30.0, frame_id=f)
        torch.cuda.synchronize()
    want = final(ref)

    for attempt in range(3):
        sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=512, device=dev)
        side = torch.cuda.Stream(device=dev)
        staging = [[torch.empty_like(t) for t in wl.heads[0]] for _ in range(3)]
        for f in range(steps):
            if mode == "resident" or (mode == "mixed" and f % 3 == 0):
                sb.step(wl.heads[f % F], now=1.7e9 + f / 30.0, frame_id=f, heads_ready=True)
            elif mode == "mixed" and f % 3 == 1:
                sb.step(wl.heads[f % F], now=1.7e9 + f / 30.0, frame_id=f)
            else:
                # the frame's heads are produced on a side stream (stand-in for the backbone's stream)
                buf = staging[f % 3]
                if f >= 3:
                    torch.cuda.current_stream().synchronize()          # the buffer's previous frame has been consumed
                with torch.cuda.stream(side):
                    for d, src in zip(buf, wl.heads[f % F]):
                        d.copy_(src, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(side)
                sb.step(buf, now=1.7e9 + f / 30.0, frame_id=f, heads_ready=ready)
        got = final(sb)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[4], want[4])
        for a, b in zip(got[2] + got[3] + got[5], want[2] + want[3] + want[5]):
            np.testing.assert_array_equal(a, b)


def test_heads_produced_right_before_the_step_are_seen_whole(pkg):
    """Default mode: the head tensors are ordered on the current stream like any other input - also when the kernel
    in front of the step is their producer (a conv stack here, plus an elementwise tail), with no synchronisation in
    between.  (The scan used to be a programmatic dependent launch on the caller's stream, which has no ordering
    against a producer that releases its dependents early: ADVICE r1.)"""
    import torch
    B, F = 3, 6
    dev = torch.device("cuda", 0)
    frames = moving_heads(pkg, B, F, seed=21)
    convs = [torch.nn.Conv2d(144, 144, 1, bias=False).to(dev).to(torch.bfloat16) for _ in range(3)]
    for c in convs:
        with torch.no_grad():
            c.weight.copy_(torch.eye(144).reshape(144, 144, 1, 1))         # identity: the producer is real, the values known
    zones = [pkg.synth.make_zones(seed=b, num_zones=3, width=1920, height=1080, dwell_time_sec=0.0, cooldown_sec=0.1) for b in range(B)]

    def run(sync_between):
        sb = pkg.StreamBatch(B, zones, src_hw=(1080, 1920), classes=WANTED, max_tracks=256)
        out = []
        for f, heads in enumerate(frames):
            src = [torch.from_numpy(h).to(torch.bfloat16).to(dev) for h in heads]
            torch.cuda.synchronize()
            with torch.no_grad():
                produced = [(c(s) * 1.0).contiguous() for c, s in zip(convs, src)]   # enqueued right in front of the step
            if sync_between:
                torch.cuda.synchronize()
            sb.step(produced, now=9.0 + f / 30.0, frame_id=f)
            torch.cuda.synchronize()
            out.append((sb.read_detections(), sb.read_tracks()[1].tolist(), [len(e) for e in sb.read_events()]))
        sb.close()
        return out

    ref, got = run(True), run(False)
    for (d0, n0, e0), (d1, n1, e1) in zip(ref, got):
        assert n0 == n1 and e0 == e1
        for a, b in zip(d0, d1):
            for k in ("xyxy", "confidence", "class_id", "anchor", "track_id"):
                np.testing.assert_array_equal(a[k], b[k])
    assert sum(len(d["confidence"]) for d in ref[-1][0]) > 10


def test_recycled_workspace_addresses_start_afresh(pkg):
    """Batches are created and dropped in a loop: the allocator hands the workspace address out again, and the library's
    per-workspace bookkeeping (slot, counters, its own stream) must not survive into the next owner - in both step modes."""
    import torch
    B, F = 4, 5
    frames = moving_heads(pkg, B, F, seed=8)
    dev = torch.device("cuda", 0)
    heads = [[torch.from_numpy(h).to(torch.bfloat16).to(dev).contiguous() for h in fr] for fr in frames]
    seen, ptrs = [], set()
    for rep in range(6):
        sb = pkg.StreamBatch(B, None, src_hw=(1080, 1920), classes=WANTED, max_tracks=128)
        ptrs.add(sb.workspace.data_ptr())
        for f in range(F + rep % 3):                                 # a different number of steps each time: slots differ
            sb.step(heads[f % F], now=1.0 + f, frame_id=f, heads_ready=True if rep % 2 else None)
        torch.cuda.synchronize()
        if rep % 3 == 0:
            seen.append((sb.read_tracks()[1].tolist(), [d["anchor"].tolist() for d in sb.read_detections()]))
        sb.close()
        del sb
    assert seen[0] == seen[1]


def test_host_feeder_copies_event_heads_only_and_says_so(pkg):
    """The host-fed step copies the first `event_prefix` events of every stream back; a stream that emits more raises
    instead of dropping them, and the device buffers are sized for a whole crowd firing at once."""
    import torch
    B, n_obj = 2, 12
    frames = moving_heads(pkg, B, 3, seed=3, n_obj=n_obj)
    everywhere = [[dict(name=f"z{k}", polygon=[[0, 0], [1920, 0], [1920, 1080], [0, 1080]], dwell_time_sec=0.0, cooldown_sec=0.0)
                   for k in range(3)] for _ in range(B)]
    sb = pkg.StreamBatch(B, everywhere, src_hw=(1080, 1920), classes=WANTED, max_tracks=64)
    assert sb.zones.event_stride == 64 * 3
    feeder = pkg.HostFeeder(sb, torch.bfloat16, event_prefix=8)
    res = None
    for f, heads in enumerate(frames):
        res = feeder.step([torch.from_numpy(h).to(torch.bfloat16) for h in heads], now=5.0 + f, frame_id=f)
    res.wait()
    with pytest.raises(pkg.RtmError):
        res.events()                                               # 3 zones x ~12 tracks > 8 per stream
    assert len(sb.read_events()[0]) > 8                            # all of them are on the device
    feeder = pkg.HostFeeder(sb, torch.bfloat16, event_prefix=64)
    res = feeder.step([torch.from_numpy(h).to(torch.bfloat16) for h in frames[-1]], now=9.0, frame_id=9)
    assert [len(e) for e in res.events()] == [len(e) for e in sb.read_events()]


def test_host_feeder_results_outlived_by_their_slot(pkg):
    """A result whose slot has been handed to a later step (two slots: two steps on) is complete but gone: waiting on
    it returns at once - it must not wait for the LATER step, which would stall a pipelined caller (that cost the
    end-to-end bench 4 % once) - and reading it raises instead of returning the later step's data."""
    import torch
    B = 2
    frames = moving_heads(pkg, B, 4, seed=4, n_obj=6)
    sb = pkg.StreamBatch(B, None, src_hw=(1080, 1920), classes=WANTED, max_tracks=64)
    feeder = pkg.HostFeeder(sb, torch.bfloat16)
    res = [feeder.step([torch.from_numpy(h).to(torch.bfloat16) for h in heads], now=5.0 + f, frame_id=f)
           for f, heads in enumerate(frames)]
    assert [r.stale for r in res] == [True, True, False, False]
    assert res[0].wait() is res[0]
    with pytest.raises(pkg.RtmError, match="overwritten"):
        res[1].detections()
    got = res[3].detections()
    dev = sb.read_detections()
    for b in range(B):
        np.testing.assert_array_equal(got[b]["track_id"], dev[b]["track_id"])
        np.testing.assert_array_equal(got[b]["xyxy"], dev[b]["xyxy"])
    assert len(res[2].detections()) == B


@pytest.mark.parametrize("kalman", [False, True])
def test_state_export_import_resumes_bit_exactly(pkg, kalman):
    """rtm_state_export / rtm_state_import (checkpoint / resume): a batch resumed from a blob in
    another object continues exactly like the one that was never interrupted; a blob of another shape is refused."""
    import torch
    B, F = 3, 24
    frames = moving_heads(pkg, B, F, seed=13)
    dev = torch.device("cuda", 0)
    heads = [[torch.from_numpy(h).to(torch.bfloat16).to(dev).contiguous() for h in fr] for fr in frames]
    zones = [pkg.synth.make_zones(seed=b, num_zones=4, width=1920, height=1080, dwell_time_sec=0.2, cooldown_sec=0.4) for b in range(B)]
    make = lambda n=B, z=zones: pkg.StreamBatch(n, z, src_hw=(1080, 1920), classes=WANTED, max_tracks=128, use_kalman=kalman)
    a = make()
    for f in range(11):                                            # an odd number of steps: the tables' parity differs
        a.step(heads[f], now=3.0 + f / 30.0, frame_id=f)
    blob = a.export_state()
    b = make()
    b.import_state(blob)
    assert b.frame_id == a.frame_id == 11
    for f in range(11, F):
        for sb in (a, b):
            sb.step(heads[f], now=3.0 + f / 30.0, frame_id=f)
        ta, na = a.read_tracks()
        tb, nb = b.read_tracks()
        assert na.tolist() == nb.tolist()
        for x, y in zip(ta, tb):
            assert [(t["track_id"], t["age"], t["time_since_update"], t["xyxy"].tolist()) for t in x] == \
                   [(t["track_id"], t["age"], t["time_since_update"], t["xyxy"].tolist()) for t in y]
        ea, eb = a.read_events(), b.read_events()
        assert [[(e.track_id, e.zone_name, e.dwell_time_sec) for e in s] for s in ea] == \
               [[(e.track_id, e.zone_name, e.dwell_time_sec) for e in s] for s in eb]
    assert sum(len(t) for t in ta) > 10
    with pytest.raises(pkg.RtmError):
        make(2, zones[:2]).import_state(blob)


@pytest.mark.parametrize("case", ["f16", "rectangle_384x640", "crowded_over_1024_candidates"])
@pytest.mark.parametrize("heads_ready", [None, True])
def test_step_kernel_variants_match_the_oracle_chain(pkg, case, heads_ready):
    """The one-launch step beyond the bench shape: f16 head tensors, the stride-32 rectangle ultralytics uses for .pt
    models (1080p -> 384 x 640, 5040 anchors), and frames whose candidates outgrow the shared-memory NMS (> 1024 per
    stream: the workspace's spill arrays) - each held to the oracle chain frame by frame, in both step modes."""
    import torch
    from oracle import chain
    B, F = 3, 6
    dev = torch.device("cuda", 0)
    imgsz = (384, 640) if case.startswith("rectangle") else (640, 640)
    tdt = torch.float16 if case == "f16" else torch.bfloat16
    n_obj = 130 if case.startswith("crowded") else 14
    rng = np.random.default_rng(77)
    wh = np.stack([rng.uniform(30, 110, (B, n_obj)), rng.uniform(30, 120, (B, n_obj))], -1)
    c = np.stack([rng.uniform(60, imgsz[1] - 60, (B, n_obj)), rng.uniform(60, imgsz[0] - 60, (B, n_obj))], -1)
    v = rng.uniform(-1.5, 1.5, (B, n_obj, 2))
    cls = rng.choice(np.asarray(WANTED), (B, n_obj))
    frames = []
    for f in range(F):
        c = c + v
        levels = [[], [], []]
        for b in range(B):
            boxes = np.concatenate([c[b] - wh[b] / 2, c[b] + wh[b] / 2], 1)
            for l, t in enumerate(pkg.synth.plant_head(rng, boxes, cls[b], imgsz=imgsz, distractor_frac=0.0, logit_range=(0.0, 3.0))):
                levels[l].append(t)
        frames.append([torch.from_numpy(np.stack(l)).to(tdt) for l in levels])
    dev_frames = [[t.to(dev).contiguous() for t in fr] for fr in frames]
    zones = [pkg.synth.make_zones(seed=b, num_zones=4, width=1920, height=1080, dwell_time_sec=0.1, cooldown_sec=0.2) for b in range(B)]
    sb = pkg.StreamBatch(B, zones, src_hw=(1080, 1920), imgsz=imgsz, classes=WANTED, max_tracks=512, device=dev)
    res = chain.run_chain_parity(sb, lambda f: dev_frames[f], lambda f: [t.float() for t in frames[f]], zones, F,
                                 src_hw=(1080, 1920), classes=WANTED, heads_ready=heads_ready, imgsz=imgsz, digests=False)
    sb.close()
    assert res["ok"], res
    assert res["detections_checked"] > B * F * 8
    if case.startswith("crowded"):
        ws_cand = sum(int((torch.sigmoid(t[:, 64:].float()).amax(1) > 0.35).sum()) for t in frames[0]) / B
        assert ws_cand > 1024, ws_cand                              # the spill path was the one that ran
