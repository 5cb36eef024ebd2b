"""rtm_track_step against the reference's golden state and against the oracle (bit-exact)."""

import numpy as np
import pytest

from conftest import golden_clip, golden_state, load_golden
from oracle import tracker_ref

pytestmark = pytest.mark.gpu


def assert_state_equal(tracks, next_id, ref_state, ref_next):
    assert next_id == ref_next
    assert [t["track_id"] for t in tracks] == ref_state["track_id"].tolist()
    np.testing.assert_array_equal(np.array([t["xyxy"] for t in tracks], np.float32).reshape(-1, 4), ref_state["xyxy"])
    np.testing.assert_array_equal(np.array([t["confidence"] for t in tracks], np.float64), np.asarray(ref_state["conf"], np.float64))
    assert [t["class_id"] for t in tracks] == ref_state["cls"].tolist()
    assert [t["age"] for t in tracks] == ref_state["age"].tolist()
    assert [t["time_since_update"] for t in tracks] == ref_state["tsu"].tolist()


@pytest.mark.parametrize("name", ["cfg1_clip.npz", "crowd_clip.npz", "gaps_clip.npz", "churn_clip.npz"])
def test_facade_replays_reference_golden(pkg, name):
    g = load_golden(name)
    kw = dict(track_thresh=0.5, track_buffer=int(g.get("track_buffer", 30)), match_thresh=float(g.get("match_thresh", 0.8)))
    trk = pkg.MultiObjectTracker("bytetrack", bytetrack=dict(kw, mot20=False), max_tracks=512, max_dets=128)
    import types
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        out = trk.update(types.SimpleNamespace(xyxy=xyxy, confidence=conf, class_id=cls))
        assert out == []
        assert_state_equal(trk._core._tracks, trk._core._next_id, golden_state(g, f), int(g["next_id"][f]))


def run_batch_against_oracle(pkg, B, F, slots, clip_kw, max_tracks, seed=100, track_kw=None, check_every=1, oracle_kw=None):
    import torch
    track_kw = track_kw or {}
    xyxy, conf, cls, count = pkg.synth.scripted_batch(B, F, slots, seed=seed, **clip_kw)
    sb = pkg.StreamBatch(B, None, max_det=slots, max_tracks=max_tracks, **track_kw)
    oracles = [tracker_ref.TrackerOracle(**(track_kw if oracle_kw is None else oracle_kw)) for _ in range(B)]
    dev = sb.device
    for f in range(F):
        sb.track_only(torch.from_numpy(xyxy[f]).to(dev), torch.from_numpy(conf[f]).to(dev),
                      torch.from_numpy(cls[f]).to(dev), torch.from_numpy(count[f]).to(dev), now=f / 30.0)
        exp = [o.step(xyxy[f, b, :count[f, b]], conf[f, b, :count[f, b]], cls[f, b, :count[f, b]])
               for b, o in enumerate(oracles)]
        if f % check_every and f != F - 1:
            continue
        tracks, next_id = sb.read_tracks()
        dets = sb.read_detections() if False else None
        tid = sb.det_track_id.cpu().numpy()
        kind = sb.det_kind.cpu().numpy()
        for b, o in enumerate(oracles):
            ref = dict(track_id=o.track_id, xyxy=o.xyxy, conf=o.conf, cls=o.cls, age=o.age, tsu=o.tsu)
            assert_state_equal(tracks[b], int(next_id[b]), ref, o.next_id)
            n = count[f, b]
            np.testing.assert_array_equal(tid[b, :n], exp[b][0])
            np.testing.assert_array_equal(kind[b, :n], exp[b][1])
    return sb


def test_batched_streams_match_oracle_every_frame(pkg):
    run_batch_against_oracle(pkg, B=16, F=60, slots=64, clip_kw=dict(num_objects=30, w_range=(20, 80), h_range=(30, 120),
                                                                     vmax=4.0, dropout=0.08), max_tracks=1024)


def test_batched_equals_single_stream_runs(pkg):
    """B streams in one launch == B independent one-stream runs (tier 4)."""
    import torch
    B, F, slots = 6, 40, 32
    kw = dict(num_objects=15, w_range=(20, 60), h_range=(30, 90), vmax=4.0)
    xyxy, conf, cls, count = pkg.synth.scripted_batch(B, F, slots, seed=7, **kw)
    sb = pkg.StreamBatch(B, None, max_det=slots, max_tracks=512)
    singles = [pkg.StreamBatch(1, None, max_det=slots, max_tracks=512) for _ in range(B)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(sb.device)
    for f in range(F):
        sb.track_only(t(xyxy[f]), t(conf[f]), t(cls[f]), t(count[f]), now=0.0)
        for b, s in enumerate(singles):
            s.track_only(t(xyxy[f, b:b + 1]), t(conf[f, b:b + 1]), t(cls[f, b:b + 1]), t(count[f, b:b + 1]), now=0.0)
    tracks, next_id = sb.read_tracks()
    for b, s in enumerate(singles):
        tr, ni = s.read_tracks()
        assert ni[0] == next_id[b]
        assert [(x["track_id"], x["age"], x["time_since_update"], x["xyxy"].tolist()) for x in tr[0]] == \
               [(x["track_id"], x["age"], x["time_since_update"], x["xyxy"].tolist()) for x in tracks[b]]


def test_dense_crowd_matches_oracle(pkg):
    """Config 5 of BASELINE.json at reduced stream count: 1000 objects per stream."""
    run_batch_against_oracle(pkg, B=3, F=8, slots=1024, clip_kw=pkg.synth.dense_crowd_kwargs(1000),
                             max_tracks=4096, seed=11, check_every=4)


def test_thresholds_ties_and_degenerate_frames(pkg):
    """IoU exactly at the threshold, duplicate detections (column conflicts), zero-area boxes,
    all-low and empty frames."""
    import types
    frames = []
    box = np.array([[0, 0, 10, 10]], np.float32)
    frames.append((box, [0.9], [0]))
    frames.append((np.array([[0, 0, 10, 8]], np.float32), [0.9], [1]))        # IoU = 0.8 vs thresh 0.8
    frames.append((np.array([[0, 0, 10, 8], [0, 0, 10, 8], [50, 50, 50, 50]], np.float32), [0.9, 0.95, 0.7], [1, 2, 3]))
    frames.append((np.zeros((0, 4), np.float32), [], []))
    frames.append((np.array([[0, 0, 10, 8], [0, 0, 10, 8.5]], np.float32), [0.4, 0.45], [4, 5]))   # all low
    frames.append((np.array([[0, 0, 10, 8.5], [50, 50, 50, 50]], np.float32), [0.6, 0.3], [6, 7]))
    trk = pkg.MultiObjectTracker(max_tracks=64, max_dets=16)
    orc = tracker_ref.TrackerOracle()
    for xyxy, conf, cls in frames * 3:
        conf, cls = np.asarray(conf, np.float32), np.asarray(cls, np.int32)
        trk.update(types.SimpleNamespace(xyxy=xyxy, confidence=conf, class_id=cls))
        exp_tid, exp_kind = orc.step(xyxy, conf, cls)
        ref = dict(track_id=orc.track_id, xyxy=orc.xyxy, conf=orc.conf, cls=orc.cls, age=orc.age, tsu=orc.tsu)
        assert_state_equal(trk._core._tracks, trk._core._next_id, ref, orc.next_id)
        tid, kind = trk._core.assignments()
        np.testing.assert_array_equal(tid, exp_tid)
        np.testing.assert_array_equal(kind, exp_kind)


def test_track_table_overflow_is_reported_not_truncated(pkg):
    import types
    trk = pkg.MultiObjectTracker(max_tracks=8, max_dets=16)
    trk._core.auto_grow = False
    xy = np.arange(12, dtype=np.float32)[:, None] * 100 + np.array([0, 0, 10, 10], np.float32)
    with pytest.raises(pkg.RtmError, match="capacity"):
        trk.update(types.SimpleNamespace(xyxy=xy, confidence=np.full(12, 0.9, np.float32), class_id=np.zeros(12, np.int32)))


@pytest.mark.parametrize("mode", ["greedy", "lapjv", "kalman"])
def test_facade_tables_grow_like_the_reference_lists(pkg, mode):
    """tracker.py:55, 127-134: the reference's track list and detection arrays have no capacity.  A facade that
    starts with 8 rows / 4 detection slots ends up, frame by frame, where the oracle does (the overflowing step is
    repeated on doubled tables: ids, order and state are those of a table that was large enough all along)."""
    import types
    rng = np.random.default_rng(5)
    kw = dict(use_kalman=True) if mode == "kalman" else dict(assignment=mode)
    trk = pkg.MultiObjectTracker(max_tracks=8, max_dets=4, **kw)
    big = pkg.MultiObjectTracker(max_tracks=1024, max_dets=256, **kw)
    pos = rng.uniform(0, 1500, (150, 2)).astype(np.float32)
    for f in range(12):
        n = min(150, 3 + 14 * f)                           # 3, 17, 31, ... detections: slots and rows overflow repeatedly
        p = pos[:n] + np.float32(2.0 * f)
        xy = np.concatenate([p, p + np.float32(60)], axis=1)
        conf = np.where(np.arange(n) % 5 == 4, 0.3, 0.9).astype(np.float32)
        det = types.SimpleNamespace(xyxy=xy, confidence=conf, class_id=(np.arange(n) % 3).astype(np.int32))
        trk.update(det)
        big.update(det)
        a, b = trk._core._tracks, big._core._tracks
        assert len(a) == len(b) and trk._core._next_id == big._core._next_id
        for x, y in zip(a, b):
            assert {k: v for k, v in x.items() if k != "xyxy"} == {k: v for k, v in y.items() if k != "xyxy"}
            np.testing.assert_array_equal(x["xyxy"], y["xyxy"])
        np.testing.assert_array_equal(trk._core.assignments()[0], big._core.assignments()[0])
        if mode == "greedy":
            if f == 0:
                orc = tracker_ref.TrackerOracle()
            orc.step(xy, conf, det.class_id)
            assert [t["track_id"] for t in a] == orc.track_id.tolist() and trk._core._next_id == orc.next_id
    assert trk._core._tables[0].capacity >= 128 and trk._core.max_dets >= 150


@pytest.mark.parametrize("clip", ["slow", "gaps", "crowd"])
def test_motion_model_matches_oracle(pkg, clip):
    """Opt-in Kalman mode (row K; the reference has none): ids, assignments and stored state
    bit-exact against the oracle, filter means / covariances within the stated 1e-4 relative
    (in fact bit for bit: both sides spell out the same float32 operations)."""
    kw = dict(slow=dict(num_objects=20, w_range=(60, 160), h_range=(120, 320), vmax=2.0, dropout=0.05),
              gaps=dict(num_objects=12, w_range=(80, 160), h_range=(120, 300), vmax=3.0, dropout=0.35),
              crowd=dict(num_objects=200, w_range=(20, 60), h_range=(50, 150), vmax=1.5, dropout=0.05))[clip]
    slots = 256 if clip == "crowd" else 64
    sb = run_batch_against_oracle(pkg, B=5, F=50, slots=slots, clip_kw=kw, max_tracks=1024, seed=21,
                                  track_kw=dict(use_kalman=True))
    # replay the oracle once more to compare the filter state itself
    xyxy, conf, cls, count = pkg.synth.scripted_batch(5, 50, slots, seed=21, **kw)
    host = sb.table.to_host()
    matched_more_than_once = 0
    for b in range(5):
        o = tracker_ref.TrackerOracle(use_kalman=True)
        for f in range(50):
            o.step(xyxy[f, b, :count[f, b]], conf[f, b, :count[f, b]], cls[f, b, :count[f, b]])
        n = len(o)
        assert host["count"][b] == n
        np.testing.assert_allclose(host["kf_mean"][b, :n], o.kf_mean, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(host["kf_cov"][b, :n], o.kf_cov, rtol=1e-4, atol=1e-9)
        np.testing.assert_array_equal(host["kf_mean"][b, :n], o.kf_mean)
        np.testing.assert_array_equal(host["kf_cov"][b, :n], o.kf_cov)
        matched_more_than_once += int((o.age > 2).sum())
    assert matched_more_than_once > 0                      # the update path ran, not only initiate / predict


def test_motion_model_off_is_the_reference(pkg):
    """use_kalman=False and rtm_track_step_ex with NULL Kalman pointers are rtm_track_step."""
    run_batch_against_oracle(pkg, B=4, F=30, slots=64, clip_kw=dict(num_objects=25, w_range=(30, 90), h_range=(40, 140),
                                                                    vmax=3.0, dropout=0.1), max_tracks=512, seed=5,
                             track_kw=dict(use_kalman=False))


@pytest.mark.parametrize("clip", ["sparse", "crowd", "dense"])
def test_optimal_assignment_matches_lapjv_emulation(pkg, clip):
    """assignment="lapjv" (tracker.py:168-181, the branch taken when `lap` is installed) against the
    scipy emulation of lap.lapjv(extend_cost, cost_limit): bit-exact state and assignments.  The clips
    are dense enough for the optimal and the greedy assignment to differ."""
    kw = dict(sparse=dict(num_objects=20, w_range=(60, 160), h_range=(120, 320), vmax=2.0, dropout=0.05),
              crowd=dict(num_objects=250, w_range=(20, 60), h_range=(50, 150), vmax=1.5, dropout=0.05),
              dense=pkg.synth.dense_crowd_kwargs(1000))[clip]
    B, F, slots = (4, 40, 64) if clip == "sparse" else ((3, 20, 256) if clip == "crowd" else (2, 6, 1024))
    run_batch_against_oracle(pkg, B=B, F=F, slots=slots, clip_kw=kw, max_tracks=4096, seed=31,
                             track_kw=dict(assignment="lapjv"), oracle_kw=dict(assign=tracker_ref.assign_lapjv_emulated))
    if clip == "crowd":
        # the same clip under the greedy rule ends elsewhere: the test above is not vacuous
        xyxy, conf, cls, count = pkg.synth.scripted_batch(B, F, slots, seed=31, **kw)
        a, b = tracker_ref.TrackerOracle(), tracker_ref.TrackerOracle(assign=tracker_ref.assign_lapjv_emulated)
        for f in range(F):
            a.step(xyxy[f, 0, :count[f, 0]], conf[f, 0, :count[f, 0]], cls[f, 0, :count[f, 0]])
            b.step(xyxy[f, 0, :count[f, 0]], conf[f, 0, :count[f, 0]], cls[f, 0, :count[f, 0]])
        assert a.next_id != b.next_id or not np.array_equal(a.track_id, b.track_id) or not np.array_equal(a.xyxy, b.xyxy)


def test_optimal_assignment_resolves_conflicts_the_greedy_rule_cannot(pkg):
    """Two tracks whose best detection is the same one: greedy gives it to the first track and leaves
    the second unmatched (no second choice); the optimal assignment matches both."""
    import types
    t0 = np.array([[0, 0, 100, 100], [4, 0, 104, 100]], np.float32)          # tracks A, B
    d1 = np.array([[3, 0, 103, 100], [9, 0, 109, 100]], np.float32)          # X best for both; Y admissible for B only
    conf, cls = np.full(2, 0.9, np.float32), np.zeros(2, np.int32)
    for mode, assign in (("greedy", tracker_ref.assign_rowloop), ("lapjv", tracker_ref.assign_lapjv_emulated)):
        trk = pkg.MultiObjectTracker(max_tracks=16, max_dets=8, assignment=mode)
        orc = tracker_ref.TrackerOracle(assign=assign)
        for boxes in (t0, d1):
            trk.update(types.SimpleNamespace(xyxy=boxes, confidence=conf, class_id=cls))
            exp_tid, exp_kind = orc.step(boxes, conf, cls)
            tid, kind = trk._core.assignments()
            np.testing.assert_array_equal(tid, exp_tid)
            np.testing.assert_array_equal(kind, exp_kind)
        assert trk._core._next_id == orc.next_id == (4 if mode == "greedy" else 3)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("assignment,kalman", [("greedy", False), ("greedy", True), ("lapjv", False), ("lapjv", True)])
def test_tracker_modes_soak(pkg, seed, assignment, kalman):
    """Seeded clips with randomly drawn density, speed and dropout through every combination of the
    assignment rule and the motion model: state, ids and per-detection assignments bit-exact."""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.choice([5, 40, 150]))
    kw = dict(num_objects=n, w_range=(20, float(rng.uniform(40, 160))), h_range=(30, float(rng.uniform(60, 300))),
              vmax=float(rng.uniform(0.5, 6.0)), dropout=float(rng.uniform(0.0, 0.4)))
    assign = tracker_ref.assign_lapjv_emulated if assignment == "lapjv" else tracker_ref.assign_rowloop
    run_batch_against_oracle(pkg, B=3, F=25, slots=256, clip_kw=kw, max_tracks=2048, seed=700 + seed,
                             track_kw=dict(assignment=assignment, use_kalman=kalman),
                             oracle_kw=dict(assign=assign, use_kalman=kalman))


def test_active_tracks_is_a_view_not_a_step(pkg):
    """Reading the active view twice in a frame leaves the trails alone (one point per update, tracker.py:241-258)."""
    import types
    trk = pkg.MultiObjectTracker()
    for f in range(4):
        box = np.array([[10 + f, 10, 60 + f, 80]], np.float32)
        trk.update(types.SimpleNamespace(xyxy=box, confidence=np.array([0.9], np.float32), class_id=np.array([2], np.int32)))
        a = trk.active_tracks()
        b = trk.active_tracks()
        assert len(a) == len(b) == 1 and a[0].trail == b[0].trail and len(a[0].trail) == f + 1


@pytest.mark.parametrize("case", ["thresh_0.3", "thresh_0.05", "no_scratch"])
def test_optimal_assignment_has_no_size_limits(pkg, case):
    """lap.lapjv solves a problem of any size (tracker.py:168-181).  In a 1000-object crowd at match_thresh = 0.05 a
    stage holds ~3,500 admissible pairs (more in later frames) in conflict components of up to ~750 rows (at 0.3: 1,200
    pairs, components of 23): beyond the shared-memory solver (4096 pairs, 32 x 32 components) the general solver in
    global scratch takes over - same state and assignments as the scipy emulation, no status bit.  Without scratch the same
    clip reports RTM_STATUS_ASSIGN_LIMIT instead of guessing."""
    import torch
    kw = pkg.synth.dense_crowd_kwargs(1000)
    if case != "no_scratch":
        thr = float(case.split("_")[1])
        sb = run_batch_against_oracle(pkg, B=2, F=5, slots=1024, clip_kw=kw, max_tracks=4096, seed=77,
                                      track_kw=dict(assignment="lapjv", match_thresh=thr, max_pairs=200_000),
                                      oracle_kw=dict(assign=tracker_ref.assign_lapjv_emulated, match_thresh=thr))
        assert not sb.status.cpu().numpy().any()
        return
    xyxy, conf, cls, count = pkg.synth.scripted_batch(1, 3, 1024, seed=77, **kw)
    sb = pkg.StreamBatch(1, None, max_det=1024, max_tracks=4096, max_pairs=0, assignment="lapjv", match_thresh=0.05)
    sb.assign_scratch = None                                        # a caller of the C ABI that passes no scratch
    sb._io_cache.clear()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(sb.device)
    with pytest.raises(pkg.RtmError):
        for f in range(3):
            sb.track_only(t(xyxy[f]), t(conf[f]), t(cls[f]), t(count[f]), now=0.0)
            sb.check_status()
