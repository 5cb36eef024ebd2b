"""rtm_zone_step against the reference's golden events, the oracle and cv2 (bit-exact)."""

import ctypes as C
import types

import numpy as np
import pytest

from conftest import golden_clip, load_golden
from oracle import tracker_ref, zone_ref
from test_oracle_golden import tracker_params, zones_for

pytestmark = pytest.mark.gpu


def event_rows(events, zone_index):
    return [(e.frame_id, e.track_id, zone_index[(e.zone_name, e.event_type)], e.class_id, *e.centroid,
             e.dwell_time_sec, *e.bbox_xyxy) for e in events]


@pytest.mark.parametrize("name", ["cfg1_clip.npz", "crowd_clip.npz", "gaps_clip.npz", "churn_clip.npz"])
def test_fused_tracker_and_zones_replay_reference_events(pkg, name):
    """Scripted detections -> rtm_track_step -> rtm_zone_step == reference tracker + zone engine."""
    import torch
    g = load_golden(name)
    zones = zones_for(name)
    zone_index = {(z["name"], z.get("trigger", "intrusion")): i for i, z in reversed(list(enumerate(zones)))}
    p = tracker_params(g)
    slots = 128
    sb = pkg.StreamBatch(1, [zones], max_det=slots, max_tracks=512, src_hw=(720, 1280), **p)
    rows = []
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        n = len(conf)
        bx = np.zeros((1, slots, 4), np.float32); bx[0, :n] = xyxy
        cf = np.zeros((1, slots), np.float32); cf[0, :n] = conf
        cl = np.zeros((1, slots), np.int32); cl[0, :n] = cls
        t = lambda a: torch.from_numpy(a).to(sb.device)
        sb.track_only(t(bx), t(cf), t(cl), t(np.array([n], np.int32)), now=float(g["t0"]) + f / float(g["fps"]), frame_id=f)
        rows += event_rows(sb.read_events()[0], zone_index)
    np.testing.assert_array_equal(np.array(rows, np.float64).reshape(-1, 11), g["events"])


def test_point_in_polygon_kernel_matches_cv2_golden(pkg):
    """One stream per golden polygon, one track per query point; dwell = cooldown = 0 makes every
    inside test visible as an event."""
    import torch
    g = load_golden("pip_cases.npz")
    off = g["poly_offsets"]
    P, Q = len(off) - 1, g["points"].shape[1]
    zones = [[dict(name="z", polygon=g["poly_xy"][off[p]:off[p + 1]].tolist(), dwell_time_sec=0.0, cooldown_sec=0.0)]
             for p in range(P)]
    dev = torch.device("cuda:0")
    zt = pkg.ZoneTables(zones, Q, dev, max_events=Q)
    tt = pkg.DeviceTrackTable(P, Q, dev)
    pts = g["points"].astype(np.float32)
    box = np.concatenate([pts, pts], 2)                     # centroid == the point
    tt.xyxy.copy_(torch.from_numpy(box))
    tt.track_id.copy_(torch.arange(1, Q + 1, dtype=torch.int32).repeat(P, 1))
    tt.time_since_update.fill_(1)
    tt.count.fill_(Q)
    status = torch.zeros(P, dtype=torch.int32, device=dev)
    lib = pkg._lib.lib()
    st = zt.state_in()[2]
    pkg._lib.check(lib.rtm_zone_step(C.byref(zt.zone_set), C.byref(tt.struct), None, C.byref(st), C.byref(st), 5.0, None, 0,
                                     zt.events.data_ptr(), zt.event_stride, zt.event_count.data_ptr(), status.data_ptr(),
                                     pkg._lib.cuda_stream()))
    assert not status.cpu().numpy().any()
    cnt = zt.event_count.cpu().numpy()
    recs = zt.events.cpu().numpy().view(np.dtype(pkg._lib.EVENT_DTYPE)).reshape(P, -1)
    for p in range(P):
        inside = np.zeros(Q, bool)
        inside[recs[p, :cnt[p]]["row"]] = True
        np.testing.assert_array_equal(inside, g["result"][p] >= 0, err_msg=f"polygon {p}")
        np.testing.assert_array_equal(recs[p, :cnt[p]]["row"], np.flatnonzero(inside))       # row order


def test_facade_matches_oracle_with_arbitrary_track_lists(pkg, tmp_path):
    """Ids vanish and come back, call order differs from first-seen order, duplicate zone names
    share state, default dwell / cooldown, negative coordinates, table growth."""
    rng = np.random.default_rng(42)
    zones = pkg.synth.make_zones(seed=1, num_zones=7, width=640, height=480, kmin=3, kmax=8, dwell_time_sec=0.2, cooldown_sec=0.5)
    zones[4]["name"] = zones[1]["name"]
    zones[6] = dict(name="defaults", polygon=[[0, 0], [640, 0], [640, 480], [0, 480]])
    zones[5] = dict(name="left_half", polygon=[[0, 0], [320, 0], [320, 480], [0, 480]], dwell_time_sec=0.2, cooldown_sec=0.4)
    clock = {"t": 100.0}
    eng = pkg.ZoneEventEngine(zones, log_path=str(tmp_path / "ev.jsonl"), clock=lambda: clock["t"], initial_rows=8)
    orc = zone_ref.ZoneOracle(zones)
    pos = rng.uniform(-20, 660, (40, 2))
    total = 0
    for f in range(200):
        clock["t"] = 100.0 + f * 0.11
        pos += rng.uniform(-4, 4, pos.shape)
        ids = rng.permutation(40)[: int(rng.integers(0, 41) if f % 17 else 0)]
        tracks = []
        for i in ids:
            b = np.array([pos[i, 0] - 5, pos[i, 1] - 9, pos[i, 0] + 5, pos[i, 1] + 9], np.float32)
            tracks.append(types.SimpleNamespace(track_id=int(i) + 1, xyxy=b, class_id=int(i) % 5, class_name=f"c{i % 5}"))
        got = eng.process(tracks, f)
        exp = orc.process([(t.track_id, t.xyxy, t.class_id) for t in tracks], f, clock["t"])
        assert [(e.track_id, e.zone_name, e.event_type, e.class_id, e.centroid, e.dwell_time_sec, e.bbox_xyxy, e.frame_id, e.class_name)
                for e in got] == \
               [(e.track_id, e.zone_name, e.event_type, e.class_id, e.centroid, e.dwell_time_sec, e.bbox_xyxy, e.frame_id, f"c{(e.track_id - 1) % 5}")
                for e in exp]
        total += len(got)
    assert total > 20
    lines = open(tmp_path / "ev.jsonl").read().splitlines()
    assert len(lines) == total
    import json
    rec = json.loads(lines[0])
    assert set(rec) == {"timestamp_utc", "event_type", "zone_name", "track_id", "class_id", "class_name",
                        "dwell_time_sec", "bbox_xyxy", "centroid", "frame_id", "metadata"}
    assert [n for n, _ in eng.get_zone_polygons()] == [z["name"] for z in zones]


def test_reference_dwell_cooldown_schedule(pkg, tmp_path):
    """SURVEY.md section 8a Z2 probe: dwell 2 s / cooldown 10 s at 30 fps fires at frames 60, 360, 660."""
    clock = {"t": 0.0}
    eng = pkg.ZoneEventEngine([dict(name="a", polygon=[[0, 0], [100, 0], [100, 100], [0, 100]])],
                              log_path=str(tmp_path / "e.jsonl"), clock=lambda: clock["t"])
    trk = [types.SimpleNamespace(track_id=1, xyxy=np.array([40, 40, 60, 60], np.float32), class_id=0)]
    fired = []
    for f in range(700):
        clock["t"] = 1_700_000_000.0 + f / 30.0
        for e in eng.process(trk, f):
            fired.append((f, e.dwell_time_sec))
    assert fired == [(60, 2.0), (360, 12.0), (660, 22.0)]


def strip_stamp(line: str) -> str:
    import re
    return re.sub(r'"timestamp_utc": "[^"]*"', '"timestamp_utc": ""', line)


@pytest.mark.parametrize("name", ["cfg1_clip.npz", "crowd_clip.npz", "churn_clip.npz"])
def test_event_sink_lines_are_the_references_bytes(pkg, name, tmp_path):
    """The event sink (zone_engine.py:44-45, 153-157): StreamBatch.write_events appends, once per step, the lines
    the UNMODIFIED reference wrote to its own events.jsonl for the same clip (captured by make_goldens.py) -
    byte for byte, apart from `timestamp_utc`, which is the wall clock of the run that wrote them."""
    import torch
    g = load_golden(name)
    zones = zones_for(name)
    want = bytes(g["jsonl"]).decode().splitlines()
    assert len(want) == len(g["events"]) > 0
    slots = 128
    sb = pkg.StreamBatch(1, [zones], max_det=slots, max_tracks=512, src_hw=(720, 1280), **tracker_params(g))
    log = tmp_path / "logs" / "events.jsonl"
    written = 0
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        n = len(conf)
        bx = np.zeros((1, slots, 4), np.float32); bx[0, :n] = xyxy
        cf = np.zeros((1, slots), np.float32); cf[0, :n] = conf
        cl = np.zeros((1, slots), np.int32); cl[0, :n] = cls
        t = lambda a: torch.from_numpy(a).to(sb.device)
        sb.track_only(t(bx), t(cf), t(cl), t(np.array([n], np.int32)), now=float(g["t0"]) + f / float(g["fps"]), frame_id=f)
        written += sb.write_events(log)
    got = open(log).read().splitlines()
    assert written == len(got) == len(want)
    assert [strip_stamp(l) for l in got] == [strip_stamp(l) for l in want]


def test_facade_event_log_is_the_references_bytes(pkg, tmp_path):
    """The same for the single-stream drop-in: ZoneEventEngine.process fed the reference's own active-track
    lists writes the reference's lines."""
    g = load_golden("cfg1_clip.npz")
    zones = zones_for("cfg1_clip.npz")
    want = bytes(g["jsonl"]).decode().splitlines()
    clock = {"t": 0.0}
    eng = pkg.ZoneEventEngine(zones, log_path=str(tmp_path / "events.jsonl"), clock=lambda: clock["t"])
    from conftest import golden_state
    for f in range(len(g["next_id"])):
        st = golden_state(g, f)
        act = np.flatnonzero(st["tsu"] == 1)
        clock["t"] = float(g["t0"]) + f / float(g["fps"])
        eng.process([types.SimpleNamespace(track_id=int(st["track_id"][r]), xyxy=st["xyxy"][r], class_id=int(st["cls"][r])) for r in act], f)
    got = open(tmp_path / "events.jsonl").read().splitlines()
    assert [strip_stamp(l) for l in got] == [strip_stamp(l) for l in want]


def test_facade_does_work_proportional_to_the_call_and_handles_repeated_ids(pkg, tmp_path):
    """Thousands of ids come and go: the device table stays as small as the largest call (the reference keeps dicts;
    so does the facade).  An id passed twice in one call is evaluated one after the other, as the reference's loop does."""
    rng = np.random.default_rng(5)
    zones = pkg.synth.make_zones(seed=2, num_zones=5, width=640, height=480, kmin=3, kmax=8, dwell_time_sec=0.1, cooldown_sec=0.3)
    clock = {"t": 10.0}
    eng = pkg.ZoneEventEngine(zones, log_path=str(tmp_path / "ev.jsonl"), clock=lambda: clock["t"], initial_rows=8)
    orc = zone_ref.ZoneOracle(zones)
    total = 0
    for f in range(150):
        clock["t"] = 10.0 + f * 0.05
        base = 1 + 20 * (f // 10)                                   # the id range moves on: 300+ ids over the run
        ids = base + rng.permutation(30)[: int(rng.integers(1, 25))]
        if f % 7 == 3:
            ids = np.concatenate([ids, ids[:2]])                    # the same ids again, later in the same call
        tracks = []
        for k, i in enumerate(ids):
            c = np.array([(37 * int(i)) % 640 + 0.4 * f, (53 * int(i)) % 480 + 0.2 * f], np.float32)   # slow drift: they dwell
            tracks.append(types.SimpleNamespace(track_id=int(i), xyxy=np.array([c[0] - 6, c[1] - 8, c[0] + 6, c[1] + 8], np.float32),
                                                class_id=int(i) % 3))
        got = eng.process(tracks, f)
        exp = orc.process([(t.track_id, t.xyxy, t.class_id) for t in tracks], f, clock["t"])
        assert [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec, e.frame_id) for e in got] == \
               [(e.track_id, e.zone_name, e.centroid, e.dwell_time_sec, e.frame_id) for e in exp]
        total += len(got)
    assert total > 20
    assert eng._tables.capacity <= 32 and len(eng._cooldown) > 10     # state on the host, a small table on the device
