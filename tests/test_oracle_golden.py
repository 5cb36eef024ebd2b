"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_goldens.py)."""

import numpy as np
import pytest

from conftest import golden_clip, golden_state, load_golden
from oracle import tracker_ref, zone_ref

CLIPS = ["cfg1_clip.npz", "crowd_clip.npz", "gaps_clip.npz", "churn_clip.npz"]


def zones_for(name):
    import importlib
    synth = importlib.import_module("rtmodt_b200").synth
    if name == "cfg1_clip.npz":
        return synth.make_zones(seed=0, num_zones=4)
    if name == "crowd_clip.npz":
        z = synth.make_zones(seed=3, num_zones=16, kmin=4, kmax=12, dwell_time_sec=0.2, cooldown_sec=0.5)
        z[5]["name"] = z[2]["name"]
        z[5]["trigger"] = "crossing"
        z[7].pop("dwell_time_sec")
        return z
    if name == "gaps_clip.npz":
        return synth.make_zones(seed=5, num_zones=3)
    return synth.make_zones(seed=9, num_zones=6, kmin=3, kmax=9, dwell_time_sec=0.1, cooldown_sec=0.3)


def tracker_params(g):
    return dict(track_buffer=int(g.get("track_buffer", 30)), match_thresh=float(g.get("match_thresh", 0.8)))


@pytest.mark.parametrize("name", CLIPS)
@pytest.mark.parametrize("assign", [tracker_ref.assign_rowloop, tracker_ref.assign_columnwin])
def test_tracker_oracle_matches_reference_state(name, assign):
    g = load_golden(name)
    trk = tracker_ref.TrackerOracle(assign=assign, **tracker_params(g))
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        trk.step(xyxy, conf, cls)
        ref = golden_state(g, f)
        assert trk.next_id == g["next_id"][f]
        np.testing.assert_array_equal(trk.track_id, ref["track_id"])
        np.testing.assert_array_equal(trk.xyxy, ref["xyxy"])           # bit-exact f32
        np.testing.assert_array_equal(trk.conf.astype(np.float64), ref["conf"])
        np.testing.assert_array_equal(trk.cls, ref["cls"])
        np.testing.assert_array_equal(trk.age, ref["age"])
        np.testing.assert_array_equal(trk.tsu, ref["tsu"])
    assert g["returned"].sum() == 0          # the reference returns [] every frame


@pytest.mark.parametrize("name", CLIPS)
def test_zone_oracle_matches_reference_events(name):
    g = load_golden(name)
    trk = tracker_ref.TrackerOracle(**tracker_params(g))
    eng = zone_ref.ZoneOracle(zones_for(name))
    rows = []
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        trk.step(xyxy, conf, cls)
        act = trk.active_rows()
        now = float(g["t0"]) + f / float(g["fps"])
        for e in eng.process(zip(trk.track_id[act], trk.xyxy[act], trk.cls[act]), f, now):
            rows.append((f, e.track_id, e.zone_index, e.class_id, *e.centroid, e.dwell_time_sec, *e.bbox_xyxy))
    got = np.array(rows, np.float64).reshape(-1, 11)
    np.testing.assert_array_equal(got, g["events"])


def test_point_in_polygon_matches_cv2_golden():
    g = load_golden("pip_cases.npz")
    off = g["poly_offsets"]
    for p in range(len(off) - 1):
        poly = g["poly_xy"][off[p]:off[p + 1]]
        got = [zone_ref.point_in_polygon(poly, int(x), int(y)) for x, y in g["points"][p]]
        np.testing.assert_array_equal(np.array(got, np.int8), g["result"][p])


def test_point_in_polygon_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(123)
    for _ in range(300):
        k = int(rng.integers(3, 10))
        poly = rng.integers(0, 30, (k, 2)).astype(np.int32)
        for x, y in rng.integers(-1, 31, (40, 2)):
            assert zone_ref.point_in_polygon(poly, int(x), int(y)) == int(
                cv2.pointPolygonTest(poly, (int(x), int(y)), False))


def test_assign_forms_agree_on_ties_and_conflicts():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        t, n = int(rng.integers(1, 9)), int(rng.integers(1, 9))   # the reference never assigns on an empty side
        iou = rng.choice(np.array([0.0, 0.5, 0.8, 0.85, 0.9], np.float32), (t, n))
        assert tracker_ref.assign_rowloop(iou, 0.8) == tracker_ref.assign_columnwin(iou, 0.8)


def test_centroid_truncates_toward_zero():
    assert zone_ref.centroid(np.array([-5.5, -3.0, 2.2, 0.5], np.float32)) == (-1, -1)
    assert zone_ref.centroid(np.array([10.0, 20.0, 15.0, 25.0], np.float32)) == (12, 22)
