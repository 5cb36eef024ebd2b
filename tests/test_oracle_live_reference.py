"""The oracle against the UNMODIFIED reference run live, on randomly parameterised clips.

The golden files pin the oracle on four fixed clips; here the reference's own ``tracker.py`` and
``zone_engine.py`` are imported from /root/reference (dev container only - the GPU box has no reference,
the test skips itself there) and driven through ``oracle/make_goldens.run_reference_clip`` on clips whose
object count, box sizes, speeds, dropout, thresholds, buffer length, zone count, dwell and cooldown are
drawn per seed, with empty and all-low-score frames mixed in.  The oracle has to reproduce the
reference's track tables after every frame and its event list, bit for bit.
"""

import importlib
import os

import numpy as np
import pytest

from conftest import golden_clip, golden_state
from oracle import tracker_ref, zone_ref

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "tracking")),
                                reason="the reference lives in the dev container only")


@pytest.fixture(scope="module")
def ref():
    mg = importlib.import_module("oracle.make_goldens")
    return mg, mg.load_reference()


def random_case(seed):
    synth = importlib.import_module("rtmodt_b200").synth
    rng = np.random.default_rng(10_000 + seed)
    n_obj = int(rng.choice([3, 12, 40, 90]))
    small = bool(rng.integers(0, 2))
    clip = synth.scripted_clip(seed=100 + seed, num_frames=int(rng.integers(40, 90)), num_objects=n_obj,
                               w_range=(18, 60) if small else (50, 160), h_range=(30, 100) if small else (90, 300),
                               vmax=float(rng.uniform(0.5, 5.0)), dropout=float(rng.uniform(0.0, 0.2)))
    for f in rng.choice(len(clip), size=int(rng.integers(0, 8)), replace=False):       # empty frames (tracker.py:70-73)
        clip[f] = (np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32))
    for f in rng.choice(len(clip), size=int(rng.integers(0, 5)), replace=False):       # all-low frames: stage 2 only
        b, c, k = clip[f]
        clip[f] = (b, np.minimum(c, np.float32(0.45)), k)
    params = dict(track_thresh=0.5, track_buffer=int(rng.choice([3, 12, 30])),
                  match_thresh=float(rng.choice([0.5, 0.7, 0.8, 0.9])))
    zones = synth.make_zones(seed=200 + seed, num_zones=int(rng.integers(1, 10)), kmin=3, kmax=10,
                             dwell_time_sec=float(rng.choice([0.0, 0.1, 0.5])), cooldown_sec=float(rng.choice([0.0, 0.3, 2.0])))
    if len(zones) > 3 and seed % 3 == 0:
        zones[3]["name"] = zones[1]["name"]          # duplicate names share state (zone_engine.py keys by name)
        zones[3]["trigger"] = "crossing"
    return clip, params, zones


@pytest.mark.parametrize("seed", range(12))
def test_oracle_reproduces_the_live_reference(ref, seed):
    mg, (ref_tracker, ref_zones) = ref
    clip, params, zones = random_case(seed)
    g = mg.run_reference_clip(ref_tracker, ref_zones, clip, zones,
                              tracker_kwargs=dict(bytetrack=dict(mot20=False, **params)))
    trk = tracker_ref.TrackerOracle(track_buffer=params["track_buffer"], match_thresh=params["match_thresh"])
    eng = zone_ref.ZoneOracle(zones)
    rows = []
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        trk.step(xyxy, conf, cls)
        want = golden_state(g, f)
        assert trk.next_id == g["next_id"][f]
        np.testing.assert_array_equal(trk.track_id, want["track_id"])
        np.testing.assert_array_equal(trk.xyxy, want["xyxy"])
        np.testing.assert_array_equal(trk.conf.astype(np.float64), want["conf"])
        np.testing.assert_array_equal(trk.cls, want["cls"])
        np.testing.assert_array_equal(trk.age, want["age"])
        np.testing.assert_array_equal(trk.tsu, want["tsu"])
        act = trk.active_rows()
        now = float(g["t0"]) + f / float(g["fps"])
        for e in eng.process(zip(trk.track_id[act], trk.xyxy[act], trk.cls[act]), f, now):
            rows.append((f, e.track_id, e.zone_index, e.class_id, *e.centroid, e.dwell_time_sec, *e.bbox_xyxy))
    np.testing.assert_array_equal(np.array(rows, np.float64).reshape(-1, 11), g["events"])
    assert g["returned"].sum() == 0


def test_oracle_reproduces_the_live_reference_on_a_dense_crowd(ref):
    """BASELINE.json configs[4] scale: 1000 objects per frame (MOT20-like), 16 zones - the reference's own
    `_ByteTrackCore.update` at T = N = 1000 (tracker.py:58-141) and its zone engine, live, against the oracle."""
    mg, (ref_tracker, ref_zones) = ref
    synth = importlib.import_module("rtmodt_b200").synth
    clip = synth.scripted_clip(seed=900, num_frames=12, **synth.dense_crowd_kwargs(1000))
    zones = synth.make_zones(seed=0, num_zones=16, width=1920, height=1080, kmin=4, kmax=12, dwell_time_sec=0.1, cooldown_sec=0.2)
    g = mg.run_reference_clip(ref_tracker, ref_zones, clip, zones, tracker_kwargs=dict(bytetrack=dict(mot20=False)))
    trk, eng = tracker_ref.TrackerOracle(), zone_ref.ZoneOracle(zones)
    rows = []
    for f, (xyxy, conf, cls) in enumerate(golden_clip(g)):
        trk.step(xyxy, conf, cls)
        want = golden_state(g, f)
        assert trk.next_id == g["next_id"][f] and len(trk) >= 700
        np.testing.assert_array_equal(trk.track_id, want["track_id"])
        np.testing.assert_array_equal(trk.xyxy, want["xyxy"])
        np.testing.assert_array_equal(trk.age, want["age"])
        np.testing.assert_array_equal(trk.tsu, want["tsu"])
        act = trk.active_rows()
        now = float(g["t0"]) + f / float(g["fps"])
        for e in eng.process(zip(trk.track_id[act], trk.xyxy[act], trk.cls[act]), f, now):
            rows.append((f, e.track_id, e.zone_index, e.class_id, *e.centroid, e.dwell_time_sec, *e.bbox_xyxy))
    assert len(rows) > 100
    np.testing.assert_array_equal(np.array(rows, np.float64).reshape(-1, 11), g["events"])
