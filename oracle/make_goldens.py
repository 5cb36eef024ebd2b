"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in the dev container.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Run from the repo root:

    python oracle/make_goldens.py            # needs /root/reference (dev container only)

What is executed is the reference's own code, imported from where it lies:
``/root/reference/src/tracking/tracker.py`` (``MultiObjectTracker`` -> ``_ByteTrackCore``,
greedy branch because ``lap`` is absent) and ``/root/reference/src/events/zone_engine.py``
(``ZoneEventEngine``, with ``zone_engine.time.time`` patched to a scripted 30 fps clock and
the log file in a temp dir).  Inputs come from ``synth.py`` with fixed seeds and are stored
in the golden files next to the reference's outputs, so the files are self-contained: the
GPU box, where /root/reference does not exist, replays them.

Because ``MultiObjectTracker.update`` always returns ``[]`` the frozen
surface is the internal state ``tracker._core._tracks`` / ``_next_id`` after every frame;
the zone engine is fed the "matched or born this frame" view (``time_since_update == 1``).
"""

from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")

synth = importlib.import_module("real-time-multi-object-detection---tracking-system_b200.synth")


def load_reference():
    """Import the reference's tracker / zone modules from /root/reference, quietly."""
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_goldens.py needs the reference at /root/reference (dev container)")
    sys.path.insert(0, REFERENCE)
    from loguru import logger
    logger.remove()
    tracker = importlib.import_module("src.tracking.tracker")
    zones = importlib.import_module("src.events.zone_engine")
    return tracker, zones


def pack_states(states):
    """Ragged per-frame track tables -> flat arrays + offsets."""
    off = np.zeros(len(states) + 1, np.int64)
    for f, s in enumerate(states):
        off[f + 1] = off[f] + len(s)
    cat = lambda key, dt: np.array([t[key] for s in states for t in s], dt)
    xy = np.array([t["xyxy"] for s in states for t in s], np.float32).reshape(-1, 4)
    return dict(state_offsets=off, state_track_id=cat("track_id", np.int32), state_xyxy=xy,
                state_conf=cat("confidence", np.float64), state_cls=cat("class_id", np.int32),
                state_age=cat("age", np.int32), state_tsu=cat("time_since_update", np.int32))


def pack_clip(clip):
    off = np.zeros(len(clip) + 1, np.int64)
    for f, (_, c, _) in enumerate(clip):
        off[f + 1] = off[f] + len(c)
    return dict(det_offsets=off,
                det_xyxy=np.concatenate([b for b, _, _ in clip]).astype(np.float32).reshape(-1, 4),
                det_conf=np.concatenate([c for _, c, _ in clip]).astype(np.float32),
                det_cls=np.concatenate([k for _, _, k in clip]).astype(np.int32))


def run_reference_clip(ref_tracker, ref_zones, clip, zone_cfgs, fps=30.0, tracker_kwargs=None):
    """One stream through reference tracker + zone engine; returns golden arrays."""
    trk = ref_tracker.MultiObjectTracker("bytetrack", **(tracker_kwargs or {}))
    tmp = tempfile.mkdtemp(prefix="rtm_golden_")
    eng = ref_zones.ZoneEventEngine(zone_cfgs, log_path=os.path.join(tmp, "events.jsonl"))
    states, next_ids, returned = [], [], []
    ev_rows = []
    clock = {"t": 0.0}
    ref_zones.time.time = lambda: clock["t"]              # zone_engine.py:84 reads this
    for f, (xyxy, conf, cls) in enumerate(clip):
        dets = types.SimpleNamespace(xyxy=xyxy, confidence=conf, class_id=cls)
        out = trk.update(dets)
        returned.append(len(out))
        states.append([dict(t, xyxy=np.array(t["xyxy"], np.float32)) for t in trk._core._tracks])
        next_ids.append(trk._core._next_id)
        active = [types.SimpleNamespace(track_id=t["track_id"], xyxy=t["xyxy"], class_id=t["class_id"])
                  for t in trk._core._tracks if t["time_since_update"] == 1]
        clock["t"] = 1_700_000_000.0 + f / fps
        for e in eng.process(active, f):
            # zone index: first zone whose name+trigger match and that contains the centroid
            zi = next(i for i, z in enumerate(eng.zones)
                      if z.name == e.zone_name and z.trigger == e.event_type)
            ev_rows.append((f, e.track_id, zi, e.class_id, e.centroid[0], e.centroid[1],
                            e.dwell_time_sec, *e.bbox_xyxy))
    ev = np.array(ev_rows, np.float64).reshape(-1, 11)
    # the event sink's wire format: the very lines ZoneEventEngine._write appended (zone_engine.py:153-157);
    # `timestamp_utc` in them is the wall clock of this run (time.gmtime), everything else is deterministic
    log = os.path.join(tmp, "events.jsonl")
    jsonl = open(log, "rb").read() if os.path.exists(log) else b""
    out = pack_states(states)
    out["jsonl"] = np.frombuffer(jsonl, np.uint8).copy()
    out.update(pack_clip(clip))
    out.update(next_id=np.array(next_ids, np.int64), returned=np.array(returned, np.int64),
               events=ev, fps=np.float64(fps), t0=np.float64(1_700_000_000.0))
    return out


def golden_pip(ref_zones, n_poly=400, n_pts=60, seed=7):
    """``cv2.pointPolygonTest`` exactly as zone_engine.py:94 calls it, on random and
    degenerate int32 polygons with points biased onto vertices / edges."""
    import cv2
    rng = np.random.default_rng(seed)
    polys, offs, pts, res = [], [0], [], []
    for p in range(n_poly):
        k = int(rng.integers(3, 13))
        span = int(rng.choice([6, 40, 1000]))
        poly = rng.integers(0, span + 1, (k, 2)).astype(np.int32)
        if p % 7 == 0:
            poly[1] = poly[0]                              # repeated vertex
        if p % 11 == 0:
            poly[:, 1] = poly[0, 1]                        # collinear / zero area
        q = rng.integers(-2, span + 3, (n_pts, 2)).astype(np.int32)
        q[: k] = poly                                      # exactly on vertices
        mid = (poly + np.roll(poly, -1, 0)) // 2
        q[k: 2 * k] = mid[: max(0, min(k, n_pts - k))]     # near / on edges
        r = [int(cv2.pointPolygonTest(poly, (int(x), int(y)), False)) for x, y in q]
        polys.append(poly)
        offs.append(offs[-1] + k)
        pts.append(q)
        res.append(r)
    return dict(poly_xy=np.concatenate(polys), poly_offsets=np.array(offs, np.int64),
                points=np.stack(pts), result=np.array(res, np.int8))


def main():
    ref_tracker, ref_zones = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # config 1 of BASELINE.json: 300 frames, 1280x720, 20 objects, 4 zones
    clip = synth.scripted_clip(seed=0)
    zones = synth.make_zones(seed=0, num_zones=4)
    g = run_reference_clip(ref_tracker, ref_zones, clip, zones)
    np.savez_compressed(os.path.join(GOLDEN, "cfg1_clip.npz"), **g)
    print("cfg1:", "next_id", g["next_id"][-1], "events", len(g["events"]),
          "max tracks", np.diff(g["state_offsets"]).max(), "returned", g["returned"].sum())

    # a crowded clip (100 objects, 16 zones incl. a duplicated name, short dwell) where greedy
    # conflicts, stage-2 matches, prunes and cooldown re-fires all occur
    clip = synth.scripted_clip(seed=3, num_frames=120, num_objects=100, w_range=(40, 120),
                               h_range=(60, 200), vmax=3.0, dropout=0.1)
    zones = synth.make_zones(seed=3, num_zones=16, kmin=4, kmax=12, dwell_time_sec=0.2,
                             cooldown_sec=0.5)
    zones[5]["name"] = zones[2]["name"]                    # duplicate names share state
    zones[5]["trigger"] = "crossing"                       # ... but stay tellable apart in events
    zones[7].pop("dwell_time_sec")                         # default 2.0 (zone_engine.py:148)
    g = run_reference_clip(ref_tracker, ref_zones, clip, zones,
                           tracker_kwargs=dict(bytetrack=dict(track_thresh=0.5, track_buffer=12,
                                                              match_thresh=0.7, mot20=False)))
    g["track_buffer"] = np.int64(12)
    g["match_thresh"] = np.float64(0.7)
    np.savez_compressed(os.path.join(GOLDEN, "crowd_clip.npz"), **g)
    print("crowd:", "next_id", g["next_id"][-1], "events", len(g["events"]),
          "max tracks", np.diff(g["state_offsets"]).max())

    # a clip with empty frames (age-only path, tracker.py:70-73) and all-low frames
    clip = synth.scripted_clip(seed=5, num_frames=90, num_objects=8)
    for f in (10, 11, 12, 40) + tuple(range(50, 85)):
        clip[f] = (np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32))
    for f in (20, 21):
        b, c, k = clip[f]
        clip[f] = (b, np.minimum(c, np.float32(0.45)), k)
    g = run_reference_clip(ref_tracker, ref_zones, clip, synth.make_zones(seed=5, num_zones=3))
    np.savez_compressed(os.path.join(GOLDEN, "gaps_clip.npz"), **g)
    print("gaps:", "next_id", g["next_id"][-1], "events", len(g["events"]),
          "max tracks", np.diff(g["state_offsets"]).max(), "max tsu", g["state_tsu"].max())

    # high churn: small fast boxes at the default IoU >= 0.8 floor -> many births, stage-2
    # matches and prunes at track_buffer = 30
    clip = synth.scripted_clip(seed=9, num_frames=150, num_objects=40, w_range=(18, 50),
                               h_range=(30, 90), vmax=4.0, dropout=0.08)
    g = run_reference_clip(ref_tracker, ref_zones, clip,
                           synth.make_zones(seed=9, num_zones=6, kmin=3, kmax=9,
                                            dwell_time_sec=0.1, cooldown_sec=0.3))
    np.savez_compressed(os.path.join(GOLDEN, "churn_clip.npz"), **g)
    print("churn:", "next_id", g["next_id"][-1], "events", len(g["events"]),
          "max tracks", np.diff(g["state_offsets"]).max())

    np.savez_compressed(os.path.join(GOLDEN, "pip_cases.npz"), **golden_pip(ref_zones))
    print("pip cases written")


if __name__ == "__main__":
    main()
