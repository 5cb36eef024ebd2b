"""TEST INFRASTRUCTURE (not product code): the CUDA step of a StreamBatch held, frame by frame and stream by
stream, to the oracle chain - detect_ref (ultralytics' decode + NMS + rescale restated; **parity unpinned**, see
detect_ref.py) -> tracker_ref (pinned to the unmodified reference) -> zone_ref (pinned to the unmodified
reference).  Used by tests/ and by bench.py's parity leg; nothing under the package imports it.

Detections: same count, same anchors (= the same NMS keep set in the same order), boxes within 1e-4 relative
(the float decode is the tolerance-checked stage).  Everything behind the detector is compared bit for bit, with
the oracles fed the DEVICE detections so that a last-ulp difference of the decode cannot masquerade as a
tracker difference: per-detection track ids, the whole track table (`_core._tracks`) and `_next_id`
(tracker.py:58-141), and the event lists (zone_engine.py:82-132) in order.
"""

from __future__ import annotations

import hashlib

import numpy as np

from . import detect_ref, tracker_ref, zone_ref


def event_key(e):
    """Everything of a ZoneEvent but timestamp_utc (the wall clock at decode time, zone_engine.py:108)."""
    return (e.event_type, e.zone_name, int(e.track_id), int(e.class_id), float(e.dwell_time_sec),
            tuple(float(v) for v in e.bbox_xyxy), tuple(int(v) for v in e.centroid), int(e.frame_id))


def run_chain_parity(sb, get_heads, get_host_heads, zones, num_frames, src_hw=(1080, 1920), classes=None,
                     t0=1_700_000_000.0, fps=30.0, heads_ready=None, digests=True, imgsz=(640, 640)):
    """Steps ``sb`` (a fresh StreamBatch over n streams) through ``num_frames`` frames and compares every stream
    with the oracle chain after every frame.

    ``get_heads(f)`` -> the three device head tensors of frame f for the n streams; ``get_host_heads(f)`` -> the
    same as float32 host tensors; ``zones[s]`` -> the zone configs of stream s.  Returns a dict of mismatch
    counters (all zero = parity) and, with ``digests``, one sha1 per stream over everything the CUDA path
    produced (for comparing two CUDA runs, e.g. different shardings, bit for bit).
    """
    n = sb.B
    trk = [tracker_ref.TrackerOracle() for _ in range(n)]
    zon = [zone_ref.ZoneOracle(zones[s]) for s in range(n)]
    hashes = [hashlib.sha1() for _ in range(n)]
    flips = box_bad = id_bad = table_bad = ev_bad = dets = events = 0
    for f in range(num_frames):
        now = t0 + f / fps
        sb.step(get_heads(f), now=now, frame_id=f, heads_ready=heads_ready)
        got = sb.read_detections()
        tracks, next_id = sb.read_tracks()
        evs = sb.read_events()
        ref = detect_ref.detect_post(get_host_heads(f), src_hw, imgsz=imgsz, classes=classes)
        for s in range(n):
            r, g = ref[s], got[s]
            if digests:
                h = hashes[s]
                for k in ("xyxy", "confidence", "class_id", "anchor", "keep", "track_id", "kind"):
                    h.update(np.ascontiguousarray(g[k]).tobytes())
                h.update(np.asarray([[t["track_id"], t["age"], t["time_since_update"], t["class_id"]] for t in tracks[s]], np.int64).tobytes())
                h.update(np.asarray([t["xyxy"] for t in tracks[s]], np.float32).tobytes())
                h.update(repr([event_key(e) for e in evs[s]]).encode())
            dets += len(r["conf"])
            if len(r["conf"]) != len(g["confidence"]) or not np.array_equal(r["anchor"], g["anchor"]):
                flips += 1      # the oracles below still follow the device detections
            else:
                box_bad += int(not np.allclose(g["xyxy"], r["xyxy"], rtol=1e-4, atol=1e-2))
                box_bad += int(not np.array_equal(g["class_id"], r["cls"]))
            tid, _ = trk[s].step(g["xyxy"], g["confidence"], g["class_id"])
            id_bad += int(not np.array_equal(tid, g["track_id"])) + int(trk[s].next_id != int(next_id[s]))
            o = trk[s]
            same = (len(tracks[s]) == len(o.track_id)
                    and [t["track_id"] for t in tracks[s]] == o.track_id.tolist()
                    and [t["age"] for t in tracks[s]] == o.age.tolist()
                    and [t["time_since_update"] for t in tracks[s]] == o.tsu.tolist()
                    and [t["class_id"] for t in tracks[s]] == o.cls.tolist()
                    and np.array_equal(np.asarray([t["xyxy"] for t in tracks[s]], np.float32).reshape(-1, 4), o.xyxy))
            table_bad += int(not same)
            act = o.active_rows()
            exp = zon[s].process(zip(o.track_id[act], o.xyxy[act], o.cls[act]), f, now)
            events += len(exp)
            ev_bad += int([event_key(e) for e in evs[s]] != [event_key(e) for e in exp])
    out = {"ok": not (flips or box_bad or id_bad or table_bad or ev_bad), "frames": num_frames, "streams": n,
           "detections_checked": dets, "events_checked": events, "nms_index_flips": flips, "box_mismatch": box_bad,
           "track_id_mismatch": id_bad, "track_table_mismatch": table_bad, "event_mismatch": ev_bad,
           "checked_against": "oracle chain: torch-CPU decode + torchvision NMS + scale_boxes (restated, unpinned) -> "
                              "tracker / zone restatements pinned to the unmodified reference"}
    if digests:
        out["digests"] = [h.hexdigest() for h in hashes]
    return out
