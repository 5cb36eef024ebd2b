"""BASELINE.json configs[2] at its full size (64 streams x 8400 anchors x 80 classes, 4 zones): every stream and
every frame of the bench's 16-frame cycle against the oracle chain (a few seconds of CPU), in both step modes, and
size-independent properties on top:

* every stream's detections are what ultralytics' post-process can emit: at most max_det, score order,
  above the confidence threshold, wanted classes only, inside the source frame, `keep` indices strictly
  increasing within equal scores (stable sort), and no same-class pair left whose IoU exceeds the NMS
  threshold (checked in letterbox coordinates, where the suppression test is made);
* track ids are unique, ascending in table order and below next_id (tracker.py:128-135); every
  detection that was matched or born points at a live track;
* results do not depend on how the streams are batched: 64 streams in one batch == two batches of 32
  (the property sharding over GPUs relies on, SURVEY section 8 e), bit for bit, also in scan_async mode.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WANTED = [0, 1, 2, 3, 5, 7]
CONF, IOU, MAX_DET = 0.35, 0.45, 100
S, F, STEPS = 64, 4, 12
SRC_H, SRC_W = 1080, 1920


def run(pkg, wl_heads, zones, streams, heads_ready=None):
    import torch
    dev = torch.device("cuda", 0)
    sb = pkg.StreamBatch(len(streams), [zones[s] for s in streams], src_hw=(SRC_H, SRC_W), classes=WANTED,
                         max_tracks=512, device=dev)
    sel = torch.as_tensor(list(streams), device=dev)
    frames = [[t.index_select(0, sel).contiguous() for t in lv] for lv in wl_heads]
    out = []
    for f in range(STEPS):
        sb.step(frames[f % F], now=1.7e9 + f / 30.0, frame_id=f, heads_ready=heads_ready)
        if f >= STEPS - 2:                                     # keep the last two steps' detections and events
            torch.cuda.synchronize()
            out.append((sb.read_detections(), sb.read_events()))
    tracks, next_id = sb.read_tracks()
    return out, tracks, next_id


def iou_matrix(b):
    x1 = np.maximum(b[:, None, 0], b[None, :, 0]); y1 = np.maximum(b[:, None, 1], b[None, :, 1])
    x2 = np.minimum(b[:, None, 2], b[None, :, 2]); y2 = np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (area[:, None] + area[None, :] - inter + 1e-12)


def ev_key(e):   # everything but timestamp_utc, which is the wall clock at decode time (zone_engine.py:108)
    return (e.event_type, e.zone_name, e.track_id, e.class_id, e.dwell_time_sec, tuple(e.bbox_xyxy), tuple(e.centroid), e.frame_id)


@pytest.fixture(scope="module")
def full(pkg):
    import torch
    from rtmodt_b200.workload import PostBackboneWorkload
    wl = PostBackboneWorkload(S, F, first_stream=0, device=torch.device("cuda", 0), dtype=torch.bfloat16)
    whole = run(pkg, wl.heads, wl.zones, range(S))
    return wl, whole


def test_detections_are_valid_post_process_output(pkg, full):
    wl, (steps, _, _) = full
    gain = min(640 / SRC_H, 640 / SRC_W)
    pad_x, pad_y = round((640 - SRC_W * gain) / 2 - 0.1), round((640 - SRC_H * gain) / 2 - 0.1)
    total = 0
    for dets, _ in steps:
        assert len(dets) == S
        for d in dets:
            n = len(d["confidence"])
            total += n
            assert n <= MAX_DET
            if n == 0:
                continue
            c, k, xy = d["confidence"], d["keep"], d["xyxy"]
            assert np.all(c > np.float32(CONF)) and np.all(np.diff(c) <= 0)
            assert np.all((np.diff(c) < 0) | (np.diff(k) > 0))          # stable sort: equal scores keep list order
            assert set(np.unique(d["class_id"])) <= set(WANTED)
            assert np.all(xy[:, [0, 2]] >= 0) and np.all(xy[:, [0, 2]] <= SRC_W)
            assert np.all(xy[:, [1, 3]] >= 0) and np.all(xy[:, [1, 3]] <= SRC_H)
            assert len(np.unique(d["anchor"])) == n and d["anchor"].min() >= 0 and d["anchor"].max() < 8400
            # back to letterbox coordinates (exact up to rounding unless the box was clipped at the frame border)
            inside = (xy[:, 0] > 0) & (xy[:, 1] > 0) & (xy[:, 2] < SRC_W) & (xy[:, 3] < SRC_H)
            lb = xy * gain + np.array([pad_x, pad_y, pad_x, pad_y], np.float32)
            iou = iou_matrix(lb.astype(np.float64))
            same = (d["class_id"][:, None] == d["class_id"][None, :]) & inside[:, None] & inside[None, :]
            np.fill_diagonal(same, False)
            assert not np.any(same & (iou > IOU + 1e-4))
    assert total > S * 10                                               # the planted objects were found


def test_track_tables_are_consistent(pkg, full):
    _, (steps, tracks, next_id) = full
    dets_last, _ = steps[-1]
    for b in range(S):
        ids = np.array([t["track_id"] for t in tracks[b]], np.int64)
        assert len(ids) > 0 and len(np.unique(ids)) == len(ids)
        assert np.all(np.diff(ids) > 0) and ids.min() >= 1 and ids.max() < next_id[b]
        # tracker.py:137-141: every track is aged after the update, so "updated this frame" reads 1, never 0
        fresh = {int(t["track_id"]) for t in tracks[b] if t["time_since_update"] == 1}
        tid = dets_last[b]["track_id"]
        assert {int(i) for i in tid[tid > 0]} == fresh                  # matched or born this frame <=> updated this frame
        for t in tracks[b]:
            assert 1 <= t["time_since_update"] <= 30 and t["age"] >= 1


@pytest.mark.parametrize("heads_ready", [None, True])
def test_results_do_not_depend_on_the_batching(pkg, full, heads_ready):
    wl, whole = full
    halves = [run(pkg, wl.heads, wl.zones, range(0, S // 2), heads_ready), run(pkg, wl.heads, wl.zones, range(S // 2, S), heads_ready)]
    steps_w, tracks_w, next_w = whole
    for h, (steps_h, tracks_h, next_h) in enumerate(halves):
        off = h * (S // 2)
        np.testing.assert_array_equal(next_h, next_w[off:off + S // 2])
        for (dw, ew), (dh, eh) in zip(steps_w, steps_h):
            for b in range(S // 2):
                for key in ("xyxy", "confidence", "class_id", "anchor", "keep", "track_id", "kind"):
                    np.testing.assert_array_equal(dh[b][key], dw[off + b][key])
                assert [ev_key(e) for e in eh[b]] == [ev_key(e) for e in ew[off + b]]
        for b in range(S // 2):
            assert len(tracks_h[b]) == len(tracks_w[off + b])
            for th, tw in zip(tracks_h[b], tracks_w[off + b]):
                assert th["track_id"] == tw["track_id"] and th["age"] == tw["age"]
                np.testing.assert_array_equal(th["xyxy"], tw["xyxy"])


@pytest.mark.parametrize("heads_ready", [None, True])
def test_every_stream_and_frame_matches_the_oracle_chain(pkg, heads_ready):
    """All 64 streams x 16 frames of configs[2] (the bench workload itself): detections (keep sets, boxes within
    1e-4), per-detection track ids, whole track tables, next ids and event lists, frame by frame."""
    import torch
    from oracle import chain
    from rtmodt_b200.workload import PostBackboneWorkload
    dev = torch.device("cuda", 0)
    frames = 16
    wl = PostBackboneWorkload(S, frames, first_stream=0, device=dev, dtype=torch.bfloat16)
    sb = pkg.StreamBatch(S, wl.zones, src_hw=(SRC_H, SRC_W), classes=WANTED, max_tracks=512, device=dev)
    res = chain.run_chain_parity(sb, lambda f: wl.heads[f], lambda f: wl.host_frame(f), wl.zones, frames,
                                 src_hw=(SRC_H, SRC_W), classes=WANTED, heads_ready=heads_ready)
    sb.close()
    assert res["streams"] == S and res["frames"] == frames
    assert res["detections_checked"] > S * frames * 10 and res["events_checked"] > 0
    bad = {k: res[k] for k in ("nms_index_flips", "box_mismatch", "track_id_mismatch", "track_table_mismatch", "event_mismatch")}
    assert res["ok"], bad
