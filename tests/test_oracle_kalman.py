"""The decoupled float32 form of the opt-in motion model (what the kernel computes) against the
canonical 8 x 8 matrix form of ByteTrack's KalmanFilter - no GPU needed."""

import numpy as np

from oracle import kalman_ref as kr, tracker_ref

TOL = 1e-4          # relative, the tolerance BASELINE.json states for Kalman states (fp32)


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.all(np.abs(a - b) <= TOL * np.maximum(np.abs(b), 1.0) + 1e-9), (a, b)


def test_decoupled_filter_equals_canonical_matrix_form():
    rng = np.random.default_rng(0)
    kf = kr.KalmanXYAH()
    for _ in range(20):
        c = rng.uniform(100, 900, 2)
        wh = rng.uniform(20, 300, 2)
        v = rng.uniform(-6, 6, 2)
        box = np.r_[c - wh / 2, c + wh / 2].astype(np.float32)
        mean64, cov64 = kf.initiate(kr.xyxy_to_xyah64(box))
        m32, c32 = kr.initiate32(box[None])
        close(kr.to_full(m32[0], c32[0])[0], mean64)
        close(kr.to_full(m32[0], c32[0])[1], cov64)
        tsu = 1
        for step in range(40):
            c = c + v
            wh = wh * rng.uniform(0.98, 1.02, 2)
            box = (np.r_[c - wh / 2, c + wh / 2] + rng.normal(0, 1.0, 4)).astype(np.float32)
            missed = rng.uniform() < 0.2
            if tsu > 1:
                mean64 = mean64.copy()
                mean64[7] = 0.0                            # STrack.predict for a track that is not 'Tracked'
            mean64, cov64 = kf.predict(mean64, cov64)
            pbox = kr.predicted_box32(m32, [tsu])[0]
            m32, c32 = kr.predict32(m32, c32, [tsu])
            close(kr.xyah_to_xyxy32(*[np.float32(x) for x in mean64[:4]]), pbox)
            if not missed:
                mean64, cov64 = kf.update(mean64, cov64, kr.xyxy_to_xyah64(box))
                m32, c32 = kr.update32(m32, c32, box[None])
                tsu = 1
            else:
                tsu += 1
            full_m, full_c = kr.to_full(m32[0], c32[0])
            close(full_m, mean64)
            close(full_c, cov64)
            # the canonical covariance stays block-diagonal per coordinate: nothing is lost by the 12-number form
            off = cov64.copy()
            for i in range(4):
                off[i, i] = off[i, 4 + i] = off[4 + i, i] = off[4 + i, 4 + i] = 0
            assert np.all(off == 0)


def test_tracker_oracle_with_motion_model_bridges_a_detection_gap():
    """Objects at 4 px / frame match the stored box frame to frame (IoU ~0.85 >= the reference's
    0.8 floor) in both modes.  After six frames without detections the stored box is 28 px behind
    (IoU ~0.5): the reference's tracker starts new identities, the predicted box keeps the old ones.
    (With the 0.8 floor a filter can only help once a track has been matched a few times: a track
    born with zero velocity predicts its own birth box.)"""
    rng = np.random.default_rng(3)
    n_obj, frames, gap = 8, 60, range(30, 36)
    c = np.stack([np.linspace(150, 1700, n_obj), rng.uniform(300, 700, n_obj)], 1)
    wh = np.stack([rng.uniform(75, 90, n_obj), rng.uniform(120, 160, n_obj)], 1)
    v = np.stack([4.0 * rng.choice([-1, 1], n_obj), np.zeros(n_obj)], 1)
    plain, kal = tracker_ref.TrackerOracle(), tracker_ref.TrackerOracle(use_kalman=True)
    for f in range(frames):
        c = c + v
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        if f in gap:
            boxes = boxes[:0]
        conf = np.full(len(boxes), 0.9, np.float32)
        cls = np.zeros(len(boxes), np.int32)
        plain.step(boxes, conf, cls)
        kal.step(boxes, conf, cls)
    assert plain.next_id - 1 == 2 * n_obj                 # every identity broke at the gap
    assert kal.next_id - 1 == n_obj                       # none did
    assert len(kal.kf_mean) == len(kal) and kal.kf_mean.dtype == np.float32


def test_lapjv_emulation_is_the_exhaustive_optimum_on_small_matrices():
    """The scipy emulation of lap.lapjv(extend_cost=True, cost_limit) against brute force: over all
    partial matchings that use only pairs cheaper than the limit, it attains the minimum of
    sum(cost of matched pairs) + limit/2 * (unmatched rows + unmatched columns)."""
    import itertools
    rng = np.random.default_rng(11)
    thresh = 0.8
    limit = 1 - thresh
    for _ in range(300):
        t, n = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        iou = rng.uniform(0.6, 1.0, (t, n)).astype(np.float32)
        cost = (np.float32(1) - iou).astype(np.float64)
        best = None
        for k in range(0, min(t, n) + 1):
            for rows in itertools.combinations(range(t), k):
                for cols in itertools.permutations(range(n), k):
                    if any(cost[r, c] >= limit for r, c in zip(rows, cols)):
                        continue
                    total = sum(cost[r, c] for r, c in zip(rows, cols)) + limit / 2 * ((t - k) + (n - k))
                    if best is None or total < best - 1e-15:
                        best = total
        rows, cols, ur, uc = tracker_ref.assign_lapjv_emulated(iou, thresh)
        got = sum(cost[r, c] for r, c in zip(rows, cols)) + limit / 2 * (len(ur) + len(uc))
        assert abs(got - best) < 1e-12, (iou, rows, cols)
        assert all(cost[r, c] < limit for r, c in zip(rows, cols))
        assert sorted(rows + ur) == list(range(t)) and sorted(cols + uc) == list(range(n))
