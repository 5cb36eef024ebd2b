"""Oracle: CPU restatement of the reference's ByteTrack-style association core.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/src/tracking/tracker.py``:

  * ``pairwise_iou``      <- ``_ByteTrackCore._batch_iou``          (tracker.py:150-161)
  * ``assign_rowloop``    <- ``_ByteTrackCore._linear_assignment``  (tracker.py:182-194,
                             the greedy branch; ``lap`` is not installed so this is the
                             branch the reference runs here - SURVEY.md §0 F4)
  * ``assign_columnwin``  -  the order-free form of the same rule used by the CUDA
                             kernel (first arg-max per row, smallest admissible row
                             wins the column, no second choice)
  * ``TrackerOracle.step``<- ``_ByteTrackCore.update`` + ``_age_tracks``
                             (tracker.py:58-148)

State is kept as parallel arrays (one row per track, creation order) instead of
the reference's list of dicts; ``as_dicts()`` renders the reference's view for
comparison with ``tracker._core._tracks``.

All arithmetic is float32, exactly as NumPy evaluates the reference's expressions
on float32 detections (NumPy-2 weak scalars: ``np.maximum(0, x)``, ``union + 1e-6``
and ``iou >= thresh`` all stay in float32 - SURVEY.md §7 "hard parts").
"""

from __future__ import annotations

import numpy as np

# kinds reported per detection by ``TrackerOracle.step`` (and by rtm_track_step)
KIND_NONE, KIND_STAGE1, KIND_STAGE2, KIND_BIRTH = 0, 1, 2, 3


def pairwise_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(T,4) x (N,4) float32 -> (T,N) float32, tracker.py:150-161."""
    a = np.asarray(a, np.float32).reshape(-1, 4)
    b = np.asarray(b, np.float32).reshape(-1, 4)
    ax1, ay1, ax2, ay2 = (a[:, k][:, None] for k in range(4))
    bx1, by1, bx2, by2 = (b[:, k][None, :] for k in range(4))
    iw = np.maximum(np.float32(0), np.minimum(ax2, bx2) - np.maximum(ax1, bx1))
    ih = np.maximum(np.float32(0), np.minimum(ay2, by2) - np.maximum(ay1, by1))
    inter = iw * ih
    area_a = (ax2 - ax1) * (ay2 - ay1)
    area_b = (bx2 - bx1) * (by2 - by1)
    union = (area_a + area_b) - inter
    return inter / (union + np.float32(1e-6))


def assign_rowloop(iou: np.ndarray, thresh: float):
    """Greedy row-order arg-max with no second choice, tracker.py:184-194.

    Returns ``(matched_rows, matched_cols, unmatched_rows, unmatched_cols)`` as
    ascending int lists (matched_* in row order).
    """
    t32 = np.float32(thresh)
    rows, cols = [], []
    taken = np.zeros(iou.shape[1], bool)
    hit = np.zeros(iou.shape[0], bool)
    for r in range(iou.shape[0]):
        c = int(np.argmax(iou[r]))
        if iou[r, c] >= t32 and not taken[c]:
            rows.append(r)
            cols.append(c)
            taken[c] = True
            hit[r] = True
    return rows, cols, np.flatnonzero(~hit).tolist(), np.flatnonzero(~taken).tolist()


def assign_columnwin(iou: np.ndarray, thresh: float):
    """Same result as :func:`assign_rowloop`, without the row-order loop.

    ``j*_r`` = first arg-max of row r; row r is admissible iff
    ``iou[r, j*_r] >= thresh``; column j goes to the smallest admissible r with
    ``j*_r == j``; every other row stays unmatched.  This is what
    ``rtm_track_step`` computes with one ``atomicMin`` per admissible row.
    """
    t_n, d_n = iou.shape
    if t_n == 0 or d_n == 0:
        return [], [], list(range(t_n)), list(range(d_n))
    best = np.argmax(iou, axis=1)
    ok = iou[np.arange(t_n), best] >= np.float32(thresh)
    winner = np.full(d_n, t_n, np.int64)
    np.minimum.at(winner, best[ok], np.flatnonzero(ok))
    rows = np.flatnonzero(ok & (winner[best] == np.arange(t_n)))
    cols = best[rows]
    hit = np.zeros(t_n, bool)
    hit[rows] = True
    taken = np.zeros(d_n, bool)
    taken[cols] = True
    return rows.tolist(), cols.tolist(), np.flatnonzero(~hit).tolist(), np.flatnonzero(~taken).tolist()


def assign_lapjv_emulated(iou: np.ndarray, thresh: float):
    """The branch the reference takes when ``lap`` is installed (tracker.py:168-181):
    ``lap.lapjv(1 - iou, extend_cost=True, cost_limit=1 - thresh)``.

    PARITY UNPINNED: ``lap>=0.4.0`` (requirements.txt:22) is not installed here and no reference
    test fixes its output.  Restated from its published behaviour: with ``cost_limit`` lap solves
    the square problem ``[[C, L/2], [L/2, 0]]`` of size T + N (L = cost_limit, C in float64) and
    reports row i as unmatched when its column is one of the T padding columns.  Solved here with
    ``scipy.optimize.linear_sum_assignment``; equal to lapjv wherever the optimum is unique (ties are
    broken solver by solver).
    """
    from scipy.optimize import linear_sum_assignment
    t_n, d_n = iou.shape
    if t_n == 0 or d_n == 0:
        return [], [], list(range(t_n)), list(range(d_n))
    limit = 1 - thresh                                     # tracker.py:170, Python float
    cost = (np.float32(1) - np.asarray(iou, np.float32)).astype(np.float64)   # tracker.py:167, then lap's cast
    ext = np.full((t_n + d_n, t_n + d_n), limit / 2.0)
    ext[t_n:, d_n:] = 0.0
    ext[:t_n, :d_n] = cost
    r_ind, c_ind = linear_sum_assignment(ext)
    rows = [int(r) for r, c in zip(r_ind, c_ind) if r < t_n and c < d_n]
    cols = [int(c) for r, c in zip(r_ind, c_ind) if r < t_n and c < d_n]
    hit = np.zeros(t_n, bool)
    hit[rows] = True
    taken = np.zeros(d_n, bool)
    taken[cols] = True
    return rows, cols, np.flatnonzero(~hit).tolist(), np.flatnonzero(~taken).tolist()


class TrackerOracle:
    """One stream's tracker state + step, tracker.py:43-148."""

    def __init__(self, track_thresh: float = 0.5, track_buffer: int = 30,
                 match_thresh: float = 0.8, assign=assign_rowloop, use_kalman: bool = False) -> None:
        # use_kalman: opt-in motion model that the reference does NOT have;
        # see oracle/kalman_ref.py.  Off = the reference's behaviour.
        self.use_kalman = use_kalman
        self.kf_mean = np.zeros((0, 8), np.float32)
        self.kf_cov = np.zeros((0, 12), np.float32)
        self.track_thresh = track_thresh
        self.track_buffer = track_buffer
        self.match_thresh = match_thresh
        self.assign = assign
        self.next_id = 1                                   # tracker.py:55
        self.track_id = np.zeros(0, np.int32)
        self.xyxy = np.zeros((0, 4), np.float32)
        self.conf = np.zeros(0, np.float32)
        self.cls = np.zeros(0, np.int32)
        self.age = np.zeros(0, np.int32)
        self.tsu = np.zeros(0, np.int32)                   # time_since_update

    def __len__(self) -> int:
        return len(self.track_id)

    # -- views ------------------------------------------------------------
    def as_dicts(self):
        """The reference's ``_core._tracks`` view (list of dicts, creation order)."""
        return [dict(track_id=int(i), xyxy=b.copy(), confidence=float(c), class_id=int(k),
                     age=int(a), time_since_update=int(t))
                for i, b, c, k, a, t in zip(self.track_id, self.xyxy, self.conf, self.cls,
                                            self.age, self.tsu)]

    def active_rows(self) -> np.ndarray:
        """Rows matched or born in the last step (``time_since_update == 1``).

        The reference's own return filter (``== 0``, tracker.py:141) is always empty
        because ``_age_tracks`` has already bumped every track.
        """
        return np.flatnonzero(self.tsu == 1)

    # -- one frame --------------------------------------------------------
    def step(self, xyxy, conf, cls):
        """tracker.py:58-141.  Returns ``(det_track_id i32 (N,), det_kind i32 (N,))``:
        for every input detection the id of the track it updated or created
        (0 = discarded low-score detection) and how (KIND_*)."""
        xyxy = np.asarray(xyxy, np.float32).reshape(-1, 4)
        conf = np.asarray(conf, np.float32).reshape(-1)
        cls = np.asarray(cls, np.int32).reshape(-1)
        n = len(conf)
        det_tid = np.zeros(n, np.int32)
        det_kind = np.zeros(n, np.int32)
        tsu_in = self.tsu.copy()
        if n == 0:                                         # tracker.py:70-73: age only, no prune
            if self.use_kalman and len(self):
                from . import kalman_ref
                self.kf_mean, self.kf_cov = kalman_ref.predict32(self.kf_mean, self.kf_cov, tsu_in)
            self.tsu = self.tsu + 1
            return det_tid, det_kind
        self._matched = []                                 # (rows, dets) of this step, for the filter update
        if self.use_kalman:
            from . import kalman_ref
            seen = kalman_ref.predicted_box32(self.kf_mean, tsu_in)   # what the association sees
        else:
            seen = self.xyxy

        high = np.flatnonzero(conf >= np.float32(self.track_thresh))   # tracker.py:76
        low = np.flatnonzero(~(conf >= np.float32(self.track_thresh)))  # tracker.py:77
        t_n = len(self)

        # stage 1: every retained track x high-score detections (tracker.py:91-106)
        if t_n and len(high):
            m_t, m_d, rest_t, rest_d = self.assign(pairwise_iou(seen, xyxy[high]),
                                                   self.match_thresh)
            self._commit(np.asarray(m_t, np.int64), high[np.asarray(m_d, np.int64)],
                         xyxy, conf, cls, det_tid, det_kind, KIND_STAGE1)
        else:
            rest_t, rest_d = list(range(t_n)), list(range(len(high)))
        rest_t = np.asarray(rest_t, np.int64)

        # stage 2: still-unmatched tracks x low-score detections, same threshold
        # (tracker.py:109-123); unmatched low detections are dropped
        if len(rest_t) and len(low):
            m_t, m_d, _, _ = self.assign(pairwise_iou(seen[rest_t], xyxy[low]),
                                         self.match_thresh)
            self._commit(rest_t[np.asarray(m_t, np.int64)], low[np.asarray(m_d, np.int64)],
                         xyxy, conf, cls, det_tid, det_kind, KIND_STAGE2)

        if self.use_kalman and t_n:                        # predict every track, update the matched ones
            self.kf_mean, self.kf_cov = kalman_ref.predict32(self.kf_mean, self.kf_cov, tsu_in)
            for rows, dets in self._matched:
                self.kf_mean[rows], self.kf_cov[rows] = kalman_ref.update32(self.kf_mean[rows], self.kf_cov[rows], xyxy[dets])

        # births from unmatched high detections, ascending index (tracker.py:126-135)
        born = high[np.asarray(rest_d, np.int64)]
        k = len(born)
        if k:
            ids = np.arange(self.next_id, self.next_id + k, dtype=np.int32)
            self.track_id = np.concatenate([self.track_id, ids])
            self.xyxy = np.concatenate([self.xyxy, xyxy[born]])
            self.conf = np.concatenate([self.conf, conf[born]])
            self.cls = np.concatenate([self.cls, cls[born]])
            self.age = np.concatenate([self.age, np.ones(k, np.int32)])
            self.tsu = np.concatenate([self.tsu, np.zeros(k, np.int32)])
            self.next_id += k
            det_tid[born] = ids
            det_kind[born] = KIND_BIRTH
            if self.use_kalman:
                m0, c0 = kalman_ref.initiate32(xyxy[born])
                self.kf_mean = np.concatenate([self.kf_mean, m0])
                self.kf_cov = np.concatenate([self.kf_cov, c0])

        # age everything, then drop tracks past the buffer (tracker.py:138-139, 144-147)
        self.tsu = self.tsu + 1
        live = self.tsu <= self.track_buffer
        for name in ("track_id", "xyxy", "conf", "cls", "age", "tsu") + (("kf_mean", "kf_cov") if self.use_kalman else ()):
            setattr(self, name, getattr(self, name)[live])
        return det_tid, det_kind

    def _commit(self, rows, dets, xyxy, conf, cls, det_tid, det_kind, kind) -> None:
        """Matched rows take the detection's box / score / class (tracker.py:99-104)."""
        if len(rows) == 0:
            return
        self._matched.append((rows, dets))
        self.xyxy[rows] = xyxy[dets]
        self.conf[rows] = conf[dets]
        self.cls[rows] = cls[dets]
        self.age[rows] += 1
        self.tsu[rows] = 0
        det_tid[dets] = self.track_id[rows]
        det_kind[dets] = kind
