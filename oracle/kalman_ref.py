"""Oracle for the opt-in motion model.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

PARITY UNPINNED - ABSENT: the reference has no Kalman filter at all (``src/tracking/tracker.py``
matches against the last matched box; ``filterpy`` is listed in requirements.txt:23 and never
imported - SURVEY.md section 0 F1).  BASELINE.json's north star nevertheless names a batched
8-state filter, so the CUDA tracker offers one as an opt-in (default off = the reference).  What
it is checked against is restated here from the published algorithm of ByteTrack
(ifzhang/ByteTrack, ``yolox/tracker/kalman_filter.py`` and ``STrack`` in ``byte_tracker.py``;
not vendored by the reference, not installed here):

  * :class:`KalmanXYAH`      the canonical form: 8-state mean, 8 x 8 covariance, float64 NumPy,
                             ``initiate`` / ``predict`` / ``project`` / ``update`` as published
                             (``std_weight_position = 1/20``, ``std_weight_velocity = 1/160``).
  * :func:`initiate32` / :func:`predict32` / :func:`update32`
                             the same filter in the form the kernel stores it.  With a diagonal
                             initial covariance and diagonal Q / R the 8 x 8 filter is exactly four
                             independent (position, velocity) filters; these functions spell out
                             their float32 arithmetic operation by operation, in the order
                             ``csrc/track_body.cuh`` uses, so that the kernel can be compared bit
                             for bit.  ``tests/test_oracle_kalman.py`` checks them against the
                             canonical form (1e-4 relative, the tolerance the north star states).
  * :func:`predicted_box32`  the box the association sees for a track.

How the filter enters the tracker (``TrackerOracle(use_kalman=True)`` in tracker_ref.py): every
track is predicted once per frame (a track that missed the previous frame has its height velocity
zeroed first, ``STrack.predict``), both association stages use the predicted box instead of the
stored one, a match runs ``update`` with the detection, a birth runs ``initiate``.  Thresholds,
assignment, ageing, pruning and the stored ``xyxy`` stay the reference's.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32
STD_POS = F32(1) / F32(20)
STD_VEL = F32(1) / F32(160)


# ---------------------------------------------------------------------------
# canonical form (float64, matrices) - ByteTrack kalman_filter.py
# ---------------------------------------------------------------------------
class KalmanXYAH:
    """8-dimensional state (x, y, a, h, vx, vy, va, vh), constant velocity, dt = 1."""

    def __init__(self) -> None:
        ndim, dt = 4, 1.0
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    def initiate(self, measurement):
        mean = np.r_[np.asarray(measurement, np.float64), np.zeros(4)]
        h = measurement[3]
        std = [2 * self._std_weight_position * h, 2 * self._std_weight_position * h, 1e-2,
               2 * self._std_weight_position * h, 10 * self._std_weight_velocity * h,
               10 * self._std_weight_velocity * h, 1e-5, 10 * self._std_weight_velocity * h]
        return mean, np.diag(np.square(std))

    def predict(self, mean, covariance):
        std_pos = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-2,
                   self._std_weight_position * mean[3]]
        std_vel = [self._std_weight_velocity * mean[3], self._std_weight_velocity * mean[3], 1e-5,
                   self._std_weight_velocity * mean[3]]
        motion_cov = np.diag(np.square(np.r_[std_pos, std_vel]))
        mean = np.dot(mean, self._motion_mat.T)
        covariance = np.linalg.multi_dot((self._motion_mat, covariance, self._motion_mat.T)) + motion_cov
        return mean, covariance

    def project(self, mean, covariance):
        std = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-1,
               self._std_weight_position * mean[3]]
        innovation_cov = np.diag(np.square(std))
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot((self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + innovation_cov

    def update(self, mean, covariance, measurement):
        projected_mean, projected_cov = self.project(mean, covariance)
        # the published code solves with a Cholesky factor; projected_cov is SPD
        kalman_gain = np.linalg.solve(projected_cov, np.dot(covariance, self._update_mat.T).T).T
        innovation = np.asarray(measurement, np.float64) - projected_mean
        new_mean = mean + np.dot(innovation, kalman_gain.T)
        new_covariance = covariance - np.linalg.multi_dot((kalman_gain, projected_cov, kalman_gain.T))
        return new_mean, new_covariance


def xyxy_to_xyah64(box):
    x1, y1, x2, y2 = (float(v) for v in box)
    w, h = x2 - x1, y2 - y1
    return np.array([x1 + w / 2, y1 + h / 2, w / h, h])


# ---------------------------------------------------------------------------
# decoupled float32 form - operation order of csrc/track_body.cuh
# ---------------------------------------------------------------------------
def xyxy_to_xyah32(box):
    """(..., 4) xyxy float32 -> (..., 4) x, y, a, h   (box_to_xyah)."""
    box = np.asarray(box, F32)
    w = box[..., 2] - box[..., 0]
    h = box[..., 3] - box[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        a = w / h
    return np.stack([box[..., 0] + w * F32(0.5), box[..., 1] + h * F32(0.5), a, h], -1).astype(F32)


def xyah_to_xyxy32(cx, cy, a, h):
    """xyah_to_box."""
    w = a * h
    x1 = cx - w * F32(0.5)
    y1 = cy - h * F32(0.5)
    return np.stack([x1, y1, x1 + w, y1 + h], -1).astype(F32)


def _noise(h, scale, const):
    """(n, 4) variances: (scale * h)^2 for x, y, h and const^2 for the aspect ratio."""
    s = (scale * h).astype(F32)
    v = (s * s).astype(F32)
    out = np.stack([v, v, np.full_like(v, F32(const) * F32(const)), v], -1)
    return out.astype(F32)


def initiate32(boxes):
    """kalman_initiate: (n, 4) boxes -> mean (n, 8), cov (n, 12) [var_p, cov_pv, var_v per coordinate]."""
    z = xyxy_to_xyah32(np.asarray(boxes, F32).reshape(-1, 4))
    n = len(z)
    mean = np.concatenate([z, np.zeros((n, 4), F32)], 1)
    vp = _noise(z[:, 3], F32(2) * STD_POS, 1e-2)
    vv = _noise(z[:, 3], F32(10) * STD_VEL, 1e-5)
    cov = np.zeros((n, 4, 3), F32)
    cov[:, :, 0] = vp
    cov[:, :, 2] = vv
    return mean, cov.reshape(n, 12)


def predict32(mean, cov, tsu_in):
    """kalman_predict: tracks with time_since_update > 1 (missed the previous frame) get vh = 0 first."""
    mean = np.array(mean, F32).reshape(-1, 8)
    c = np.array(cov, F32).reshape(-1, 4, 3)
    mean[np.asarray(tsu_in).reshape(-1) > 1, 7] = 0
    qp = _noise(mean[:, 3], STD_POS, 1e-2)
    qv = _noise(mean[:, 3], STD_VEL, 1e-5)
    pp, pv, vv = c[:, :, 0].copy(), c[:, :, 1].copy(), c[:, :, 2].copy()
    mean[:, :4] = mean[:, :4] + mean[:, 4:]
    c[:, :, 0] = ((pp + pv) + (pv + vv)) + qp
    c[:, :, 1] = pv + vv
    c[:, :, 2] = vv + qv
    return mean, c.reshape(-1, 12)


def update32(mean, cov, boxes):
    """kalman_update (project + update) with measurement boxes (n, 4)."""
    mean = np.array(mean, F32).reshape(-1, 8)
    c = np.array(cov, F32).reshape(-1, 4, 3)
    z = xyxy_to_xyah32(np.asarray(boxes, F32).reshape(-1, 4))
    r = _noise(mean[:, 3], STD_POS, 1e-1)
    pp, pv, vv = c[:, :, 0].copy(), c[:, :, 1].copy(), c[:, :, 2].copy()
    s = pp + r
    kp = pp / s
    kv = pv / s
    y = z - mean[:, :4]
    mean[:, :4] = mean[:, :4] + kp * y
    mean[:, 4:] = mean[:, 4:] + kv * y
    c[:, :, 0] = pp - kp * pp
    c[:, :, 1] = pv - kp * pv
    c[:, :, 2] = vv - kv * pv
    return mean, c.reshape(-1, 12)


def predicted_box32(mean, tsu_in):
    """kalman_predicted_box: the box both association stages see (needs the mean only)."""
    mean = np.asarray(mean, F32).reshape(-1, 8)
    vh = np.where(np.asarray(tsu_in).reshape(-1) > 1, F32(0), mean[:, 7]).astype(F32)
    return xyah_to_xyxy32(mean[:, 0] + mean[:, 4], mean[:, 1] + mean[:, 5], mean[:, 2] + mean[:, 6], mean[:, 3] + vh)


def to_full(mean, cov):
    """Decoupled (8,), (12,) -> canonical mean (8,) and 8 x 8 covariance (for comparison)."""
    c = np.asarray(cov, np.float64).reshape(4, 3)
    full = np.zeros((8, 8))
    for i in range(4):
        full[i, i] = c[i, 0]
        full[i, 4 + i] = full[4 + i, i] = c[i, 1]
        full[4 + i, 4 + i] = c[i, 2]
    return np.asarray(mean, np.float64), full
