"""Randomised soak of the NMS stage against torchvision (bit-exact keep indices): many shapes of
candidate sets - a handful to a few thousand candidates, one class to all classes, score ties,
huge and degenerate boxes, small max_det - so that every path of nms_body.cuh (counting-rank sort
and its bitonic fallback, warp heads / filter / tails, block-wide scan, spill arrays) is hit with
inputs nobody hand-picked."""

import numpy as np
import pytest

from test_gpu_detect import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(40))
def test_random_candidate_sets(pkg, seed):
    rng = np.random.default_rng(1000 + seed)
    A = int(rng.choice([64, 400, 2100, 8400]))
    nc = int(rng.choice([1, 3, 20, 80]))
    B = 2
    n_obj = int(rng.integers(1, max(2, A // 40)))
    dup = int(rng.integers(1, 12))
    mode = rng.choice(["plain", "ties", "quantised", "huge", "degenerate"])
    pred = np.zeros((B, 4 + nc, A), np.float32)
    pred[:, 4:] = rng.uniform(0, 0.02, (B, nc, A))
    pred[:, 0] = rng.uniform(0, 640, (B, A)); pred[:, 1] = rng.uniform(0, 640, (B, A))
    pred[:, 2] = rng.uniform(4, 120, (B, A)); pred[:, 3] = rng.uniform(4, 120, (B, A))
    size = (2000, 12000) if mode == "huge" else (10, 150)
    for b in range(B):
        for _ in range(n_obj):
            c = int(rng.integers(0, nc))
            base = np.array([rng.uniform(30, 610), rng.uniform(30, 610), rng.uniform(*size), rng.uniform(*size)])
            for a in rng.choice(A, min(dup, A), replace=False):
                box = base + rng.normal(0, 3.0, 4)
                if mode == "degenerate" and rng.uniform() < 0.3:
                    box[2:] = 0.0                                          # zero-area boxes: NaN IoU, never suppressed
                pred[b, :4, a] = box
                s = rng.uniform(0.1, 0.99)
                if mode == "ties":
                    s = float(rng.choice([0.4, 0.6, 0.8]))
                if mode == "quantised":
                    s = np.round(s * 16) / 16                              # few distinct scores: crowded sort buckets
                pred[b, 4 + c, a] = s
    kw = dict(conf=float(rng.choice([0.05, 0.25, 0.35])), iou=float(rng.choice([0.3, 0.45, 0.7])),
              classes=None if rng.uniform() < 0.6 else sorted(rng.choice(nc, max(1, nc // 2), replace=False).tolist()),
              agnostic=bool(rng.uniform() < 0.2), max_det=int(rng.choice([1, 5, 100, 300])))
    check_against_oracle(pkg, pred, **kw)
