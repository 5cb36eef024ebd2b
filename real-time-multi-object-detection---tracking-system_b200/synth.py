"""Seeded synthetic workloads for the five BASELINE.json configs (SURVEY.md §8 d).

Everything here is plain NumPy on the host so the same bytes can be fed to the
CUDA path, to the oracle and (in the dev container) to the unmodified reference.
Nothing in this module touches the GPU or the oracle.

* ``scripted_clip``      – scripted moving boxes for one stream (configs 1 and 5).
* ``scripted_batch``     – the same for B streams, padded to a fixed slot count.
* ``make_zones``         – the two zones of ``config/default.yaml:68-77`` plus
                           seeded convex polygons.
* ``plant_head``         – a synthetic YOLOv8 head tensor (3 levels, 64 DFL box
                           logits + 80 class logits per anchor) with planted
                           objects (configs 2-4).  A random-init YOLOv8s never
                           scores above 0.35, so detections have to be planted
                           (SURVEY.md §7 "hard parts", last bullet).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

#: classes kept by the reference's config (``config/default.yaml:39``)
WANTED_CLASSES = (0, 1, 2, 3, 5, 7)
#: the strides / grid sizes of a 640x640 YOLOv8 head
STRIDES = (8, 16, 32)
REG_MAX = 16
NUM_CLASSES = 80
NUM_OUT = 4 * REG_MAX + NUM_CLASSES  # 144 channels per anchor


# ---------------------------------------------------------------------------
# Scripted detections (tracker / zone workloads)
# ---------------------------------------------------------------------------
@dataclass
class ScriptedObjects:
    """State of the scripted movers of one stream."""

    cxy: np.ndarray   # (M, 2) float64 centres
    wh: np.ndarray    # (M, 2) float64 sizes
    vel: np.ndarray   # (M, 2) float64 px / frame
    cls: np.ndarray   # (M,)   int32


def _spawn(rng: np.random.Generator, m: int, width: int, height: int,
           w_range, h_range, vmax: float) -> ScriptedObjects:
    wh = np.stack([rng.uniform(*w_range, m), rng.uniform(*h_range, m)], 1)
    cxy = np.stack([rng.uniform(wh[:, 0] / 2, width - wh[:, 0] / 2),
                    rng.uniform(wh[:, 1] / 2, height - wh[:, 1] / 2)], 1)
    vel = rng.uniform(-vmax, vmax, (m, 2))
    cls = rng.choice(np.asarray(WANTED_CLASSES, np.int32), m).astype(np.int32)
    return ScriptedObjects(cxy, wh, vel, cls)


def _advance(o: ScriptedObjects, width: int, height: int) -> None:
    o.cxy += o.vel
    lim = np.array([width, height], np.float64)
    lo = o.wh / 2
    hi = lim - o.wh / 2
    under = o.cxy < lo
    over = o.cxy > hi
    o.cxy = np.where(under, 2 * lo - o.cxy, o.cxy)
    o.cxy = np.where(over, 2 * hi - o.cxy, o.cxy)
    o.vel = np.where(under | over, -o.vel, o.vel)


def scripted_clip(seed: int = 0, num_frames: int = 300, width: int = 1280, height: int = 720,
                  num_objects: int = 20, w_range=(60.0, 160.0), h_range=(120.0, 320.0),
                  vmax: float = 2.0, dropout: float = 0.05, conf_range=(0.36, 0.95)):
    """Config 1 of BASELINE.json (SURVEY.md §8 d): one stream of scripted movers.

    Returns a list with one ``(xyxy f32 (N,4), conf f32 (N,), cls i32 (N,))`` per
    frame.  About a quarter of the scores fall below ``track_thresh`` so the
    second association stage is exercised, and ``dropout`` makes tracks go
    unmatched for a frame now and then.
    """
    rng = np.random.default_rng(seed)
    objs = _spawn(rng, num_objects, width, height, w_range, h_range, vmax)
    frames = []
    for _ in range(num_frames):
        _advance(objs, width, height)
        keep = rng.uniform(size=num_objects) >= dropout
        conf = rng.uniform(*conf_range, num_objects).astype(np.float32)
        half = objs.wh / 2
        xyxy = np.concatenate([objs.cxy - half, objs.cxy + half], 1).astype(np.float32)
        frames.append((np.ascontiguousarray(xyxy[keep]), conf[keep], objs.cls[keep].copy()))
    return frames


def scripted_batch(num_streams: int, num_frames: int, slots: int, seed: int = 0, **kw):
    """B independent clips packed into fixed-slot arrays.

    Returns ``(xyxy (F,B,slots,4) f32, conf (F,B,slots) f32, cls (F,B,slots) i32,
    count (F,B) i32)``; stream ``b`` uses ``seed + b``.
    """
    xyxy = np.zeros((num_frames, num_streams, slots, 4), np.float32)
    conf = np.zeros((num_frames, num_streams, slots), np.float32)
    cls = np.zeros((num_frames, num_streams, slots), np.int32)
    count = np.zeros((num_frames, num_streams), np.int32)
    for b in range(num_streams):
        clip = scripted_clip(seed=seed + b, num_frames=num_frames, **kw)
        for f, (bx, cf, cl) in enumerate(clip):
            n = len(cf)
            if n > slots:
                raise ValueError(f"stream {b} frame {f}: {n} detections > {slots} slots")
            xyxy[f, b, :n], conf[f, b, :n], cls[f, b, :n], count[f, b] = bx, cf, cl, n
    return xyxy, conf, cls, count


def dense_crowd_kwargs(num_objects: int = 1000):
    """Config 5 (MOT20-like) parameters for :func:`scripted_clip`."""
    return dict(width=1920, height=1080, num_objects=num_objects, w_range=(20.0, 60.0),
                h_range=(50.0, 150.0), vmax=1.5)


# ---------------------------------------------------------------------------
# Zones
# ---------------------------------------------------------------------------
def make_zones(seed: int = 0, num_zones: int = 4, width: int = 1280, height: int = 720,
               dwell_time_sec: float = 0.5, cooldown_sec: float = 2.0, kmin: int = 6, kmax: int = 6):
    """Zone configs (the dict schema of ``zone_engine.py:142-151``).

    The first two are the rectangles of ``config/default.yaml:68-77``; the rest
    are seeded convex polygons with ``K`` in ``[kmin, kmax]`` (hexagons by
    default, K in [4, 12] for the dense-crowd config).
    """
    rng = np.random.default_rng(10_000 + seed)
    zones = [
        dict(name="restricted_area_1", polygon=[[100, 200], [400, 200], [400, 600], [100, 600]],
             trigger="intrusion", dwell_time_sec=dwell_time_sec, cooldown_sec=cooldown_sec),
        dict(name="exit_gate", polygon=[[800, 400], [1200, 400], [1200, 700], [800, 700]],
             trigger="crossing", direction="left_to_right", dwell_time_sec=dwell_time_sec,
             cooldown_sec=cooldown_sec),
    ][:num_zones]
    for z in range(len(zones), num_zones):
        k = int(rng.integers(kmin, kmax + 1))
        c = np.array([rng.uniform(0.15, 0.85) * width, rng.uniform(0.15, 0.85) * height])
        r = rng.uniform(0.08, 0.2) * min(width, height)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        pts = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
        zones.append(dict(name=f"zone_{z}", polygon=np.rint(pts).astype(np.int64).tolist(),
                          trigger="intrusion", dwell_time_sec=dwell_time_sec,
                          cooldown_sec=cooldown_sec))
    return zones


# ---------------------------------------------------------------------------
# Synthetic YOLOv8 head tensors
# ---------------------------------------------------------------------------
def head_shapes(imgsz=(640, 640)):
    """[(h, w)] of the three head levels for an ``imgsz`` = (H, W) network input."""
    return [(imgsz[0] // s, imgsz[1] // s) for s in STRIDES]


def num_anchors(imgsz=(640, 640)) -> int:
    return sum(h * w for h, w in head_shapes(imgsz))


def plant_head(rng: np.random.Generator, boxes_xyxy: np.ndarray, classes: np.ndarray,
               imgsz=(640, 640), logit_range=(-0.5, 3.0), distractor_frac: float = 0.1,
               bg_mean: float = -6.0, peak: float = 6.0, dtype=np.float32):
    """One frame's head tensors ``[(144, h, w)] * 3`` with planted objects.

    Background: class logits ~ N(bg_mean, 1), box logits ~ N(0, 1).  Each object
    (letterbox-pixel ``xyxy``) lights, on every level whose DFL range can reach
    its sides, the anchors within one cell of its centre (up to 3x3) with a class
    logit ~ U(*logit_range) and DFL logits peaked (two-bin interpolation, +peak)
    at the true side distances, so the decoded box is close to the planted one.
    ``distractor_frac`` of the objects get a class outside ``WANTED_CLASSES``.
    """
    shapes = head_shapes(imgsz)
    heads = []
    for (h, w) in shapes:
        t = np.empty((NUM_OUT, h, w), np.float32)
        t[:4 * REG_MAX] = rng.standard_normal((4 * REG_MAX, h, w), np.float32)
        t[4 * REG_MAX:] = rng.standard_normal((NUM_CLASSES, h, w), np.float32) + bg_mean
        heads.append(t)
    unwanted = np.setdiff1d(np.arange(NUM_CLASSES), WANTED_CLASSES)
    for (x1, y1, x2, y2), c in zip(np.asarray(boxes_xyxy, np.float64), classes):
        if rng.uniform() < distractor_frac:
            c = int(rng.choice(unwanted))
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        for t, s, (h, w) in zip(heads, STRIDES, shapes):
            gx0, gy0 = int(cx / s), int(cy / s)
            for gy in range(max(gy0 - 1, 0), min(gy0 + 2, h)):
                for gx in range(max(gx0 - 1, 0), min(gx0 + 2, w)):
                    ax, ay = (gx + 0.5) * s, (gy + 0.5) * s
                    dist = np.array([ax - x1, ay - y1, x2 - ax, y2 - ay]) / s
                    if dist.min() < 0.0 or dist.max() > REG_MAX - 1.001:
                        continue
                    t[4 * REG_MAX + int(c), gy, gx] = rng.uniform(*logit_range)
                    for side in range(4):
                        lo = int(np.floor(dist[side]))
                        fr = dist[side] - lo
                        base = side * REG_MAX
                        # softmax over {lo, lo+1} with weights (1-fr, fr) reproduces dist
                        t[base + lo, gy, gx] = peak + np.log(max(1.0 - fr, 1e-3))
                        t[base + lo + 1, gy, gx] = peak + np.log(max(fr, 1e-3))
    return [t.astype(dtype) if dtype is not np.float32 else t for t in heads]


def letterbox_params(src_hw, imgsz=(640, 640)):
    """Scale-preserving letterbox geometry (ultralytics ``LetterBox``, auto=False).

    Returns ``dict(r, new_w, new_h, top, left)`` following SURVEY.md §3.2:
    ``r = min(H/h0, W/w0)``, ``new_unpad = round(w0*r), round(h0*r)``, padding split
    with the -0.1 / +0.1 rounding trick.
    """
    h0, w0 = src_hw
    r = min(imgsz[0] / h0, imgsz[1] / w0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = (imgsz[1] - new_w) / 2, (imgsz[0] - new_h) / 2
    return dict(r=r, new_w=new_w, new_h=new_h, top=int(round(dh - 0.1)), left=int(round(dw - 0.1)))


def scale_params(src_hw, imgsz=(640, 640)):
    """``(gain, pad_x, pad_y)`` of ultralytics ``scale_boxes`` (SURVEY.md §8 a, N3)."""
    h0, w0 = src_hw
    gain = min(imgsz[0] / h0, imgsz[1] / w0)
    pad_x = round((imgsz[1] - w0 * gain) / 2 - 0.1)
    pad_y = round((imgsz[0] - h0 * gain) / 2 - 0.1)
    return float(gain), float(pad_x), float(pad_y)


def synthetic_frame(rng: np.random.Generator, height: int = 1080, width: int = 1920,
                    boxes_xyxy=None) -> np.ndarray:
    """A u8 HWC BGR frame: seeded noise with filled rectangles (config 3 input)."""
    img = rng.integers(0, 256, (height, width, 3), dtype=np.uint8)
    if boxes_xyxy is not None:
        for k, (x1, y1, x2, y2) in enumerate(np.asarray(boxes_xyxy)):
            img[max(int(y1), 0):int(y2), max(int(x1), 0):int(x2)] = (37 * k + 11) % 256
    return img


# ---------------------------------------------------------------------------
# Vectorised planting (bench-sized workloads: configs 3 / 4 of BASELINE.json)
# ---------------------------------------------------------------------------
def pingpong_objects(seed: int, num_frames: int, num_objects: int = 30, src_hw=(1080, 1920),
                     imgsz=(640, 640), w_range=(90.0, 240.0), h_range=(180.0, 480.0), vmax: float = 3.0,
                     dropout: float = 0.05, logit_range=(-0.5, 3.0), distractor_frac: float = 0.1):
    """Movers of one stream over a cycle of ``num_frames`` frames that can be replayed forever:
    each object runs forward for half the cycle and back again, so frame ``F-1 -> 0`` is as
    continuous as any other step (config-1 dynamics scaled to 1080p).

    Returns ``dict(boxes (F, M, 4) letterbox px, cls (M,), logit (F, M), present (F, M))``.
    """
    rng = np.random.default_rng(seed)
    h0, w0 = src_hw
    m = num_objects
    wh = np.stack([rng.uniform(*w_range, m), rng.uniform(*h_range, m)], 1)
    vel = rng.uniform(-vmax, vmax, (m, 2))
    reach = np.abs(vel) * (num_frames // 2)
    lim = np.array([w0, h0], np.float64)
    lo = wh / 2 + np.maximum(-vel, 0) * (num_frames // 2)
    hi = lim - wh / 2 - np.maximum(vel, 0) * (num_frames // 2)
    c0 = rng.uniform(lo, np.maximum(hi, lo + 1e-3))
    del reach
    tri = np.minimum(np.arange(num_frames), num_frames - np.arange(num_frames)).astype(np.float64)
    c = c0[None] + vel[None] * tri[:, None, None]                       # (F, M, 2) source px
    gain, pad_x, pad_y = scale_params(src_hw, imgsz)
    half = wh[None] / 2
    boxes = np.concatenate([c - half, c + half], -1) * gain
    boxes[..., [0, 2]] += pad_x
    boxes[..., [1, 3]] += pad_y
    cls = rng.choice(np.asarray(WANTED_CLASSES, np.int32), m).astype(np.int32)
    unwanted = np.setdiff1d(np.arange(NUM_CLASSES), WANTED_CLASSES)
    distract = rng.uniform(size=m) < distractor_frac
    cls[distract] = rng.choice(unwanted, int(distract.sum()))
    logit = rng.uniform(*logit_range, (num_frames, m))
    present = rng.uniform(size=(num_frames, m)) >= dropout
    return dict(boxes=boxes, cls=cls, logit=logit, present=present)


def plant_cells(boxes, cls, logit, owner, rng: np.random.Generator, imgsz=(640, 640), peak: float = 6.0):
    """Which head cells to overwrite for a flat list of M objects (vectorised :func:`plant_head`).

    ``owner[m]`` is the index of the (frame, stream) head tensor object m belongs to.  Returns one
    dict per level with int64 ``owner, gy, gx, cls`` (K,), float32 ``logit`` (K,), int64 ``lo`` (K, 4)
    and float32 ``vlo, vhi`` (K, 4): class channel ``64 + cls`` gets ``logit``; DFL bins ``lo`` /
    ``lo + 1`` of side k get ``vlo`` / ``vhi``.  A cell claimed by two objects goes to the first.
    """
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    x1, y1, x2, y2 = (boxes[:, k] for k in range(4))
    cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
    dy, dx = (a.reshape(-1) for a in np.meshgrid([-1, 0, 1], [-1, 0, 1], indexing="ij"))
    out = []
    for s, (h, w) in zip(STRIDES, head_shapes(imgsz)):
        gx = np.floor(cx / s).astype(np.int64)[:, None] + dx[None]
        gy = np.floor(cy / s).astype(np.int64)[:, None] + dy[None]
        ax, ay = (gx + 0.5) * s, (gy + 0.5) * s
        dist = np.stack([ax - x1[:, None], ay - y1[:, None], x2[:, None] - ax, y2[:, None] - ay], -1) / s
        ok = (gx >= 0) & (gx < w) & (gy >= 0) & (gy < h) & (dist.min(-1) >= 0) & (dist.max(-1) <= REG_MAX - 1.001)
        m_idx, c_idx = np.nonzero(ok)
        own, ggy, ggx = np.asarray(owner, np.int64)[m_idx], gy[ok], gx[ok]
        key = (own * h + ggy) * w + ggx
        _, first = np.unique(key, return_index=True)
        first.sort()
        m_idx, c_idx, own, ggy, ggx = m_idx[first], c_idx[first], own[first], ggy[first], ggx[first]
        d = dist[ok][first]
        lo = np.floor(d).astype(np.int64)
        fr = d - lo
        jitter = np.where(c_idx == 4, 0.0, rng.uniform(0.0, 1.0, len(m_idx)))     # centre cell scores best
        out.append(dict(owner=own, gy=ggy, gx=ggx, cls=np.asarray(cls, np.int64)[m_idx],
                        logit=(np.asarray(logit, np.float64)[m_idx] - jitter).astype(np.float32), lo=lo,
                        vlo=(peak + np.log(np.maximum(1.0 - fr, 1e-3))).astype(np.float32),
                        vhi=(peak + np.log(np.maximum(fr, 1e-3))).astype(np.float32)))
    return out


def scatter_cells(head, cells: dict) -> None:
    """Write planted cells into ``head`` (N, 144, h, w): a NumPy array or a torch tensor (any device)."""
    n, ch, h, w = head.shape
    own, gy, gx = cells["owner"], cells["gy"], cells["gx"]
    base = (own * ch * h + gy) * w + gx                                  # channel 0 of the cell
    plane = h * w
    idx = [base + (4 * REG_MAX + cells["cls"]) * plane]
    val = [cells["logit"]]
    for k in range(4):
        idx += [base + (k * REG_MAX + cells["lo"][:, k]) * plane, base + (k * REG_MAX + cells["lo"][:, k] + 1) * plane]
        val += [cells["vlo"][:, k], cells["vhi"][:, k]]
    idx, val = np.concatenate(idx), np.concatenate(val)
    if isinstance(head, np.ndarray):
        head.reshape(-1)[idx] = val.astype(head.dtype)
    else:
        import torch
        flat = head.view(-1)
        flat[torch.from_numpy(idx).to(head.device)] = torch.from_numpy(val).to(head.device).to(head.dtype)
