"""Oracle: CPU restatement of the detector's pre- and post-processing.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PARITY UNPINNED: the arithmetic lives in
``ultralytics>=8.1.0`` (requirements.txt:14), reached from
``/root/reference/src/detection/detector.py:100-111`` and not installable here; the reference
has no test that fixes a number at this boundary.  What follows restates ultralytics' published
8.1-era algorithm with the same third-party kernels ultralytics itself calls:

  * ``letterbox``          <- ``ultralytics.data.augment.LetterBox.__call__`` (auto=False,
                              scaleup=True, center=True): ``cv2.resize(INTER_LINEAR)`` +
                              ``cv2.copyMakeBorder(114)``
  * ``preprocess``         <- ``BasePredictor.preprocess``: BGR->RGB, HWC->CHW, cast, /255
  * ``resize_fixedpoint``  -  NumPy model of OpenCV's 8-bit INTER_LINEAR (what the CUDA kernel
                              implements); checked against ``cv2.resize`` in the tests
  * ``decode_head``        <- ``Detect._inference`` + ``DFL.forward`` + ``dist2bbox``
  * ``non_max_suppression``<- ``ultralytics.utils.ops.non_max_suppression`` (multi_label=False,
                              max_nms=30000, max_wh=7680) calling ``torchvision.ops.nms``
  * ``scale_boxes``        <- ``ultralytics.utils.ops.scale_boxes`` + ``clip_boxes``
  * ``nms_plain``          -  loop restatement of torchvision's CPU ``nms`` kernel, used to
                              cross-check torchvision itself on ties / thresholds
"""

from __future__ import annotations

import numpy as np

STRIDES = (8, 16, 32)
REG_MAX = 16
MAX_WH = 7680
MAX_NMS = 30000


# ---------------------------------------------------------------------------
# P1 letterbox / preprocess
# ---------------------------------------------------------------------------
def letterbox_geometry(src_hw, new_shape=(640, 640), auto=False, stride=32):
    """LetterBox.__call__ geometry.  ``auto=True`` (what ultralytics uses for .pt models) pads only
    to the next multiple of ``stride``: ``dw, dh = np.mod(dw, stride), np.mod(dh, stride)``."""
    h0, w0 = src_hw
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    new_unpad = int(round(w0 * r)), int(round(h0 * r))          # (w, h)
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, (top, bottom, left, right)


def letterbox(img: np.ndarray, new_shape=(640, 640), auto=False) -> np.ndarray:
    """u8 HWC BGR -> u8 HWC BGR letterboxed, via OpenCV exactly as ultralytics does."""
    import cv2
    new_unpad, (top, bottom, left, right) = letterbox_geometry(img.shape[:2], new_shape, auto)
    if img.shape[:2][::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def preprocess(img_lb: np.ndarray, dtype="bf16"):
    """Letterboxed u8 HWC BGR -> (3, H, W) torch tensor / 255 in ``dtype`` (f32 | f16 | bf16)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(img_lb[..., ::-1].transpose(2, 0, 1)))
    t = t.to({"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[dtype])
    return t / 255


def _taps(dst: int, src: int, clamp_like_x: bool):
    """OpenCV's per-destination-index source offset and 11-bit fixed-point weights."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if clamp_like_x:                       # horizontal rule: reset the fraction at the borders
        lo = s < 0
        s[lo], f[lo] = 0, 0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0
        s0, s1 = s, np.minimum(s + 1, src - 1)
    else:                                  # vertical rule: keep the weights, clip the rows
        s0, s1 = np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1)
    c0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    c1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s0, s1, c0, c1


def resize_fixedpoint(img: np.ndarray, new_wh) -> np.ndarray:
    """NumPy model of cv2.resize(u8, INTER_LINEAR)."""
    nw, nh = new_wh
    h, w = img.shape[:2]
    x0, x1, a0, a1 = _taps(nw, w, True)
    y0, y1, b0, b1 = _taps(nh, h, False)
    src = img.astype(np.int64)
    hor = src[:, x0] * a0[None, :, None] + src[:, x1] * a1[None, :, None]        # (h, nw, 3)
    s0, s1 = hor[y0], hor[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------
# D1 head decode
# ---------------------------------------------------------------------------
def make_anchors(shapes):
    """``utils.tal.make_anchors`` (offset 0.5): anchor points (2, A) and strides (1, A)."""
    import torch
    pts, strides = [], []
    for (h, w), s in zip(shapes, STRIDES):
        sx = torch.arange(w, dtype=torch.float32) + 0.5
        sy = torch.arange(h, dtype=torch.float32) + 0.5
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        strides.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts).transpose(0, 1), torch.cat(strides).transpose(0, 1)


def decode_head(heads, nc: int = 80):
    """[(B, 64+nc, h, w)] x 3 -> (B, 4+nc, A) float32: xywh in letterbox px + class probs."""
    import torch
    heads = [torch.as_tensor(h).float() for h in heads]
    b = heads[0].shape[0]
    shapes = [tuple(h.shape[2:]) for h in heads]
    x_cat = torch.cat([h.reshape(b, 4 * REG_MAX + nc, -1) for h in heads], 2)
    box, cls = x_cat.split((4 * REG_MAX, nc), 1)
    a = box.shape[2]
    prob = box.view(b, 4, REG_MAX, a).transpose(2, 1).softmax(1)                  # DFL
    dist = (prob * torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)).sum(1)
    anchors, strides = make_anchors(shapes)
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
    return torch.cat((dbox, cls.sigmoid()), 1)


# ---------------------------------------------------------------------------
# N1 + N2 non_max_suppression
# ---------------------------------------------------------------------------
def xywh2xyxy(x):
    import torch
    y = torch.empty_like(x)
    dw, dh = x[..., 2] / 2, x[..., 3] / 2
    y[..., 0], y[..., 1] = x[..., 0] - dw, x[..., 1] - dh
    y[..., 2], y[..., 3] = x[..., 0] + dw, x[..., 1] + dh
    return y


def non_max_suppression(pred, conf_thres=0.35, iou_thres=0.45, classes=None, agnostic=False,
                        max_det=100, nms=None):
    """Per image: ``(dets (n, 6) = xyxy, conf, cls ; keep idx (n,) ; anchor idx (n,))``."""
    import torch
    import torchvision
    nms = nms or torchvision.ops.nms
    pred = torch.as_tensor(pred).float()
    nc = pred.shape[1] - 4
    xc = pred[:, 4:4 + nc].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    out = []
    for xi, x in enumerate(pred):
        anchor = torch.arange(x.shape[0])[xc[xi]]
        x = x[xc[xi]]
        box, cls = x.split((4, nc), 1)
        conf, j = cls.max(1, keepdim=True)
        keep = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float()), 1)[keep]
        anchor = anchor[keep]
        if classes is not None:
            sel = (x[:, 5:6] == torch.tensor(classes)).any(1)
            x, anchor = x[sel], anchor[sel]
        n = x.shape[0]
        if not n:
            out.append((torch.zeros((0, 6)), torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)))
            continue
        if n > MAX_NMS:
            top = x[:, 4].argsort(descending=True)[:MAX_NMS]
            x, anchor = x[top], anchor[top]
        c = x[:, 5:6] * (0 if agnostic else MAX_WH)
        i = nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out.append((x[i], i, anchor[i]))
    return out


def nms_plain(boxes, scores, iou_threshold: float):
    """torchvision's CPU kernel as a loop: stable descending order, float32 IoU > double thr."""
    import torch
    b = np.asarray(boxes, np.float32)
    s = np.asarray(scores, np.float32)
    order = np.argsort(-s, kind="stable")
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    dead = np.zeros(len(s), bool)
    keep = []
    for _i, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        w = np.maximum(np.float32(0), np.minimum(b[i, 2], b[rest, 2]) - np.maximum(b[i, 0], b[rest, 0]))
        h = np.maximum(np.float32(0), np.minimum(b[i, 3], b[rest, 3]) - np.maximum(b[i, 1], b[rest, 1]))
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            ovr = inter / (area[i] + area[rest] - inter)
        dead[rest[ovr.astype(np.float64) > iou_threshold]] = True
    return torch.as_tensor(np.array(keep, np.int64))


# ---------------------------------------------------------------------------
# N3 scale_boxes
# ---------------------------------------------------------------------------
def scale_boxes(img1_hw, boxes, img0_hw):
    """Letterbox px -> source px, clipped (``ops.scale_boxes`` with padding=True)."""
    import torch
    boxes = torch.as_tensor(boxes).clone()
    gain = min(img1_hw[0] / img0_hw[0], img1_hw[1] / img0_hw[1])
    pad = (round((img1_hw[1] - img0_hw[1] * gain) / 2 - 0.1),
           round((img1_hw[0] - img0_hw[0] * gain) / 2 - 0.1))
    boxes[..., [0, 2]] -= pad[0]
    boxes[..., [1, 3]] -= pad[1]
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, img0_hw[1])
    boxes[..., 1].clamp_(0, img0_hw[0])
    boxes[..., 2].clamp_(0, img0_hw[1])
    boxes[..., 3].clamp_(0, img0_hw[0])
    return boxes


def detect_post(heads, src_hw, imgsz=(640, 640), conf=0.35, iou=0.45, classes=None, agnostic=False,
                max_det=100, nc=80):
    """decode -> NMS -> rescale for a batch; returns per image dict(xyxy, conf, cls, keep, anchor)."""
    pred = decode_head(heads, nc)
    res = []
    for dets, keep, anchor in non_max_suppression(pred, conf, iou, classes, agnostic, max_det):
        xyxy = scale_boxes(imgsz, dets[:, :4], src_hw) if len(dets) else dets[:, :4]
        res.append(dict(xyxy=xyxy.numpy().astype(np.float32), conf=dets[:, 4].numpy().astype(np.float32),
                        cls=dets[:, 5].numpy().astype(np.int32), keep=keep.numpy(), anchor=anchor.numpy()))
    return res
