"""Same-shape PyTorch stand-in for the YOLOv8s network the reference loads through
``ultralytics.YOLO`` (detector.py:84): identical layer graph, channel widths and head layout
("YOLOv8s module"), so the three head tensors have the shapes, dtype
and channel order the post-backbone kernels consume.  ``ultralytics`` itself is not
installable offline; weights are random-initialised (there is no network for checkpoints),
which is all the post-backbone benchmark needs - the conv forward is not a parity subject.

The module returns the RAW head tensors ``[(B, 64 + nc, H/8, W/8), (.., H/16, ..), (.., H/32, ..)]``
(what ``Detect.forward`` concatenates per level before decoding); decoding is
``rtm_decode_nms``'s job.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn


class Conv(nn.Module):
    """Conv2d(no bias) + BatchNorm(eps 1e-3, momentum 0.03) + SiLU."""

    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True):
        super().__init__()
        self.cv1 = Conv(c1, c2, 3, 1)
        self.cv2 = Conv(c2, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m(x)
        y2 = self.m(y1)
        return self.cv2(torch.cat((x, y1, y2, self.m(y2)), 1))


class DetectHead(nn.Module):
    """YOLOv8 Detect: per level a box branch (4 x reg_max DFL logits) and a class branch."""

    reg_max = 16

    def __init__(self, nc=80, ch=(128, 256, 512), strides=(8, 16, 32)):
        super().__init__()
        self.nc, self.strides = nc, strides
        c2 = max(16, ch[0] // 4, self.reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        for a, b, s in zip(self.cv2, self.cv3, strides):           # Detect.bias_init
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:nc] = math.log(5 / nc / (640 / s) ** 2)

    def forward(self, feats):
        return [torch.cat((self.cv2[i](x), self.cv3[i](x)), 1).contiguous() for i, x in enumerate(feats)]


class YOLOv8s(nn.Module):
    """depth 0.33 / width 0.50 instance of the YOLOv8 graph (about 11.2 M parameters)."""

    def __init__(self, nc: int = 80):
        super().__init__()
        self.b0, self.b1 = Conv(3, 32, 3, 2), Conv(32, 64, 3, 2)
        self.b2 = C2f(64, 64, 1, True)
        self.b3, self.b4 = Conv(64, 128, 3, 2), C2f(128, 128, 2, True)
        self.b5, self.b6 = Conv(128, 256, 3, 2), C2f(256, 256, 2, True)
        self.b7, self.b8 = Conv(256, 512, 3, 2), C2f(512, 512, 1, True)
        self.b9 = SPPF(512, 512, 5)
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.h12 = C2f(768, 256, 1)
        self.h15 = C2f(384, 128, 1)
        self.h16, self.h18 = Conv(128, 128, 3, 2), C2f(384, 256, 1)
        self.h19, self.h21 = Conv(256, 256, 3, 2), C2f(768, 512, 1)
        self.detect = DetectHead(nc)

    def forward(self, x):
        p3 = self.b4(self.b3(self.b2(self.b1(self.b0(x)))))
        p4 = self.b6(self.b5(p3))
        p5 = self.b9(self.b8(self.b7(p4)))
        n4 = self.h12(torch.cat((self.up(p5), p4), 1))
        n3 = self.h15(torch.cat((self.up(n4), p3), 1))
        o4 = self.h18(torch.cat((self.h16(n3), n4), 1))
        o5 = self.h21(torch.cat((self.h19(o4), p5), 1))
        return self.detect([n3, o4, o5])


#: COCO class names (what ``result.names`` holds for the stock yolov8s checkpoint, detector.py:128)
COCO_NAMES = (
    "person", "bicycle", "car", "motorcycle", "airplane", "bus", "train", "truck", "boat", "traffic light",
    "fire hydrant", "stop sign", "parking meter", "bench", "bird", "cat", "dog", "horse", "sheep", "cow",
    "elephant", "bear", "zebra", "giraffe", "backpack", "umbrella", "handbag", "tie", "suitcase", "frisbee",
    "skis", "snowboard", "sports ball", "kite", "baseball bat", "baseball glove", "skateboard", "surfboard",
    "tennis racket", "bottle", "wine glass", "cup", "fork", "knife", "spoon", "bowl", "banana", "apple",
    "sandwich", "orange", "broccoli", "carrot", "hot dog", "pizza", "donut", "cake", "chair", "couch",
    "potted plant", "bed", "dining table", "toilet", "tv", "laptop", "mouse", "remote", "keyboard",
    "cell phone", "microwave", "oven", "toaster", "sink", "refrigerator", "book", "clock", "vase",
    "scissors", "teddy bear", "hair drier", "toothbrush",
)
