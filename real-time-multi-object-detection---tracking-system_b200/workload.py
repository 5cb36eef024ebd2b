"""The benchmark workload of BASELINE.json configs 3 / 4: S synthetic 1080p streams, a cycle of F
frames of planted YOLOv8 head tensors per stream (8400 anchors x (64 + 80) channels) and 4 zones
per stream.  Planted cells come from NumPy (seeded per stream, identical wherever they are
generated); the background noise is drawn with torch on the device that holds the tensors.
Used by bench.py (both arms) and by the tests; touches neither the CUDA library nor the oracle.
"""

from __future__ import annotations

import numpy as np

from . import synth


class PostBackboneWorkload:
    """``heads[f]`` = three tensors (S, 144, h, w) for frame f of the cycle; ``zones[s]`` = zone configs."""

    def __init__(self, num_streams: int, num_frames: int = 16, first_stream: int = 0, device="cpu",
                 dtype=None, num_objects: int = 30, num_zones: int = 4, src_hw=(1080, 1920),
                 imgsz=(640, 640), dwell_time_sec: float = 0.5, cooldown_sec: float = 2.0) -> None:
        import torch
        self.S, self.F = int(num_streams), int(num_frames)
        self.src_hw, self.imgsz = tuple(src_hw), tuple(imgsz)
        self.device = torch.device(device)
        self.dtype = dtype or torch.bfloat16
        self.stream_ids = list(range(first_stream, first_stream + self.S))
        self.zones = [synth.make_zones(seed=sid, num_zones=num_zones, width=src_hw[1], height=src_hw[0],
                                       dwell_time_sec=dwell_time_sec, cooldown_sec=cooldown_sec)
                      for sid in self.stream_ids]
        self.objects = [synth.pingpong_objects(1000 + sid, self.F, num_objects, src_hw, imgsz) for sid in self.stream_ids]
        self.heads = [self._make_frame(f) for f in range(self.F)]

    def _make_frame(self, f: int):
        import torch
        boxes, cls, logit, owner = [], [], [], []
        for s, o in enumerate(self.objects):
            keep = o["present"][f]
            boxes.append(o["boxes"][f][keep])
            cls.append(o["cls"][keep])
            logit.append(o["logit"][f][keep])
            owner.append(np.full(int(keep.sum()), s, np.int64))
        rng = np.random.default_rng(77_000 + 131 * self.stream_ids[0] + f)
        cells = synth.plant_cells(np.concatenate(boxes), np.concatenate(cls), np.concatenate(logit),
                                  np.concatenate(owner), rng, self.imgsz)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(5_000_011 * (self.stream_ids[0] + 1) + f)
        out = []
        for (h, w), c in zip(synth.head_shapes(self.imgsz), cells):
            t = torch.randn((self.S, synth.NUM_OUT, h, w), generator=gen, device=self.device, dtype=torch.float32)
            t[:, 4 * synth.REG_MAX:] -= 6.0                               # background class logits ~ N(-6, 1)
            synth.scatter_cells(t, c)
            out.append(t.to(self.dtype).contiguous())
        return out

    @property
    def bytes_per_stream_frame(self) -> int:
        import torch
        return synth.num_anchors(self.imgsz) * synth.NUM_OUT * torch.empty(0, dtype=self.dtype).element_size()

    def host_frame(self, f: int, streams=None):
        """Frame f of the cycle as float32 host tensors (for the oracle), optionally a subset of streams."""
        sel = slice(None) if streams is None else list(streams)
        return [t[sel].float().cpu() for t in self.heads[f % self.F]]
