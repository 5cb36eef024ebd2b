"""The benchmark workload of BASELINE.json configs 3 / 4: S synthetic 1080p streams, a cycle of F
frames of planted YOLOv8 head tensors per stream (8400 anchors x (64 + 80) channels) and 4 zones
per stream.  Every stream's tensors are a function of its GLOBAL stream id and the frame only:
planted cells come from NumPy generators seeded per (stream, frame), the background noise from a
counter-based integer hash evaluated with exact integer arithmetic (the same bits on the CPU and on
any GPU) - so a stream's head tensor does not depend on how streams are sharded over ranks, nor on
which arm of the bench generates it ("identical head tensors").
Used by bench.py (both arms) and by the tests; touches neither the CUDA library nor the oracle.
"""

from __future__ import annotations

import numpy as np

from . import synth


def hashed_noise(stream_ids, frame: int, level: int, numel: int, device):
    """(len(stream_ids), numel) float32 noise, mean 0, std 1.155 (sum of the four bytes of a 32-bit hash of
    (stream, frame, level, element), centred and divided by 128): integer arithmetic below 2^63 and an exact
    conversion, hence bit-identical on the CPU and on any GPU."""
    import torch
    idx = torch.arange(numel, dtype=torch.int64, device=device)[None]
    seed = torch.tensor([(sid * 1_000_003 + frame * 7_919 + level * 104_729 + 12_345) & 0x7FFFFFFF for sid in stream_ids],
                        dtype=torch.int64, device=device)[:, None]
    x = (idx * 2_654_435_761 + seed) & 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45D9F3B) & 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45D9F3B) & 0xFFFFFFFF
    x = (x >> 16) ^ x
    s = (x & 255) + ((x >> 8) & 255) + ((x >> 16) & 255) + (x >> 24)
    return (s - 510).to(torch.float32) / 128.0


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
        cells = None
        for s, (sid, o) in enumerate(zip(self.stream_ids, self.objects)):
            keep = o["present"][f]
            n = int(keep.sum())
            rng = np.random.default_rng([77, sid, f])                      # per (global stream, frame)
            c = synth.plant_cells(o["boxes"][f][keep], o["cls"][keep], o["logit"][f][keep],
                                  np.full(n, s, np.int64), rng, self.imgsz)
            cells = c if cells is None else [{k: np.concatenate([a[k], b[k]]) for k in a} for a, b in zip(cells, c)]
        out = []
        for level, ((h, w), c) in enumerate(zip(synth.head_shapes(self.imgsz), cells)):
            t = torch.empty((self.S, synth.NUM_OUT, h, w), device=self.device, dtype=torch.float32)
            for s0 in range(0, self.S, 16):
                ids = self.stream_ids[s0:s0 + 16]
                t[s0:s0 + len(ids)] = hashed_noise(ids, f, level, synth.NUM_OUT * h * w, self.device).view(len(ids), synth.NUM_OUT, h, w)
            t[:, 4 * synth.REG_MAX:] -= 6.0                               # background class logits ~ (-6, 1.15)
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
