"""Drop-in for ``src/detection/detector.py``: same ``Detector`` / ``Detections`` API, with
letterbox, head decode, NMS and rescale on the GPU (``rtm_letterbox`` + ``rtm_decode_nms``)
around an unchanged PyTorch conv forward.

What differs from the reference, and why:

* the reference delegates everything to ``ultralytics.YOLO.predict`` (detector.py:100-111);
  ultralytics is not available offline, so the network is the same-shape stand-in of
  :mod:`.yolov8s`.  ``model_path`` may point to a ``state_dict`` saved from that module;
  ``model=`` accepts any ``nn.Module`` returning the three raw head tensors.  A missing file
  raises ``FileNotFoundError`` exactly like detector.py:90.
* ``half=True`` runs the network in bfloat16 (BASELINE.json's north_star: uint8 -> bf16
  letterbox); decode / NMS arithmetic is float32 either way.
* ``detect_batch`` handles a batch of equally sized frames in one launch sequence.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from .. import _lib
from ..synth import scale_params


@dataclass
class Detections:
    """Structured detection output for a single frame (detector.py:29-48)."""
    xyxy: np.ndarray
    confidence: np.ndarray
    class_id: np.ndarray
    class_names: list = field(default_factory=list)

    def __len__(self) -> int:
        return len(self.confidence)

    def filter_classes(self, keep: Sequence[int]) -> "Detections":
        mask = np.isin(self.class_id, keep)
        return Detections(xyxy=self.xyxy[mask], confidence=self.confidence[mask], class_id=self.class_id[mask],
                          class_names=[n for n, m in zip(self.class_names, mask) if m])


class Detector:
    """YOLOv8 detector with a B200-native pre/post-process (detector.py:54-135)."""

    _WARMUP_ITERATIONS = 10

    def __init__(self, model_path: Optional[str], fallback_model: Optional[str] = None,
                 input_size: tuple = (640, 640), confidence: float = 0.35, iou: float = 0.45,
                 classes: Optional[list] = None, half: bool = True, device: str = "cuda:0",
                 max_det: int = 100, agnostic_nms: bool = False, *, model=None, names=None,
                 num_classes: int = 80, warmup: bool = True, auto: Optional[bool] = None) -> None:
        import torch
        from .yolov8s import COCO_NAMES, YOLOv8s
        self.input_size = tuple(input_size)
        # auto=True: LetterBox pads only to the next multiple of 32 (what ultralytics does for .pt models:
        # 1080p -> 384 x 640, 5040 anchors); False: a fixed square input (exported engines; the BASELINE configs:
        # 8400 anchors).  None (default) follows ultralytics' own rule, `auto = model is a .pt file`
        # (detector.py:84-87 loads either an engine or the .pt fallback); a module passed in is taken as fixed-shape.
        self.auto = auto
        self.confidence, self.iou, self.classes = confidence, iou, classes
        self.half = half and torch.cuda.is_available()
        self.device, self.max_det, self.agnostic_nms = device, max_det, agnostic_nms
        self._lib = _lib.lib()
        self._dev = torch.device(device)
        self._dtype = torch.bfloat16 if self.half else torch.float32
        self.names = dict(enumerate(names if names is not None else COCO_NAMES))
        self.nc = int(num_classes)

        if model is not None:
            net = model
        else:
            primary = Path(model_path) if model_path else None
            if primary is not None and primary.exists():
                chosen = primary
            elif fallback_model and Path(fallback_model).exists():
                chosen = Path(fallback_model)
            else:
                raise FileNotFoundError(f"No model found at {model_path} or {fallback_model}")
            if self.auto is None:
                self.auto = chosen.suffix == ".pt"
            net = YOLOv8s(self.nc)
            state = torch.load(str(chosen), map_location="cpu", weights_only=True)
            net.load_state_dict(state)
        self.auto = bool(self.auto)
        self.model = net.to(self._dev).to(self._dtype).eval()
        self._params = _lib.make_nms_params(confidence, iou, max_det, agnostic_nms, classes, self.nc)
        self._bufs = {}
        if warmup:
            self._warmup()

    # ------------------------------------------------------------------
    def _buffers(self, B: int, src_hw):
        import torch
        key = (B, tuple(src_hw))
        if key not in self._bufs:
            (H, W), geom = self._geometry(src_hw)
            A = sum((H // s) * (W // s) for s in (8, 16, 32))
            dev = self._dev
            gain, px, py = scale_params(src_hw, (H, W))
            self._bufs = {key: dict(
                hw=(H, W), geom=geom,
                frames=torch.empty((B, src_hw[0], src_hw[1], 3), dtype=torch.uint8, device=dev),
                net_in=torch.empty((B, 3, H, W), dtype=self._dtype, device=dev),
                scale=torch.tensor([[gain, px, py, src_hw[1], src_hw[0]]] * B, dtype=torch.float32, device=dev),
                xyxy=torch.zeros((B, self.max_det, 4), dtype=torch.float32, device=dev),
                conf=torch.zeros((B, self.max_det), dtype=torch.float32, device=dev),
                cls=torch.zeros((B, self.max_det), dtype=torch.int32, device=dev),
                count=torch.zeros(B, dtype=torch.int32, device=dev),
                status=torch.zeros(B, dtype=torch.int32, device=dev),
                ws=torch.zeros(self._lib.rtm_nms_workspace_bytes(B, A), dtype=torch.uint8, device=dev))}
            self._lib.rtm_workspace_release(self._bufs[key]["ws"].data_ptr())   # a recycled address starts afresh
        return self._bufs[key]

    def _geometry(self, src_hw):
        """LetterBox.__call__ (scaleup=True, center=True): network input (H, W) and where the resized
        source goes in it: (new_h, new_w, top, left)."""
        h0, w0 = src_hw
        Hf = Wf = int(self.input_size[0])          # detector.py:102 passes imgsz=self.input_size[0]: a square target
        r = min(Hf / h0, Wf / w0)
        new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
        dw, dh = Wf - new_w, Hf - new_h
        if self.auto:
            dw, dh = dw % 32, dh % 32
        H, W = new_h + dh, new_w + dw
        top, left = int(round(dh / 2 - 0.1)), int(round(dw / 2 - 0.1))
        return (H, W), (new_h, new_w, top, left)

    def detect(self, frame: np.ndarray) -> Detections:
        """Run inference on a single BGR frame and return ``Detections`` (detector.py:98-112)."""
        return self.detect_batch(frame[None])[0]

    def detect_batch(self, frames: np.ndarray) -> list:
        """``frames``: (B, H, W, 3) uint8 BGR, one frame per stream."""
        import torch
        frames = np.ascontiguousarray(frames)
        B, h0, w0, _ = frames.shape
        buf = self._buffers(B, (h0, w0))
        (H, W), (new_h, new_w, top, left) = buf["hw"], buf["geom"]
        with torch.cuda.device(self._dev), torch.inference_mode():
            st = _lib.cuda_stream()
            buf["frames"].copy_(torch.from_numpy(frames), non_blocking=True)
            _lib.check(self._lib.rtm_letterbox_ex(buf["frames"].data_ptr(), B, h0, w0, w0 * 3, h0 * w0 * 3,
                                                  buf["net_in"].data_ptr(), _lib.dtype_code(self._dtype), H, W,
                                                  new_h, new_w, top, left, st))
            heads = [t.contiguous() for t in self.model(buf["net_in"])]
            _lib.check(self._lib.rtm_decode_nms(
                heads[0].data_ptr(), heads[1].data_ptr(), heads[2].data_ptr(), _lib.dtype_code(heads[0].dtype),
                B, H, W, C.byref(self._params), buf["scale"].data_ptr(), buf["xyxy"].data_ptr(),
                buf["conf"].data_ptr(), buf["cls"].data_ptr(), None, None, buf["count"].data_ptr(),
                self.max_det, buf["status"].data_ptr(), buf["ws"].data_ptr(), buf["ws"].numel(), st))
            n = buf["count"].cpu().numpy()
            status = buf["status"].cpu().numpy()
            if status.any():
                buf["status"].zero_()
            _lib.raise_on_status(status, "Detector")
            xyxy, conf, cls = buf["xyxy"].cpu().numpy(), buf["conf"].cpu().numpy(), buf["cls"].cpu().numpy()
        return [self._parse(xyxy[b, :n[b]], conf[b, :n[b]], cls[b, :n[b]]) for b in range(B)]

    def _parse(self, xyxy, conf, cls) -> Detections:
        """detector.py:117-129."""
        if len(conf) == 0:
            return Detections(xyxy=np.empty((0, 4), dtype=np.float32), confidence=np.empty(0, dtype=np.float32),
                              class_id=np.empty(0, dtype=np.int32))
        names = [self.names.get(int(c), str(c)) for c in cls]
        return Detections(xyxy=xyxy.astype(np.float32), confidence=conf.astype(np.float32),
                          class_id=cls.astype(np.int32), class_names=names)

    def _warmup(self) -> None:
        dummy = np.zeros((*self.input_size[::-1], 3), dtype=np.uint8)   # detector.py:132
        for _ in range(self._WARMUP_ITERATIONS):
            self.detect(dummy)
