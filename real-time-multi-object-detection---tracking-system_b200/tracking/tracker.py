"""Drop-in for ``src/tracking/tracker.py``: same classes, arguments, errors and quirks,
with the association arithmetic on the GPU (``rtm_track_step``).

Quirks kept on purpose, because parity is with the reference's behaviour, not with canonical
ByteTrack:

* ``MultiObjectTracker.update`` returns ``[]`` on every frame: the reference ages every track
  *after* resetting the matched ones and then filters on ``time_since_update == 0``
  (tracker.py:138-147).  The tracks matched or born in the last frame are available from
  :meth:`MultiObjectTracker.active_tracks` (``time_since_update == 1``).
* no Kalman filter, IoU >= ``match_thresh`` in both stages, class-agnostic matching,
  empty frames never prune (tracker.py:70-73).
* the state the reference keeps in ``tracker._core._tracks`` / ``_next_id`` is readable here
  under the same names (copied from the device on access).

One instance tracks one stream (B = 1 slab of :class:`~..streams.DeviceTrackTable`); the
batched product path is :class:`~..streams.StreamBatch`.
"""

from __future__ import annotations

import ctypes as C
from collections import defaultdict
from dataclasses import dataclass, field

import numpy as np

from .. import _lib
from ..streams import DeviceTrackTable, stream_tracks


@dataclass
class Track:
    """Represents a single tracked object (tracker.py:27-37)."""
    track_id: int
    xyxy: np.ndarray
    confidence: float
    class_id: int
    class_name: str = ""
    age: int = 0
    time_since_update: int = 0
    trail: list = field(default_factory=list)


class _ByteTrackCore:
    """Device-resident counterpart of the reference's ``_ByteTrackCore`` (tracker.py:43-194)."""

    def __init__(self, track_thresh: float = 0.5, track_buffer: int = 30, match_thresh: float = 0.8,
                 max_tracks: int = 1024, max_dets: int = 256, device="cuda:0", use_kalman: bool = False,
                 assignment: str = "greedy") -> None:
        import torch
        self.track_thresh, self.track_buffer, self.match_thresh = track_thresh, track_buffer, match_thresh
        # opt-in: associate against the box a constant-velocity filter predicts (the reference has no
        # motion model; default off keeps its behaviour bit for bit)
        self.use_kalman = bool(use_kalman)
        # "greedy": the branch of tracker.py:163-194 the reference takes when `lap` is missing (as here);
        # "lapjv": its lap.lapjv branch, solved exactly on the GPU
        self.assignment = assignment
        self.auto_grow = True              # False: a full table raises RtmError instead (fixed memory)
        _lib.track_options(track_thresh, match_thresh, track_buffer, assignment)
        self._lib = _lib.lib()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self._tables = [DeviceTrackTable(1, max_tracks, self.device, kalman=self.use_kalman) for _ in range(2)]
            self._cur = 0
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._alloc_slots(int(max_dets))

    def _alloc_slots(self, max_dets: int) -> None:
        """Detection slots and everything sized by them or by the table capacity."""
        import torch
        cap, dev = self._tables[0].capacity, self.device
        self.max_dets = max_dets
        self._det_xyxy = torch.zeros(1, max_dets, 4, dtype=torch.float32, device=dev)
        self._det_conf = torch.zeros(1, max_dets, dtype=torch.float32, device=dev)
        self._det_cls = torch.zeros(1, max_dets, dtype=torch.int32, device=dev)
        self._det_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self._det_tid = torch.zeros(1, max_dets, dtype=torch.int32, device=dev)
        self._det_kind = torch.zeros(1, max_dets, dtype=torch.int32, device=dev)
        self._src_row = torch.zeros(1, cap, dtype=torch.int32, device=dev)
        self._assign_scratch = None                          # "lapjv": scratch of the general solver (crowds, low thresholds)
        if self.assignment == "lapjv":
            n = self._lib.rtm_assign_scratch_bytes(1, cap, max_dets, 64 * max_dets)
            self._assign_scratch = torch.zeros(n, dtype=torch.uint8, device=dev)

    def _grow(self, max_tracks: int, max_dets: int) -> None:
        """The reference's lists have no capacity (tracker.py:55, 127-134): the tables double when a frame does not
        fit.  The step that reported the overflow wrote the OTHER table only, so it is simply run again."""
        import torch
        with torch.cuda.device(self.device):
            if max_tracks > self._tables[0].capacity:
                self._tables = [t.grown(max_tracks) for t in self._tables]
            self._alloc_slots(max(max_dets, self.max_dets))

    # -- the reference's attribute names ------------------------------------
    @property
    def _tracks(self) -> list:
        return stream_tracks(self._tables[self._cur].to_host(), 0)

    @property
    def _next_id(self) -> int:
        return int(self._tables[self._cur].next_id.cpu()[0])

    def update(self, xyxy: np.ndarray, confidence: np.ndarray, class_id: np.ndarray) -> list:
        """One tracking step (tracker.py:58-141).  Returns the reference's (empty) list."""
        import torch
        n = len(confidence)
        if n > self.max_dets:
            self._grow(self._tables[0].capacity, 1 << (n - 1).bit_length())
        while True:
            with torch.cuda.device(self.device):
                if n:
                    self._det_xyxy[0, :n] = torch.as_tensor(np.ascontiguousarray(xyxy, np.float32).reshape(n, 4)).to(self.device)
                    self._det_conf[0, :n] = torch.as_tensor(np.ascontiguousarray(confidence, np.float32)).to(self.device)
                    self._det_cls[0, :n] = torch.as_tensor(np.ascontiguousarray(class_id, np.int32)).to(self.device)
                self._det_count.fill_(n)
                tin, tout = self._tables[self._cur], self._tables[self._cur ^ 1]
                opt = _lib.track_options(self.track_thresh, self.match_thresh, self.track_buffer, self.assignment,
                                         tin.kalman if self.use_kalman else None, tout.kalman if self.use_kalman else None,
                                         self._assign_scratch)
                _lib.check(self._lib.rtm_track_step_ex(
                    C.byref(tin.struct), C.byref(tout.struct), self._det_xyxy.data_ptr(),
                    self._det_conf.data_ptr(), self._det_cls.data_ptr(), self._det_count.data_ptr(),
                    self.max_dets, C.byref(opt), self._det_tid.data_ptr(), self._det_kind.data_ptr(),
                    self._src_row.data_ptr(), self._status.data_ptr(), _lib.cuda_stream()))
                st = self._status.cpu().numpy()
                if st.any():
                    self._status.zero_()
            if (int(st[0]) & _lib.STATUS_TRACK_OVERFLOW) and self.auto_grow:
                self._grow(2 * self._tables[0].capacity, self.max_dets)      # tin is untouched: run the frame again
                continue
            break
        self._cur ^= 1
        self._last_n = n
        _lib.raise_on_status(st, "MultiObjectTracker")
        # the reference filters on time_since_update == 0 AFTER ageing every track, so nothing
        # ever qualifies (tracker.py:141, 144-147)
        return []

    def assignments(self):
        """(track_id, kind) per detection of the last update (RTM_DET_* kinds)."""
        n = self._last_n
        return self._det_tid.cpu().numpy()[0, :n], self._det_kind.cpu().numpy()[0, :n]


class MultiObjectTracker:
    """High-level tracker wrapping ByteTrack (tracker.py:200-259)."""

    def __init__(self, algorithm: str = "bytetrack", **kwargs) -> None:
        self.algorithm = algorithm.lower()
        if self.algorithm == "bytetrack":
            bt_params = kwargs.get("bytetrack", kwargs)           # nested or flat, tracker.py:206
            extra = {k: kwargs[k] for k in ("max_tracks", "max_dets", "device", "use_kalman", "assignment") if k in kwargs}
            self._core = _ByteTrackCore(track_thresh=bt_params.get("track_thresh", 0.5),
                                        track_buffer=bt_params.get("track_buffer", 30),
                                        match_thresh=bt_params.get("match_thresh", 0.8), **extra)
        elif self.algorithm == "deepsort":
            raise NotImplementedError("DeepSORT adapter not yet wired. Use bytetrack.")
        else:
            raise ValueError(f"Unknown tracker: {self.algorithm}")
        self._trail_map = defaultdict(list)
        self._trail_maxlen = 30
        self._frame = 0                    # update() calls so far
        self._trail_frame = {}             # track id -> frame its trail was last extended in

    def update(self, detections) -> list:
        """``detections`` needs ``.xyxy``, ``.confidence``, ``.class_id`` (tracker.py:234-238)."""
        raw = self._core.update(detections.xyxy, detections.confidence, detections.class_id)
        self._frame += 1
        return self._wrap(raw)

    def active_tracks(self) -> list:
        """Tracks matched or born in the last frame (``time_since_update == 1``), as ``Track``
        objects with trails - what a caller of the reference presumably wanted from update()."""
        return self._wrap([t for t in self._core._tracks if t["time_since_update"] == 1])

    def _wrap(self, raw) -> list:
        tracks = []
        for r in raw:                                             # tracker.py:241-258
            tid = r["track_id"]
            cx = int((r["xyxy"][0] + r["xyxy"][2]) / 2)
            cy = int((r["xyxy"][1] + r["xyxy"][3]) / 2)
            trail = self._trail_map[tid]
            if self._trail_frame.get(tid) != self._frame:         # one point per frame, however often the view is read
                self._trail_frame[tid] = self._frame
                trail.append((cx, cy))
                if len(trail) > self._trail_maxlen:
                    trail.pop(0)
            tracks.append(Track(track_id=tid, xyxy=r["xyxy"], confidence=r["confidence"],
                                class_id=r["class_id"], age=r["age"],
                                time_since_update=r["time_since_update"], trail=list(trail)))
        return tracks
