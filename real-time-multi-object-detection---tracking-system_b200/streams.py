"""Batched, device-resident state and the fused post-backbone step for B video streams.

This is the product path: one process drives one B200; every stream of the batch owns a
row-slab of the device tables below (the reference keeps the same state in Python lists and
dicts, one tracker and one zone engine per stream - tracker.py:55-56, zone_engine.py:72-75).

    DeviceTrackTable   `_tracks` + `_next_id` of B trackers           (rtm_track_table)
    ZoneTables         zone polygons + dwell / cooldown state of B engines
                       (rtm_zone_set, rtm_zone_state)
    StreamBatch        decode + NMS -> tracker -> zones for B streams per call
                       (rtm_post_backbone_step / rtm_post_backbone_step_host)

PyTorch provides device memory and the CUDA stream; all arithmetic is in librtmodt_b200.so.
"""

from __future__ import annotations

import ctypes as C
import time
from dataclasses import asdict, dataclass, field
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib
from .synth import scale_params


# ---------------------------------------------------------------------------
# Event record (same schema as the reference's ZoneEvent, zone_engine.py:29-45)
# ---------------------------------------------------------------------------
@dataclass
class ZoneEvent:
    """Immutable event record written to the alert log."""
    timestamp_utc: str
    event_type: str
    zone_name: str
    track_id: int
    class_id: int
    class_name: str
    dwell_time_sec: float
    bbox_xyxy: list
    centroid: list
    frame_id: int
    metadata: dict = field(default_factory=dict)

    def to_json(self) -> str:
        import json
        return json.dumps(asdict(self), default=str)


@dataclass
class Zone:
    """Parsed zone config (zone_engine.py:51-58)."""
    name: str
    polygon: np.ndarray
    trigger: str
    dwell_time_sec: float = 2.0
    cooldown_sec: float = 10.0
    direction: Optional[str] = None


def parse_zone(cfg: dict) -> Zone:
    """zone_engine.py:142-151 - KeyError on a missing ``name`` / ``polygon``."""
    pts = np.array(cfg["polygon"], dtype=np.int32)
    return Zone(name=cfg["name"], polygon=pts, trigger=cfg.get("trigger", "intrusion"),
                dwell_time_sec=cfg.get("dwell_time_sec", 2.0), cooldown_sec=cfg.get("cooldown_sec", 10.0),
                direction=cfg.get("direction"))


# ---------------------------------------------------------------------------
# Track table
# ---------------------------------------------------------------------------
class DeviceTrackTable:
    """SoA track table of B streams on the device (rtm_track_table)."""

    FIELDS = ("track_id", "xyxy", "confidence", "class_id", "age", "time_since_update")

    def __init__(self, num_streams: int, capacity: int, device, kalman: bool = False) -> None:
        import torch
        self.num_streams, self.capacity, self.device = int(num_streams), int(capacity), device
        i32 = dict(dtype=torch.int32, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        B, cap = self.num_streams, self.capacity
        self.count = torch.zeros(B, **i32)
        self.next_id = torch.ones(B, **i32)                      # tracker.py:55
        self.track_id = torch.zeros(B, cap, **i32)
        self.xyxy = torch.zeros(B, cap, 4, **f32)
        self.confidence = torch.zeros(B, cap, **f32)
        self.class_id = torch.zeros(B, cap, **i32)
        self.age = torch.zeros(B, cap, **i32)
        self.time_since_update = torch.zeros(B, cap, **i32)
        self.struct = _lib.TrackTable(B, cap, self.count.data_ptr(), self.next_id.data_ptr(),
                                      self.track_id.data_ptr(), self.xyxy.data_ptr(),
                                      self.confidence.data_ptr(), self.class_id.data_ptr(),
                                      self.age.data_ptr(), self.time_since_update.data_ptr())
        # opt-in motion model (rtm_kalman_state): four (position, velocity) filters per track
        self.kf_mean = self.kf_cov = self.kalman = None
        if kalman:
            self.kf_mean = torch.zeros(B, cap, 8, **f32)
            self.kf_cov = torch.zeros(B, cap, 12, **f32)
            self.kalman = _lib.KalmanState(self.kf_mean.data_ptr(), self.kf_cov.data_ptr())

    def grown(self, capacity: int) -> "DeviceTrackTable":
        """A table of ``capacity`` >= this one's rows per stream holding the same rows and counters."""
        big = DeviceTrackTable(self.num_streams, max(int(capacity), self.capacity), self.device, kalman=self.kalman is not None)
        big.count.copy_(self.count)
        big.next_id.copy_(self.next_id)
        for k in self.FIELDS + (("kf_mean", "kf_cov") if self.kalman is not None else ()):
            getattr(big, k)[:, :self.capacity] = getattr(self, k)
        return big

    def to_host(self):
        """dict of host numpy arrays (synchronises)."""
        out = {k: getattr(self, k).cpu().numpy() for k in self.FIELDS}
        out["count"] = self.count.cpu().numpy()
        out["next_id"] = self.next_id.cpu().numpy()
        if self.kalman is not None:
            out["kf_mean"] = self.kf_mean.cpu().numpy()
            out["kf_cov"] = self.kf_cov.cpu().numpy()
        return out

    def load(self, stream: int, rows: dict, next_id: int) -> None:
        """Overwrite one stream's rows from host arrays (state import)."""
        import torch
        n = len(rows["track_id"])
        if n > self.capacity:
            raise ValueError(f"{n} tracks > capacity {self.capacity}")
        for k in self.FIELDS:
            t = getattr(self, k)
            t[stream, :n] = torch.as_tensor(np.asarray(rows[k]), dtype=t.dtype).to(self.device)
        self.count[stream] = n
        self.next_id[stream] = int(next_id)


def stream_tracks(host: dict, b: int):
    """One stream's live rows of ``DeviceTrackTable.to_host()`` as the reference's list of dicts
    (``tracker._core._tracks``, tracker.py:127-134)."""
    n = int(host["count"][b])
    return [dict(track_id=int(host["track_id"][b, r]), xyxy=host["xyxy"][b, r].copy(),
                 confidence=float(host["confidence"][b, r]), class_id=int(host["class_id"][b, r]),
                 age=int(host["age"][b, r]), time_since_update=int(host["time_since_update"][b, r]))
            for r in range(n)]


# ---------------------------------------------------------------------------
# Zones
# ---------------------------------------------------------------------------
class ZoneTables:
    """Zone polygons (read-only) and per-(track row, zone name) dwell / cooldown state."""

    def __init__(self, zones_per_stream: Sequence[Sequence[dict]], capacity: int, device,
                 max_events: Optional[int] = None) -> None:
        import torch
        self.device = device
        self.capacity = int(capacity)
        self.zones = [[parse_zone(z) for z in zs] for zs in zones_per_stream]
        B = self.num_streams = len(self.zones)
        zone_off, poly_off, pts, dwell, cool, col = [0], [0], [], [], [], []
        ncol = 1
        for zs in self.zones:
            if len(zs) > _lib.MAX_ZONES_PER_STREAM:
                raise ValueError(f"at most {_lib.MAX_ZONES_PER_STREAM} zones per stream")
            names: dict[str, int] = {}
            for z in zs:
                poly = np.asarray(z.polygon, np.int32).reshape(-1, 2)
                pts.append(poly)
                poly_off.append(poly_off[-1] + len(poly))
                dwell.append(float(z.dwell_time_sec))
                cool.append(float(z.cooldown_sec))
                col.append(names.setdefault(z.name, len(names)))   # state is keyed by NAME
            ncol = max(ncol, len(names))
            zone_off.append(zone_off[-1] + len(zs))
        self.num_columns = ncol
        self.max_zones = max([len(z) for z in self.zones] + [1])
        dev = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt).to(device).contiguous()
        self._zone_off = dev(zone_off, torch.int32)
        self._poly_off = dev(poly_off, torch.int32)
        self._poly_xy = dev(np.concatenate(pts) if pts else np.zeros((1, 2), np.int32), torch.int32)
        self._dwell = dev(dwell or [0.0], torch.float64)
        self._cool = dev(cool or [0.0], torch.float64)
        self._col = dev(col or [0], torch.int32)
        self.zone_set = _lib.ZoneSet(B, ncol, self._zone_off.data_ptr(), self._poly_off.data_ptr(),
                                     self._poly_xy.data_ptr(), self._dwell.data_ptr(),
                                     self._cool.data_ptr(), self._col.data_ptr())
        # state ping-pong: (B, columns, capacity) f64; NaN = not inside, 0.0 = never alerted
        self._state = []
        for _ in range(2):
            fs = torch.full((B, ncol, self.capacity), float("nan"), dtype=torch.float64, device=device)
            la = torch.zeros((B, ncol, self.capacity), dtype=torch.float64, device=device)
            self._state.append((fs, la, _lib.ZoneState(fs.data_ptr(), la.data_ptr())))
        self.cur = 0
        # every (track row, zone) pair of a stream may fire in one step - a crowd whose dwell timers run out on the same
        # frame does just that - up to what the kernel can stage per stream and step (kZoneStepEvents)
        self.event_stride = int(max_events or max(16, min(self.capacity * self.max_zones, _lib.ZONE_STEP_EVENTS)))
        # event slabs alternate with the state tables: the set written by a step is not the one the caller may
        # still be reading from the step before (rtm_step_io.results_alternate)
        self._events = [torch.zeros((B, self.event_stride, 64), dtype=torch.uint8, device=device) for _ in range(2)]
        self._event_count = [torch.zeros(B, dtype=torch.int32, device=device) for _ in range(2)]

    @property
    def events(self):
        """Event slabs of the last step (the set that goes with the current state)."""
        return self._events[self.cur]

    @property
    def event_count(self):
        return self._event_count[self.cur]

    def state_in(self):
        return self._state[self.cur]

    def state_out(self):
        return self._state[self.cur ^ 1]

    def swap(self) -> None:
        self.cur ^= 1

    def decode_events(self, ev_host: np.ndarray, count_host: np.ndarray, class_names=None):
        """Host bytes of rtm_zone_event -> per-stream lists of ZoneEvent, reference order."""
        recs = ev_host.view(np.dtype(_lib.EVENT_DTYPE)).reshape(self.num_streams, -1)
        stamp = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())      # zone_engine.py:108
        out = []
        for b in range(self.num_streams):
            evs = []
            for r in recs[b, :int(count_host[b])]:
                z = self.zones[b][int(r["zone"])]
                cid = int(r["class_id"])
                name = ""
                if class_names is not None:
                    name = class_names.get(cid, "") if isinstance(class_names, dict) else class_names(int(r["track_id"]))
                evs.append(ZoneEvent(timestamp_utc=stamp, event_type=z.trigger, zone_name=z.name,
                                     track_id=int(r["track_id"]), class_id=cid, class_name=name,
                                     dwell_time_sec=round(float(r["dwell"]), 2),
                                     bbox_xyxy=[float(v) for v in r["xyxy"]],
                                     centroid=[int(r["cx"]), int(r["cy"])], frame_id=int(r["frame_id"])))
            out.append(evs)
        return out


# ---------------------------------------------------------------------------
# The fused step
# ---------------------------------------------------------------------------
class StreamBatch:
    """decode + NMS -> tracker -> zones for B streams per call, state resident in HBM.

    Parameters mirror the reference's three constructors: detector (detector.py:59-71),
    tracker (tracker.py:46-51) and zone engine (zone_engine.py:67).
    """

    def __init__(self, num_streams: int, zones_per_stream: Optional[Sequence[Sequence[dict]]] = None,
                 src_hw=(1080, 1920), imgsz=(640, 640), num_classes: int = 80, confidence: float = 0.35,
                 iou: float = 0.45, classes: Optional[Sequence[int]] = None, max_det: int = 100,
                 agnostic_nms: bool = False, track_thresh: float = 0.5, track_buffer: int = 30,
                 match_thresh: float = 0.8, max_tracks: int = 1024, max_events: Optional[int] = None,
                 det_slots: Optional[int] = None, device="cuda:0", use_kalman: bool = False,
                 assignment: str = "greedy", max_pairs: Optional[int] = None) -> None:
        import torch
        self.lib = _lib.lib()
        self.device = torch.device(device)
        self.B = B = int(num_streams)
        self.imgsz, self.src_hw, self.nc = tuple(imgsz), tuple(src_hw), int(num_classes)
        self.params = _lib.make_nms_params(confidence, iou, max_det, agnostic_nms, classes, num_classes)
        self.track_thresh, self.match_thresh, self.track_buffer = float(track_thresh), float(match_thresh), int(track_buffer)
        self.det_stride = max(int(max_det), int(det_slots or 0))   # detection slots per stream
        self.num_anchors = sum((self.imgsz[0] // s) * (self.imgsz[1] // s) for s in (8, 16, 32))
        with torch.cuda.device(self.device):
            i32 = dict(dtype=torch.int32, device=self.device)
            f32 = dict(dtype=torch.float32, device=self.device)
            D = self.det_stride
            # two sets of result buffers, alternating like the track tables: a step never overwrites what the
            # caller may still be reading from the step before it (rtm_step_io.results_alternate)
            self._res = [dict(det_xyxy=torch.zeros(B, D, 4, **f32), det_conf=torch.zeros(B, D, **f32),
                              det_cls=torch.zeros(B, D, **i32), det_anchor=torch.zeros(B, D, **i32),
                              det_keep=torch.zeros(B, D, **i32), det_count=torch.zeros(B, **i32),
                              det_track_id=torch.zeros(B, D, **i32), det_kind=torch.zeros(B, D, **i32)) for _ in range(2)]
            self.status = torch.zeros(B, **i32)
            gain, px, py = scale_params(self.src_hw, self.imgsz)
            self.scale = torch.tensor([[gain, px, py, self.src_hw[1], self.src_hw[0]]] * B, **f32)
            ws_bytes = self.lib.rtm_nms_workspace_bytes(B, self.num_anchors)
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
            # the allocator may hand out an address the library has seen before: start its bookkeeping afresh
            _lib.check(self.lib.rtm_workspace_release(self.workspace.data_ptr()))
            # use_kalman: opt-in motion model the reference does not have (rtm_track_step_ex); off = reference
            self.use_kalman = bool(use_kalman)
            # assignment: "greedy" = what the reference runs without `lap` (and here); "lapjv" = its lap branch
            self.assignment = assignment
            _lib.track_options(0.5, 0.8, 30, assignment)          # validates the name
            # "lapjv": scratch of the general solver (stages with many admissible pairs or large conflict components);
            # max_pairs = admissible (track, detection) pairs a stage of one stream may hold
            self.assign_scratch = None
            if assignment == "lapjv":
                pairs = int(max_pairs if max_pairs is not None else 16 * max(self.det_stride, 64))
                n = self.lib.rtm_assign_scratch_bytes(B, int(max_tracks), self.det_stride, pairs)
                self.assign_scratch = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self.tables = [DeviceTrackTable(B, max_tracks, self.device, kalman=self.use_kalman) for _ in range(2)]
            self.src_row = torch.zeros(B, max_tracks, **i32)
            self.zones = None
            if zones_per_stream is not None:
                if len(zones_per_stream) != B:
                    raise ValueError("zones_per_stream must have one entry per stream")
                self.zones = ZoneTables(zones_per_stream, max_tracks, self.device, max_events)
        self.cur = 0
        self.frame_id = 0
        self._host = None
        self._io_cache, self._head_cache = {}, {}

    def close(self) -> None:
        """Give the library's per-workspace bookkeeping (its stream and events) back; the batch is unusable after."""
        ws, self.workspace = getattr(self, "workspace", None), None
        if ws is not None:
            import torch
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                self.lib.rtm_workspace_release(ws.data_ptr())

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state views -------------------------------------------------------
    def __getattr__(self, name):
        # det_xyxy, det_conf, ... : the result buffers of the last step (the set that goes with the current table)
        if name.startswith("det_") and "_res" in self.__dict__ and name in self._res[0]:
            return self._res[self.cur][name]
        raise AttributeError(name)

    @property
    def table(self) -> DeviceTrackTable:
        """The table holding the current state (after the last step)."""
        return self.tables[self.cur]

    def _io(self, heads, now: float, frame_id: int) -> _lib.StepIO:
        """rtm_step_io of this step.  Everything that does not change from step to step is filled in once
        per table parity and reused (a step is ~40 us of GPU work: the host side has to stay well below)."""
        key = (self.cur, self.zones.cur if self.zones is not None else 0)
        io = self._io_cache.get(key)
        if io is None:
            io = self._io_cache[key] = self._build_io()
        if heads is not None:
            hkey = (id(heads[0]), id(heads[1]), id(heads[2]))
            ent = self._head_cache.get(hkey)
            if ent is None or ent[0][0] is not heads[0] or ent[0][1] is not heads[1] or ent[0][2] is not heads[2]:
                for t, s in zip(heads, (8, 16, 32)):
                    exp = (self.B, 64 + self.nc, self.imgsz[0] // s, self.imgsz[1] // s)
                    if tuple(t.shape) != exp or not t.is_contiguous() or t.device != self.device:
                        raise ValueError(f"head level stride {s}: expected contiguous {exp} on {self.device}, got {tuple(t.shape)}")
                if len(self._head_cache) > 256:
                    self._head_cache.clear()
                ent = self._head_cache[hkey] = (tuple(heads), _lib.dtype_code(heads[0].dtype))   # keeps the tensors (and their ids) alive
            p3, p4, p5 = ent[0]
            io.head_p3, io.head_p4, io.head_p5 = p3.data_ptr(), p4.data_ptr(), p5.data_ptr()
            io.head_dtype = ent[1]
        io.now, io.frame_id = float(now), int(frame_id)
        return io

    def _build_io(self) -> _lib.StepIO:
        io = _lib.StepIO()
        io.img_h, io.img_w = self.imgsz
        io.scale = self.scale.data_ptr()
        res = self._res[self.cur ^ 1]                           # the set that goes with table_out
        io.det_xyxy, io.det_conf, io.det_cls = res["det_xyxy"].data_ptr(), res["det_conf"].data_ptr(), res["det_cls"].data_ptr()
        io.det_anchor, io.det_keep, io.det_count = res["det_anchor"].data_ptr(), res["det_keep"].data_ptr(), res["det_count"].data_ptr()
        io.det_stride = self.det_stride
        io.workspace, io.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        io.table_in = C.pointer(self.tables[self.cur].struct)
        io.table_out = C.pointer(self.tables[self.cur ^ 1].struct)
        io.track_thresh, io.match_thresh, io.track_buffer = self.track_thresh, self.match_thresh, self.track_buffer
        io.det_track_id, io.det_kind, io.src_row = res["det_track_id"].data_ptr(), res["det_kind"].data_ptr(), self.src_row.data_ptr()
        if self.zones is not None:
            io.zones = C.pointer(self.zones.zone_set)
            io.state_in = C.pointer(self.zones.state_in()[2])
            io.state_out = C.pointer(self.zones.state_out()[2])
            io.events, io.event_stride = self.zones._events[self.zones.cur ^ 1].data_ptr(), self.zones.event_stride
            io.event_count = self.zones._event_count[self.zones.cur ^ 1].data_ptr()
        io.status = self.status.data_ptr()
        if self.use_kalman:
            io.kalman_in = C.pointer(self.tables[self.cur].kalman)
            io.kalman_out = C.pointer(self.tables[self.cur ^ 1].kalman)
        if self.assignment == "lapjv":
            io.assignment, io.cost_limit = _lib.ASSIGN_OPTIMAL, 1 - self.match_thresh      # tracker.py:170
            io.assign_scratch, io.assign_scratch_bytes = self.assign_scratch.data_ptr(), self.assign_scratch.numel()
        return io

    def _advance(self) -> None:
        self.cur ^= 1
        if self.zones is not None:
            self.zones.swap()
        self.frame_id += 1

    # -- stepping ----------------------------------------------------------
    def step(self, heads, now: Optional[float] = None, frame_id: Optional[int] = None, heads_ready=None) -> None:
        """One frame of every stream from device-resident head tensors (asynchronous).

        ``heads_ready``: None (default) - the head tensors are ordered on the current stream like any
        other input.  ``True`` - they are complete already; a ``torch.cuda.Event`` - they are complete
        once it has fired (e.g. the backbone runs on its own stream).  In both cases the head scan is
        enqueued on a stream of the library's own (rtm_step_io.scan_async), so that consecutive scans
        run back to back instead of queueing behind the previous step's post kernel; results are
        ordered on the current stream exactly as in the default mode."""
        import torch
        now = time.time() if now is None else now                     # zone_engine.py:84
        fid = self.frame_id if frame_id is None else frame_id
        io = self._io(heads, now, fid)
        if heads_ready is None:
            io.scan_async, io.heads_ready_event, io.results_alternate = 0, None, 0
        else:
            io.scan_async, io.results_alternate = 1, 1           # the result buffers alternate with the tables
            io.heads_ready_event = None if heads_ready is True else heads_ready.cuda_event
        if torch.cuda.current_device() == self.device.index:
            rc = self.lib.rtm_post_backbone_step(C.byref(io), C.byref(self.params), _lib.cuda_stream(self.device.index))
        else:
            with torch.cuda.device(self.device):
                rc = self.lib.rtm_post_backbone_step(C.byref(io), C.byref(self.params), _lib.cuda_stream(self.device.index))
        if rc:
            _lib.check(rc)
        self._advance()

    def track_only(self, det_xyxy, det_conf, det_cls, det_count, now: Optional[float] = None,
                   frame_id: Optional[int] = None) -> None:
        """Tracker + zones on scripted detections already on the device (config 5 of
        BASELINE.json feeds 1000 boxes per stream, which the detector cannot emit)."""
        import torch
        now = time.time() if now is None else now
        fid = self.frame_id if frame_id is None else frame_id
        S = det_conf.shape[1]
        with torch.cuda.device(self.device):
            st = _lib.cuda_stream()
            tin, tout = self.tables[self.cur], self.tables[self.cur ^ 1]
            res = self._res[self.cur ^ 1]
            want_assign = S == self.det_stride
            opt = _lib.track_options(self.track_thresh, self.match_thresh, self.track_buffer, self.assignment,
                                     tin.kalman if self.use_kalman else None, tout.kalman if self.use_kalman else None,
                                     self.assign_scratch)
            _lib.check(self.lib.rtm_track_step_ex(
                C.byref(tin.struct), C.byref(tout.struct), det_xyxy.data_ptr(), det_conf.data_ptr(),
                det_cls.data_ptr(), det_count.data_ptr(), S, C.byref(opt),
                res["det_track_id"].data_ptr() if want_assign else None,
                res["det_kind"].data_ptr() if want_assign else None, self.src_row.data_ptr(),
                self.status.data_ptr(), st))
            if self.zones is not None:
                z = self.zones
                _lib.check(self.lib.rtm_zone_step(
                    C.byref(z.zone_set), C.byref(tout.struct), self.src_row.data_ptr(),
                    C.byref(z.state_in()[2]), C.byref(z.state_out()[2]), float(now), None, int(fid),
                    z._events[z.cur ^ 1].data_ptr(), z.event_stride, z._event_count[z.cur ^ 1].data_ptr(),
                    self.status.data_ptr(), st))
        self._advance()

    # -- checkpoint / resume -------------------------------------------------
    def _state_args(self):
        t = self.tables[self.cur]
        z = self.zones
        return (C.byref(t.struct), C.byref(z.state_in()[2]) if z is not None else None, z.num_columns if z is not None else 0,
                C.byref(t.kalman) if self.use_kalman else None)

    def export_state(self) -> bytes:
        """The batch's whole state - track tables, next ids, zone dwell / cooldown state, Kalman state - as one blob
        (rtm_state_export; synchronises).  The reference has no counterpart: its state lives in Python lists and dicts."""
        import torch
        z = self.zones
        n = self.lib.rtm_state_bytes(self.B, self.tables[0].capacity, z.num_columns if z is not None else 0, int(z is not None),
                                     int(self.use_kalman))
        blob = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)              # steps in flight on the library's stream included
            _lib.check(self.lib.rtm_state_export(*self._state_args(), blob.data_ptr(), n, _lib.cuda_stream()))
        return blob.numpy().tobytes() + int(self.frame_id).to_bytes(8, "little", signed=True)

    def import_state(self, blob: bytes) -> None:
        """Resume from :meth:`export_state` of a batch of the same shape (streams, max_tracks, zone columns, motion model)."""
        import torch
        n = len(blob) - 8
        buf = torch.from_numpy(np.frombuffer(blob, np.uint8, n).copy())
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _lib.check(self.lib.rtm_state_import(*self._state_args(), buf.data_ptr(), n, _lib.cuda_stream()))
        self.frame_id = int.from_bytes(blob[n:], "little", signed=True)

    # -- results -----------------------------------------------------------
    def check_status(self) -> None:
        st = self.status.cpu().numpy()
        if st.any():
            self.status.zero_()                       # reported once: later steps start from a clean word
        _lib.raise_on_status(st, "StreamBatch")

    def read_events(self, class_names=None):
        """Per-stream lists of ZoneEvent of the last step (synchronises the device)."""
        if self.zones is None:
            return [[] for _ in range(self.B)]
        self.check_status()
        cnt = self.zones.event_count.cpu().numpy()
        ev = self.zones.events[:, :max(int(cnt.max()), 1)].cpu().numpy()     # only the filled part of the slabs
        return self.zones.decode_events(ev, cnt, class_names)

    def event_records(self) -> np.ndarray:
        """The raw 64-byte ``rtm_zone_event`` records of the last step, all streams concatenated in stream
        order (a NumPy structured array of ``_lib.EVENT_DTYPE``; ``stream`` numbers the batch's streams)."""
        if self.zones is None:
            return np.zeros(0, np.dtype(_lib.EVENT_DTYPE))
        self.check_status()
        cnt = self.zones.event_count.cpu().numpy()
        ev = self.zones.events[:, :max(int(cnt.max()), 1)].cpu().numpy()
        recs = ev.view(np.dtype(_lib.EVENT_DTYPE)).reshape(self.B, -1)
        return np.concatenate([recs[b, :int(cnt[b])] for b in range(self.B)]) if self.B else recs.reshape(-1)

    def write_events(self, log_path, class_names=None, stream_names=None) -> int:
        """The event sink of the batch: the last step's events of ALL streams as JSONL, the reference's
        ``ZoneEvent.to_json()`` line per event (zone_engine.py:44-45), in (stream, track, zone) order, with ONE
        append for the whole step (the reference opens the file once per event, zone_engine.py:153-155).
        ``stream_names``: optional per-stream value stored under ``metadata["stream"]`` (the reference's
        ``metadata`` is ``{}``: omitted when None, so that a single-stream log is byte-identical to the
        reference's apart from ``timestamp_utc``).  Returns the number of lines written."""
        from pathlib import Path
        lines = []
        for b, evs in enumerate(self.read_events(class_names)):
            for e in evs:
                if stream_names is not None:
                    e.metadata = {"stream": stream_names[b]}
                lines.append(e.to_json() + "\n")
        if lines:
            path = Path(log_path)
            path.parent.mkdir(parents=True, exist_ok=True)
            with open(path, "a") as f:
                f.write("".join(lines))
        return len(lines)

    def read_detections(self):
        """Per-stream dict(xyxy, confidence, class_id, anchor, keep, track_id, kind) of the last step."""
        self.check_status()
        n = self.det_count.cpu().numpy()
        arrs = {k: getattr(self, "det_" + k).cpu().numpy() for k in ("xyxy", "conf", "cls", "anchor", "keep", "track_id", "kind")}
        return [dict(xyxy=arrs["xyxy"][b, :n[b]], confidence=arrs["conf"][b, :n[b]], class_id=arrs["cls"][b, :n[b]],
                     anchor=arrs["anchor"][b, :n[b]], keep=arrs["keep"][b, :n[b]],
                     track_id=arrs["track_id"][b, :n[b]], kind=arrs["kind"][b, :n[b]]) for b in range(self.B)]

    def read_tracks(self):
        """Per-stream ``_core._tracks``-style lists (synchronises)."""
        self.check_status()
        host = self.table.to_host()
        return [stream_tracks(host, b) for b in range(self.B)], host["next_id"].astype(np.int64)


# ---------------------------------------------------------------------------
# Host-fed stepping (the end-to-end path: host buffers in, host results out)
# ---------------------------------------------------------------------------
class StepResult:
    """Results of one host-fed step; ``wait()`` blocks until its device->host copies landed."""

    def __init__(self, feeder: "HostFeeder", slot: dict, index: int) -> None:
        self._feeder, self._slot, self._index = feeder, slot, index

    @property
    def stale(self) -> bool:
        """The slot has been handed to a later step (``depth`` steps on): its host buffers hold that step's results."""
        return self._feeder.k > self._index + self._feeder.depth

    def wait(self) -> "StepResult":
        # (a stale result is complete - the feeder synchronised on it before reusing the slot - and the slot's event now
        # belongs to the later step: waiting on it would wait for THAT step and stall the caller's pipeline)
        if not self.stale:
            self._slot["done"].synchronize()
            _lib.raise_on_status(self._slot["status"].numpy(), "HostFeeder")
        return self

    def _fresh(self) -> None:
        if self.stale:
            raise _lib.RtmError(f"HostFeeder: the results of step {self._index} were overwritten by step "
                                f"{self._index + self._feeder.depth}; read a step's results before {self._feeder.depth} more steps are enqueued")
        self.wait()

    @property
    def det_count(self) -> np.ndarray:
        self._fresh()
        return self._slot["det_count"].numpy().copy()

    def detections(self):
        self._fresh()
        s = self._slot
        n = s["det_count"].numpy()
        return [dict(xyxy=s["det_xyxy"].numpy()[b, :n[b]].copy(), confidence=s["det_conf"].numpy()[b, :n[b]].copy(),
                     class_id=s["det_cls"].numpy()[b, :n[b]].copy(), track_id=s["det_track_id"].numpy()[b, :n[b]].copy())
                for b in range(len(n))]

    def events(self, class_names=None):
        self._fresh()
        z = self._feeder.batch.zones
        if z is None:
            return [[] for _ in range(self._feeder.batch.B)]
        cnt = self._slot["event_count"].numpy()
        if int(cnt.max()) > self._feeder.event_prefix:
            raise _lib.RtmError(f"HostFeeder: a stream emitted {int(cnt.max())} events in one step, only the first "
                                f"{self._feeder.event_prefix} are copied to the host (raise event_prefix)")
        return z.decode_events(self._slot["events"].numpy(), cnt, class_names)


class HostFeeder:
    """Feeds a :class:`StreamBatch` from HOST head tensors through ``rtm_post_backbone_step_host``.

    Two slots, each with its own CUDA stream, pinned staging buffers and device head buffers:
    the host->device copy of step k+1 overlaps the kernels and the device->host copy of step k
    (the kernels of consecutive steps stay ordered through events because they share the track
    and zone tables).  Every step's inputs cross PCIe inside the step, and its results
    (detections with track ids, events, status) come back to pinned host memory.
    """

    def __init__(self, batch: StreamBatch, head_dtype, depth: int = 2, event_prefix: int = 64) -> None:
        import torch
        self.batch, self.depth = batch, int(depth)
        # events copied back per stream and step; a stream that emits more in one step raises (nothing is dropped silently)
        self.event_prefix = min(int(event_prefix), batch.zones.event_stride) if batch.zones is not None else 1
        self.lib = batch.lib
        B, D, dev = batch.B, batch.det_stride, batch.device
        shapes = [(B, 64 + batch.nc, batch.imgsz[0] // s, batch.imgsz[1] // s) for s in (8, 16, 32)]
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        self._shapes, self._head_dtype = shapes, head_dtype

        def views(flat):                                     # the three levels back to back in one buffer
            out, o = [], 0
            for sh in shapes:
                n = int(np.prod(sh))
                out.append(flat[o:o + n].view(sh))
                o += n
            return out

        self._views = views
        total = sum(int(np.prod(sh)) for sh in shapes)
        self.slots = []
        with torch.cuda.device(dev):
            for _ in range(self.depth):
                stream = torch.cuda.Stream(device=dev)
                done, copied = torch.cuda.Event(), torch.cuda.Event()
                done.record(stream)                          # materialise the cudaEvent_t handles
                copied.record(stream)
                ev_stride = self.event_prefix
                self.slots.append(dict(
                    stream=stream, done=done, copied=copied,
                    host_heads=views(pin((total,), head_dtype)),
                    dev_heads=views(torch.empty(total, dtype=head_dtype, device=dev)),
                    events=pin((B, ev_stride, 64), torch.uint8), event_count=pin((B,), torch.int32).zero_(),
                    det_xyxy=pin((B, D, 4), torch.float32), det_conf=pin((B, D), torch.float32),
                    det_cls=pin((B, D), torch.int32), det_track_id=pin((B, D), torch.int32),
                    det_count=pin((B,), torch.int32).zero_(), status=pin((B,), torch.int32).zero_()))
            torch.cuda.synchronize(dev)
        self.k = 0
        self.h2d_bytes = sum(int(np.prod(s)) for s in shapes) * torch.empty(0, dtype=head_dtype).element_size()
        s0 = self.slots[0]
        self.d2h_bytes = sum(s0[k].numel() * s0[k].element_size() for k in
                             ("events", "event_count", "det_xyxy", "det_conf", "det_cls", "det_track_id", "det_count", "status"))

    def alloc_pinned_heads(self):
        """Three page-locked host tensors (one per level) laid out back to back, for :meth:`step_pinned`:
        contiguous levels cross PCIe as a single transfer."""
        import torch
        total = sum(int(np.prod(sh)) for sh in self._shapes)
        return self._views(torch.empty((total,), dtype=self._head_dtype, pin_memory=True))

    def stage(self, slot_index: int, heads_host) -> None:
        """Copy host head tensors into the pinned staging buffers of a slot (plain memcpy)."""
        for dst, src in zip(self.slots[slot_index]["host_heads"], heads_host):
            dst.copy_(src)

    def step(self, heads_host=None, now: Optional[float] = None, frame_id: Optional[int] = None) -> StepResult:
        """Enqueue one step.  ``heads_host``: three host tensors (copied into the slot's pinned
        buffers first) or None when they were filled with :meth:`stage`."""
        slot_index = self.k % self.depth
        self.slots[slot_index]["done"].synchronize()          # the slot's previous results were consumed
        if heads_host is not None:
            self.stage(slot_index, heads_host)
        return self._enqueue(self.slots[slot_index]["host_heads"], now, frame_id)

    def step_pinned(self, pinned_heads, now: Optional[float] = None, frame_id: Optional[int] = None) -> StepResult:
        """Enqueue one step whose inputs already sit in page-locked host tensors (no staging copy)."""
        for t in pinned_heads:
            if not t.is_pinned():
                raise ValueError("step_pinned needs page-locked (pinned) host tensors")
        self.slots[self.k % self.depth]["done"].synchronize()
        return self._enqueue(pinned_heads, now, frame_id)

    def _enqueue(self, host_heads, now, frame_id) -> StepResult:
        import torch
        b = self.batch
        slot = self.slots[self.k % self.depth]
        prev = self.slots[(self.k - 1) % self.depth]
        now = time.time() if now is None else now
        fid = b.frame_id if frame_id is None else frame_id
        io = b._io(None, now, fid)
        io.scan_async, io.heads_ready_event, io.results_alternate = 0, None, 0   # the copy and the kernels share the slot's stream
        io.head_p3, io.head_p4, io.head_p5 = (t.data_ptr() for t in slot["dev_heads"])
        io.head_dtype = _lib.dtype_code(slot["dev_heads"][0].dtype)
        h = _lib.StepHostIO()
        h.host_head_p3, h.host_head_p4, h.host_head_p5 = (t.data_ptr() for t in host_heads)
        if b.zones is not None:
            h.host_events, h.host_event_count = slot["events"].data_ptr(), slot["event_count"].data_ptr()
        h.host_det_xyxy, h.host_det_conf = slot["det_xyxy"].data_ptr(), slot["det_conf"].data_ptr()
        h.host_det_cls, h.host_det_track_id = slot["det_cls"].data_ptr(), slot["det_track_id"].data_ptr()
        h.host_det_count, h.host_status = slot["det_count"].data_ptr(), slot["status"].data_ptr()
        h.host_event_stride = self.event_prefix
        h.wait_event = prev["done"].cuda_event if self.k > 0 and self.depth > 1 else None
        h.done_event = slot["done"].cuda_event
        # the copies of consecutive steps follow each other on the link (side by side they end side by side and the link
        # idles while both streams run their kernels)
        h.copy_wait_event = prev["copied"].cuda_event if self.k > 0 and self.depth > 1 else None
        h.copy_done_event = slot["copied"].cuda_event
        with torch.cuda.device(b.device):
            _lib.check(self.lib.rtm_post_backbone_step_host(C.byref(io), C.byref(h), C.byref(b.params),
                                                            slot["stream"].cuda_stream))
        b._advance()
        self.k += 1
        return StepResult(self, slot, self.k - 1)
