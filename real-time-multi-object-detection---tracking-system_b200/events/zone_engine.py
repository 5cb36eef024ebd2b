"""Drop-in for ``src/events/zone_engine.py``: same class, arguments, errors and event schema,
with point-in-polygon, dwell and cooldown evaluation on the GPU (``rtm_zone_step``).

Differences that are additions, not changes: the clock is injectable (``clock=`` keyword,
default ``time.time`` as zone_engine.py:84) and JSONL lines are written in one append per call
instead of one ``open()`` per event (zone_engine.py:153-155); line content is the reference's
``ZoneEvent.to_json()``.

The reference keys its state by ``track_id`` and zone NAME in unbounded dicts; here every
track id seen is given a row of a device table (rows are never reused because the cooldown
ledger is never purged either - zone_engine.py:75), the table doubling when it fills up.
"""

from __future__ import annotations

import ctypes as C
import time
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

from .. import _lib
from ..streams import DeviceTrackTable, Zone, ZoneEvent, ZoneTables

__all__ = ["ZoneEventEngine", "ZoneEvent", "Zone"]


class ZoneEventEngine:
    """Evaluate tracks against polygon zones and emit events (zone_engine.py:64-157)."""

    def __init__(self, zone_configs: list, log_path: str = "logs/events.jsonl", *,
                 clock: Callable[[], float] = time.time, device="cuda:0", initial_rows: int = 256) -> None:
        import torch
        self._lib = _lib.lib()
        self.device = torch.device(device)
        self._zone_configs = list(zone_configs)
        self.clock = clock
        self.log_path = Path(log_path)
        self.log_path.parent.mkdir(parents=True, exist_ok=True)
        self._rows: dict[int, int] = {}           # track_id -> row of the state table
        self._present: set[int] = set()           # ids passed to the previous call
        self._build(int(initial_rows))
        self.zones = self._tables.zones[0]        # list[Zone], as the reference's attribute

    def _build(self, capacity: int, old=None) -> None:
        import torch
        with torch.cuda.device(self.device):
            tables = ZoneTables([self._zone_configs], capacity, self.device)
            if old is not None:                   # grow: carry the state rows over
                n = old.capacity
                for k in (0, 1):
                    tables._state[0][k][:, :, :n] = old.state_in()[k]
                tables.cur = 0
            self._tables = tables
            self._track = DeviceTrackTable(1, capacity, self.device)
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def process(self, tracks: Sequence, frame_id: int) -> list:
        """Check all tracks against all zones; returns new events (zone_engine.py:82-132)."""
        import torch
        now = self.clock()
        ids = [int(t.track_id) for t in tracks]
        for tid in ids:
            if tid not in self._rows:
                self._rows[tid] = len(self._rows)
        if len(self._rows) > self._tables.capacity:
            self._build(max(2 * self._tables.capacity, len(self._rows)), old=self._tables)
        cap = self._tables.capacity
        n_rows = len(self._rows)
        # host staging of the "track table": rows of present tracks get tsu = 1, rows of tracks
        # absent from this call get tsu = 2 (their dwell timers are purged, zone_engine.py:128-130)
        tsu = np.full(cap, 2, np.int32)
        xyxy = np.zeros((cap, 4), np.float32)
        tid_arr = np.zeros(cap, np.int32)
        cls = np.zeros(cap, np.int32)
        order = []
        for t in tracks:
            r = self._rows[int(t.track_id)]
            tsu[r] = 1
            xyxy[r] = np.asarray(t.xyxy, np.float32)
            tid_arr[r] = int(t.track_id)
            cls[r] = int(t.class_id)
            order.append(r)
        with torch.cuda.device(self.device):
            tt = self._track
            tt.time_since_update[0] = torch.from_numpy(tsu).to(self.device)
            tt.xyxy[0] = torch.from_numpy(xyxy).to(self.device)
            tt.track_id[0] = torch.from_numpy(tid_arr).to(self.device)
            tt.class_id[0] = torch.from_numpy(cls).to(self.device)
            tt.count.fill_(n_rows)
            z = self._tables
            st = z.state_in()[2]                  # in place: rows are persistent here
            _lib.check(self._lib.rtm_zone_step(
                C.byref(z.zone_set), C.byref(tt.struct), None, C.byref(st), C.byref(st), float(now), None,
                int(frame_id), z.events.data_ptr(), z.event_stride, z.event_count.data_ptr(),
                self._status.data_ptr(), _lib.cuda_stream()))
            _lib.raise_on_status(self._status.cpu().numpy(), "ZoneEventEngine")
            names = {int(t.track_id): getattr(t, "class_name", "") for t in tracks}
            events = z.decode_events(z.events.cpu().numpy(), z.event_count.cpu().numpy(),
                                     class_names=lambda tid: names.get(tid, ""))[0]
        # the kernel emits in (row, zone) order; the reference iterates tracks in call order
        if order != sorted(order):
            pos = {tid: k for k, tid in enumerate(ids)}
            events.sort(key=lambda e: pos[e.track_id])            # stable: zone order is kept
        self._present = set(ids)
        if events:
            self._write(events)
        return events

    def get_zone_polygons(self) -> list:
        """For visualization overlay (zone_engine.py:134-136)."""
        return [(z.name, z.polygon) for z in self.zones]

    def _write(self, events) -> None:
        with open(self.log_path, "a") as f:
            f.write("".join(e.to_json() + "\n" for e in events))
