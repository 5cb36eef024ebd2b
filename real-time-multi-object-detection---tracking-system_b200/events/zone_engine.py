"""Drop-in for ``src/events/zone_engine.py``: same class, arguments, errors and event schema,
with point-in-polygon, dwell and cooldown evaluation on the GPU (``rtm_zone_step``).

Differences that are additions, not changes: the clock is injectable (``clock=`` keyword,
default ``time.time`` as zone_engine.py:84) and JSONL lines are written in one append per call
instead of one ``open()`` per event (zone_engine.py:153-155); line content is the reference's
``ZoneEvent.to_json()``.

The reference keys its state by ``track_id`` and zone NAME in two dicts (zone_engine.py:72-75).
So does this class, on the host: per call only the tracks PASSED IN travel to the device, as a
compact table in call order with their dwell / cooldown state beside them (one upload, one kernel,
one download) - the work per frame is O(tracks of this call), like the reference's, however many
ids the stream has seen.  The batched product path (:class:`~..streams.StreamBatch`) keeps the
state resident on the device instead.
"""

from __future__ import annotations

import ctypes as C
import time
from pathlib import Path
from typing import Callable, Optional, Sequence

import numpy as np

from .. import _lib
from ..streams import DeviceTrackTable, Zone, ZoneEvent, ZoneTables

__all__ = ["ZoneEventEngine", "ZoneEvent", "Zone"]


class ZoneEventEngine:
    """Evaluate tracks against polygon zones and emit events (zone_engine.py:64-157)."""

    def __init__(self, zone_configs: list, log_path: str = "logs/events.jsonl", *,
                 clock: Callable[[], float] = time.time, device="cuda:0", initial_rows: int = 64,
                 max_events: Optional[int] = None) -> None:
        import torch
        self._lib = _lib.lib()
        self.device = torch.device(device)
        self._zone_configs = list(zone_configs)
        self.clock = clock
        self.log_path = Path(log_path)
        self.log_path.parent.mkdir(parents=True, exist_ok=True)
        # the reference's two dicts, with one float64 per state column (= distinct zone name) as value
        self._occupancy: dict[int, np.ndarray] = {}   # track_id -> first_seen per column, NaN = not inside
        self._cooldown: dict[int, np.ndarray] = {}    # track_id -> last_alert per column, 0.0 = never (never purged)
        self._max_events = max_events
        self._build(int(initial_rows))
        self.zones = self._tables.zones[0]            # list[Zone], as the reference's attribute

    def _build(self, capacity: int) -> None:
        import torch
        with torch.cuda.device(self.device):
            self._tables = ZoneTables([self._zone_configs], capacity, self.device, self._max_events)
            self._track = DeviceTrackTable(1, capacity, self.device)
            self._track.time_since_update.fill_(1)    # every row of a call is a track passed to process()
            self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        z = self._tables
        # rows of one kernel call: never more than the event buffer can hold if every (row, zone) pair fires
        self._rows_per_call = max(1, min(capacity, z.event_stride // max(z.max_zones, 1)))

    def process(self, tracks: Sequence, frame_id: int) -> list:
        """Check all tracks against all zones; returns new events (zone_engine.py:82-132)."""
        now = self.clock()
        tracks = list(tracks)
        events: list = []
        # The kernel evaluates the rows of a call side by side.  The reference walks them one after the other,
        # which only matters when an id occurs twice in one call: such a call is cut where an id repeats.
        start, seen = 0, set()
        for k, t in enumerate(tracks):
            tid = int(t.track_id)
            if tid in seen or k - start == self._rows_per_call:
                events += self._process_rows(tracks[start:k], frame_id, now)
                start, seen = k, set()
            seen.add(tid)
        if start < len(tracks):
            events += self._process_rows(tracks[start:], frame_id, now)
        # tracks absent from this call lose their dwell timers, never their cooldowns (zone_engine.py:128-130)
        active = {int(t.track_id) for t in tracks}
        for tid in [tid for tid in self._occupancy if tid not in active]:
            del self._occupancy[tid]
        if events:
            self._write(events)
        return events

    def _process_rows(self, tracks, frame_id: int, now: float) -> list:
        import torch
        n = len(tracks)
        if n == 0:
            return []
        if n > self._tables.capacity:
            cap = self._tables.capacity
            while cap < n:
                cap *= 2
            self._build(cap)
        z, tt = self._tables, self._track
        ncol, cap = z.num_columns, z.capacity
        ids = np.fromiter((int(t.track_id) for t in tracks), np.int32, n)
        state = np.empty((2, ncol, cap), np.float64)
        state[0] = np.nan
        state[1] = 0.0
        for r, tid in enumerate(ids.tolist()):
            fs, la = self._occupancy.get(tid), self._cooldown.get(tid)
            if fs is not None:
                state[0, :, r] = fs
            if la is not None:
                state[1, :, r] = la
        rows = np.zeros((n, 6), np.float32)               # xyxy, then track id and class id as raw int32 bits
        rows[:, :4] = np.stack([np.asarray(t.xyxy, np.float32).reshape(4) for t in tracks])
        rows[:, 4] = ids.view(np.float32)
        rows[:, 5] = np.fromiter((int(t.class_id) for t in tracks), np.int32, n).view(np.float32)
        with torch.cuda.device(self.device):
            d_rows = torch.from_numpy(rows).to(self.device)
            tt.xyxy[0, :n] = d_rows[:, :4]
            tt.track_id[0, :n] = d_rows[:, 4].view(torch.int32)
            tt.class_id[0, :n] = d_rows[:, 5].view(torch.int32)
            tt.count.fill_(n)
            fs_dev, la_dev, st = z.state_in()
            d_state = torch.from_numpy(state).to(self.device)
            fs_dev[0].copy_(d_state[0])
            la_dev[0].copy_(d_state[1])
            self._status.zero_()
            _lib.check(self._lib.rtm_zone_step(
                C.byref(z.zone_set), C.byref(tt.struct), None, C.byref(st), C.byref(st), float(now), None,
                int(frame_id), z.events.data_ptr(), z.event_stride, z.event_count.data_ptr(),
                self._status.data_ptr(), _lib.cuda_stream()))
            out = torch.stack([fs_dev[0], la_dev[0]]).cpu().numpy()
            count = z.event_count.cpu().numpy()
            _lib.raise_on_status(self._status.cpu().numpy(), "ZoneEventEngine")
            names = {int(t.track_id): getattr(t, "class_name", "") for t in tracks}
            ev_host = z.events[:, :max(int(count[0]), 1)].cpu().numpy()
            events = z.decode_events(ev_host, count, class_names=lambda tid: names.get(tid, ""))[0]
        for r, tid in enumerate(ids.tolist()):
            fs, la = out[0, :, r], out[1, :, r]
            if np.isnan(fs).all():
                self._occupancy.pop(tid, None)
            else:
                self._occupancy[tid] = fs.copy()
            if la.any():
                self._cooldown[tid] = la.copy()
        return events                                     # (row, zone) order = the reference's (track, zone) order

    def get_zone_polygons(self) -> list:
        """For visualization overlay (zone_engine.py:134-136)."""
        return [(z.name, z.polygon) for z in self.zones]

    def _write(self, events) -> None:
        with open(self.log_path, "a") as f:
            f.write("".join(e.to_json() + "\n" for e in events))
