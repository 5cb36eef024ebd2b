"""Oracle: CPU restatement of the reference's polygon-zone event engine.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/src/events/zone_engine.py``:

  * ``centroid``            <- zone_engine.py:90-91 (float32 sum, /2, truncation toward 0)
  * ``point_in_polygon``    <- ``cv2.pointPolygonTest(poly_int32, (cx, cy), False)`` as called
                               at zone_engine.py:94 - the integer branch of OpenCV's
                               routine (OpenCV 4.13 is the version importable here);
                               ``>= 0`` means inside, edges and vertices count as inside
  * ``ZoneOracle.process``  <- ``ZoneEventEngine.process``   (zone_engine.py:82-132)

The clock is injected (``now`` argument) instead of ``time.time()``
(zone_engine.py:84); ``timestamp_utc`` (zone_engine.py:108, real wall clock) is not
part of the comparison.  State is keyed exactly like the reference: occupancy by
``track_id -> {zone_name: first_seen}``, cooldown by ``(track_id, zone_name)``, so two
zones with one name share state.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def centroid(xyxy) -> tuple[int, int]:
    """zone_engine.py:90-91 on a float32 box."""
    b = np.asarray(xyxy, np.float32)
    return int((b[0] + b[2]) / np.float32(2)), int((b[1] + b[3]) / np.float32(2))


def point_in_polygon(poly: np.ndarray, px: int, py: int) -> int:
    """+1 inside, 0 on an edge / vertex, -1 outside; integer arithmetic only.

    Restates the integer path of OpenCV's ``pointPolygonTest`` (measureDist=False,
    CV_32S contour): walk the edges (v0 -> v), skip those that cannot cross the
    ray, report 0 when the point lies on an edge, otherwise count sign-corrected
    crossings.  Checked against ``cv2.pointPolygonTest`` in
    ``tests/test_oracle_golden.py``.
    """
    k = len(poly)
    if k == 0:
        return -1
    crossings = 0
    vx, vy = int(poly[k - 1][0]), int(poly[k - 1][1])
    for i in range(k):
        v0x, v0y = vx, vy
        vx, vy = int(poly[i][0]), int(poly[i][1])
        if (v0y <= py and vy <= py) or (v0y > py and vy > py) or (v0x < px and vx < px):
            if py == vy and (px == vx or (py == v0y and ((v0x <= px <= vx) or (vx <= px <= v0x)))):
                return 0
            continue
        d = (py - v0y) * (vx - v0x) - (px - v0x) * (vy - v0y)
        if d == 0:
            return 0
        if vy < v0y:
            d = -d
        crossings += d > 0
    return 1 if crossings & 1 else -1


@dataclass
class ZoneRecord:
    """The deterministic fields of ``ZoneEvent`` (zone_engine.py:29-45)."""
    event_type: str
    zone_name: str
    zone_index: int
    track_id: int
    class_id: int
    dwell: float                 # unrounded now - first_seen
    dwell_time_sec: float        # round(dwell, 2), zone_engine.py:114
    bbox_xyxy: list
    centroid: list
    frame_id: int
    class_name: str = ""
    metadata: dict = field(default_factory=dict)


class ZoneOracle:
    """zone_engine.py:64-132 with an injected clock and no file output."""

    def __init__(self, zone_configs, pip=point_in_polygon) -> None:
        self.zones = []
        for cfg in zone_configs:                            # zone_engine.py:142-151
            self.zones.append(dict(
                name=cfg["name"], polygon=np.array(cfg["polygon"], dtype=np.int32),
                trigger=cfg.get("trigger", "intrusion"),
                dwell_time_sec=cfg.get("dwell_time_sec", 2.0),
                cooldown_sec=cfg.get("cooldown_sec", 10.0)))
        self.pip = pip
        self.occupancy: dict[int, dict[str, float]] = {}
        self.cooldown: dict[tuple[int, str], float] = {}

    def process(self, tracks, frame_id: int, now: float):
        """``tracks``: iterable of ``(track_id, xyxy, class_id)``; returns [ZoneRecord]
        in (track order, zone order)."""
        out = []
        seen = set()
        for tid, box, cid in tracks:
            tid = int(tid)
            seen.add(tid)
            cx, cy = centroid(box)
            for zi, z in enumerate(self.zones):
                if self.pip(z["polygon"], cx, cy) >= 0:
                    first = self.occupancy.setdefault(tid, {}).setdefault(z["name"], now)
                    dwell = now - first
                    if dwell >= z["dwell_time_sec"]:
                        key = (tid, z["name"])
                        if now - self.cooldown.get(key, 0.0) >= z["cooldown_sec"]:
                            out.append(ZoneRecord(
                                event_type=z["trigger"], zone_name=z["name"], zone_index=zi,
                                track_id=tid, class_id=int(cid), dwell=dwell,
                                dwell_time_sec=round(dwell, 2),
                                bbox_xyxy=[float(v) for v in np.asarray(box, np.float32)],
                                centroid=[cx, cy], frame_id=frame_id))
                            self.cooldown[key] = now
                elif tid in self.occupancy:
                    self.occupancy[tid].pop(z["name"], None)
        for tid in set(self.occupancy) - seen:              # zone_engine.py:128-130
            del self.occupancy[tid]
        return out


def cv2_pip(poly: np.ndarray, px: int, py: int) -> float:
    """The third-party call the reference itself makes (zone_engine.py:94)."""
    import cv2
    return cv2.pointPolygonTest(poly, (px, py), False)
