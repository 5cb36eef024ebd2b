"""rtmodt-b200: B200-native post-backbone hot path of RTMODT.

Drop-in classes (same names, arguments and errors as the reference):

    Detector, Detections            <- src/detection/detector.py
    MultiObjectTracker, Track       <- src/tracking/tracker.py
    ZoneEventEngine, ZoneEvent, Zone<- src/events/zone_engine.py

Batched product path: :class:`StreamBatch` (B streams per launch, state resident in HBM).
All arithmetic runs in ``librtmodt_b200.so`` (hand-written CUDA, sm_100a); there is no CPU
fallback - entry points raise ``RtmError`` when the library or the GPU is missing.
"""

from . import _lib, sharding, synth, workload  # noqa: F401
from ._lib import RtmError  # noqa: F401
from .streams import DeviceTrackTable, HostFeeder, StreamBatch, Zone, ZoneEvent, ZoneTables  # noqa: F401
from .tracking import ByteTracker, MultiObjectTracker, Track  # noqa: F401
from .events import ZoneEngine, ZoneEventEngine  # noqa: F401
from .detection import Detections, Detector  # noqa: F401

__version__ = "0.1.0"
