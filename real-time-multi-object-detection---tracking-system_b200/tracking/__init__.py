from .tracker import MultiObjectTracker, Track

ByteTracker = MultiObjectTracker  # the name BASELINE.json's north_star uses

__all__ = ["MultiObjectTracker", "Track", "ByteTracker"]
