from .zone_engine import Zone, ZoneEvent, ZoneEventEngine

ZoneEngine = ZoneEventEngine  # the name BASELINE.json's north_star uses

__all__ = ["ZoneEventEngine", "ZoneEvent", "Zone", "ZoneEngine"]
