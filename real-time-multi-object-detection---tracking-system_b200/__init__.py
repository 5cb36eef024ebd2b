"""rtmodt-b200: B200-native post-backbone hot path of RTMODT."""

from . import synth  # noqa: F401  (host-only NumPy workload generators)
