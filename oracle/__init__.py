"""CPU oracle for the RTMODT post-backbone hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of the reference's hot path so
that the CUDA path can be checked bit for bit (integers / indices) or within the
stated tolerance (floats).  It is NOT part of the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * the shipped package (``real-time-multi-object-detection---tracking-system_b200``)
    never imports it and has no CPU fallback - it raises if the CUDA library is
    missing.

Pinning status (SURVEY.md §8 c):

  * ``tracker_ref`` / ``zone_ref`` are PINNED: ``oracle/make_goldens.py`` runs the
    unmodified reference files (``/root/reference/src/tracking/tracker.py``,
    ``src/events/zone_engine.py``) in the dev container on seeded inputs and
    freezes their outputs in ``tests/golden/``; ``tests/test_oracle_golden.py``
    replays the oracle against those files.  The reference itself ships no golden
    vectors for this path (its only tests are FastAPI smoke tests).
  * ``detect_ref`` (letterbox / head decode / NMS / rescale) restates
    ``ultralytics>=8.1.0`` (un-vendored, not installable here) using the very
    third-party kernels ultralytics calls (``cv2.resize``, ``cv2.copyMakeBorder``,
    ``torchvision.ops.nms``): PARITY UNPINNED at that boundary - there is neither a
    runnable reference nor a reference test that fixes a number.
  * ``kalman_ref`` has no reference counterpart at all (the reference tracker has
    no Kalman filter): PARITY UNPINNED - it follows the canonical ByteTrack filter.
"""
