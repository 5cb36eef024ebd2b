"""The C-ABI library loads and exports every symbol include/rtmodt_b200.h declares (no GPU)."""

import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rtmodt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for name in ("rtm_letterbox", "rtm_decode_nms", "rtm_nms_pred", "rtm_track_step", "rtm_zone_step",
                 "rtm_post_backbone_step", "rtm_post_backbone_step_host", "rtm_version", "rtm_last_error"):
        assert name in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(header_symbols()) == set(pkg._lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.rtm_version() == 100


def test_struct_layouts_match_the_header(pkg):
    L = pkg._lib
    assert ctypes.sizeof(L.ZoneEventRec) == 64
    import numpy as np
    assert np.dtype(L.EVENT_DTYPE).itemsize == 64
    assert ctypes.sizeof(L.NmsParams) == 8 + 4 * 4 + 32
    assert ctypes.sizeof(L.TrackTable) == 8 + 8 * 8
    assert ctypes.sizeof(L.ZoneSet) == 8 + 6 * 8
    assert ctypes.sizeof(L.StepHostIO) == 14 * 8


def test_compute_entry_points_refuse_to_run_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RtmError):
        pkg.MultiObjectTracker()
    with pytest.raises(pkg.RtmError):
        pkg.ZoneEventEngine([])
    with pytest.raises(pkg.RtmError):
        pkg.StreamBatch(2)


def test_reference_error_behaviour_is_kept(pkg):
    with pytest.raises(NotImplementedError):
        pkg.MultiObjectTracker("deepsort")
    with pytest.raises(ValueError):
        pkg.MultiObjectTracker("sort")
    with pytest.raises(KeyError):
        pkg.streams.parse_zone({"polygon": [[0, 0], [1, 1], [2, 0]]})
    with pytest.raises(KeyError):
        pkg.streams.parse_zone({"name": "z"})
    z = pkg.streams.parse_zone({"name": "z", "polygon": [[0, 0], [4, 0], [4, 4]]})
    assert (z.trigger, z.dwell_time_sec, z.cooldown_sec) == ("intrusion", 2.0, 10.0)


def test_nms_params_class_mask(pkg):
    p = pkg._lib.make_nms_params(classes=[0, 1, 2, 3, 5, 7])
    assert p.class_mask[0] == 0b10101111 and p.class_mask[1] == 0
    p = pkg._lib.make_nms_params(classes=None, num_classes=80)
    assert p.class_mask[0] == 0xFFFFFFFF and p.class_mask[2] == 0xFFFF and p.class_mask[3] == 0
