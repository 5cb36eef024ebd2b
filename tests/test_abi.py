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
    assert ctypes.sizeof(L.StepHostIO) == 16 * 8


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


def test_ctypes_structs_match_the_header_as_a_c_compiler_lays_it_out(pkg, tmp_path):
    """include/rtmodt_b200.h is plain C: gcc compiles it, prints sizeof / offsetof of every struct the Python binding
    mirrors, and the numbers have to agree with the ctypes Structures field by field."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    L = pkg._lib
    pairs = [("rtm_nms_params", L.NmsParams), ("rtm_track_table", L.TrackTable), ("rtm_kalman_state", L.KalmanState),
             ("rtm_track_options", L.TrackOptions), ("rtm_zone_set", L.ZoneSet), ("rtm_zone_state", L.ZoneState),
             ("rtm_zone_event", L.ZoneEventRec), ("rtm_step_io", L.StepIO), ("rtm_step_host_io", L.StepHostIO)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtmodt_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
