"""pytest configuration: the ``gpu`` marker and shared helpers.

``-m "not gpu"`` runs here on the CPU (oracle vs. the reference's golden vectors, host
logic, C-ABI symbol checks); ``-m gpu`` are the parity tests proper and need a B200.
"""

import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name is not a Python identifier)."""
    return importlib.import_module("rtmodt_b200")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def golden_clip(g):
    """[(xyxy, conf, cls)] per frame from a golden file."""
    o = g["det_offsets"]
    return [(g["det_xyxy"][o[f]:o[f + 1]], g["det_conf"][o[f]:o[f + 1]], g["det_cls"][o[f]:o[f + 1]])
            for f in range(len(o) - 1)]


def golden_state(g, f):
    """The reference's ``_core._tracks`` after frame f as parallel arrays."""
    o = g["state_offsets"]
    s = slice(o[f], o[f + 1])
    return dict(track_id=g["state_track_id"][s], xyxy=g["state_xyxy"][s], conf=g["state_conf"][s],
                cls=g["state_cls"][s], age=g["state_age"][s], tsu=g["state_tsu"][s])
