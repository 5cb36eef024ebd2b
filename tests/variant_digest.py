"""Helper of tests/test_gpu_variants.py (not a test): one small batch stepped through the library in this process'
configuration (environment switches such as RTM_STEP_LAZY are read once per process), printed as one digest line.

    python tests/variant_digest.py [async|sync] [bf16|f16|f32]

Six streams x eight frames of the bench workload plus a letterbox of two frame sizes; the digest covers everything the
CUDA path produced (detections, anchors, keep indices, track ids, track tables, events, letterbox planes)."""
import hashlib
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200 import _lib
from rtmodt_b200.workload import PostBackboneWorkload

mode = sys.argv[1] if len(sys.argv) > 1 else "async"
dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
dev = torch.device("cuda", 0)
S, F = 6, 8
wl = PostBackboneWorkload(S, F, first_stream=3, device=dev, dtype=dtype)
sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=256, device=dev)
h = hashlib.sha1()
for f in range(F):
    sb.step(wl.heads[f], now=1.7e9 + f / 30.0, frame_id=f, heads_ready=True if mode == "async" else None)
    det = sb.read_detections()
    tracks, next_id = sb.read_tracks()
    evs = sb.read_events()
    for s in range(S):
        for k in ("xyxy", "confidence", "class_id", "anchor", "keep", "track_id", "kind"):
            h.update(np.ascontiguousarray(det[s][k]).tobytes())
        h.update(np.asarray([[t["track_id"], t["age"], t["time_since_update"], t["class_id"]] for t in tracks[s]], np.int64).tobytes())
        h.update(np.asarray([t["xyxy"] for t in tracks[s]], np.float32).tobytes())
        h.update(repr([(e.event_type, e.zone_name, e.track_id, e.frame_id, e.dwell_time_sec) for e in evs[s]]).encode())
    h.update(np.asarray(next_id, np.int64).tobytes())
sb.check_status()
lib = _lib.lib()
rng = np.random.default_rng(5)
for (hh, ww) in ((1080, 1920), (720, 1280)):
    frames = torch.from_numpy(rng.integers(0, 256, (2, hh, ww, 3), dtype=np.uint8)).to(dev)
    out = torch.zeros(2, 3, 640, 640, dtype=dtype, device=dev)
    _lib.check(lib.rtm_letterbox(frames.data_ptr(), 2, hh, ww, ww * 3, hh * ww * 3, out.data_ptr(), _lib.dtype_code(dtype), 640, 640,
                                 _lib.cuda_stream()))
    h.update(out.cpu().view(torch.uint8).numpy().tobytes())
print("digest", mode, h.hexdigest())
