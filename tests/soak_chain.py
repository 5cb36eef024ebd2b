"""One-off soak (not collected by pytest; run by hand on the GPU box): the step kernel against the oracle chain on
workloads of very different candidate densities - 3 to 200 planted objects per stream, i.e. from a handful to several
thousand candidates per stream-frame (candidate queue under back-pressure, NMS spill path, crowded class segments) -
in both step modes.

    python tests/soak_chain.py [streams] [frames]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

pkg = importlib.import_module("rtmodt_b200")
from rtmodt_b200.workload import PostBackboneWorkload
from oracle import chain

S = int(sys.argv[1]) if len(sys.argv) > 1 else 12
F = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
bad = 0
for objects in (3, 30, 90, 200):
    for dtype in (torch.bfloat16, torch.float16):
        wl = PostBackboneWorkload(S, F, first_stream=7 * objects, device=dev, dtype=dtype, num_objects=objects)
        for ready in (None, True):
            sb = pkg.StreamBatch(S, wl.zones, src_hw=(1080, 1920), classes=[0, 1, 2, 3, 5, 7], max_tracks=1024, device=dev)
            out = chain.run_chain_parity(sb, lambda f: wl.heads[f], lambda f: wl.host_frame(f), wl.zones, F,
                                         classes=[0, 1, 2, 3, 5, 7], heads_ready=ready, digests=False)
            sb.close()
            print(f"objects {objects:3d} {str(dtype)[6:]:8s} {'async' if ready else 'sync '}: ok={out['ok']} detections {out['detections_checked']} "
                  f"events {out['events_checked']} flips {out['nms_index_flips']} box {out['box_mismatch']} ids {out['track_id_mismatch']} "
                  f"tables {out['track_table_mismatch']} events {out['event_mismatch']}", flush=True)
            bad += not out["ok"]
print("soak", "FAILED" if bad else "ok")
sys.exit(1 if bad else 0)
