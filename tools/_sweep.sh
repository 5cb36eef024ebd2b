timeout 300 python -m pytest tests/test_gpu_detect.py -x -q -m gpu -k "letterbox or detector" 2>&1 | tail -2
for impl in fast direct; do
RTM_LETTERBOX_IMPL=$impl python - <<'P'
import sys, importlib, json, os
sys.path.insert(0, '.')
import torch, bench
pkg = importlib.import_module("rtmodt_b200")
dev = torch.device("cuda", 0)
r = bench.letterbox_bench(pkg, pkg._lib.lib(), dev, 64, 6549.4)
print(os.environ.get("RTM_LETTERBOX_IMPL"), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
P
done
