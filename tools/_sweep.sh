python - <<'P'
import sys, importlib, json
sys.path.insert(0, '.')
import torch, bench
pkg = importlib.import_module("rtmodt_b200")
print(json.dumps(bench.dense_crowd_bench(pkg, torch.device("cuda", 0))))
P
