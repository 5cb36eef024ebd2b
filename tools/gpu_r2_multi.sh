#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_n2.log
: > $L
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "bench N=$N rc=$?" >> $L
tail -c 1500 gpurun_out/bench_r2_n$N.err >> $L
python - <<PY >> $L
import json
d=json.load(open('gpurun_out/bench_r2_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})
print('single', d['single_stream']['ms_per_step'], 'weak', d['weak_scaling'])
print('parity', d['parity'])
print('e2e', d['e2e'])
print('lat', d.get('latency_ms_per_step'))
print('summary', d['summary'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/diag_dist.py --dist --reps 5 >> $L 2>&1
cat $L | cut -c1-1200
