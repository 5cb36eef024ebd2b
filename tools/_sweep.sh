timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python tools/post_timeline.py 48 | grep -E "count|stage 1|stage 2|whole"
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']*1e3,1),'us/step', {k:round(v['avg_us'],1) for k,v in d['kernels'].items()}, d['parity']['ok'], d['latency_ms_per_step'])"
