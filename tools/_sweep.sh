timeout 300 python -m pytest tests/test_gpu_detect.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -1
for i in 1 2; do
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']*1e3,1),'us/step', {k:round(v['avg_us'],1) for k,v in d['kernels'].items()}, d['parity']['ok'])"
done
