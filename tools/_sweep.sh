python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), d['opt_in_modes'], d['dense_crowd']['ms_per_step'], d['letterbox']['avg_launch_us'])"
