#!/bin/bash
# Next experiment on the step sequence (profiles/r1g_overlap_probes.txt): let consecutive head scans
# overlap (RTM_SCAN_TRIGGER=1) while leaving room for the post kernel BY CONSTRUCTION - at most two scan CTAs per SM
# (76 KB each through RTM_TMA_SMEM_PAD_KB=6: three no longer fit, also when two scans overlap) beside a post kernel
# built small (-DRTM_NMS_SMEM_CAND=1024 -DRTM_TRACK_PREF_ROWS=128 -DRTM_ZONE_PREF_VERTICES=512
# -DRTM_POST_CTAS_PER_SM=2: ~71 KB and 64 registers; 2 x 76 + 71 = 223 KB of the SM's 227).
#   here:      python tools/exp_leave_room.sh --build        (no GPU needed)
#   GPU box:   gpurun --timeout 900 -- 'bash tools/exp_leave_room.sh'
cd "$(dirname "$0")/.."
if [ "$1" = "--build" ]; then
  python - <<'PY'
import importlib
b = importlib.import_module("real-time-multi-object-detection---tracking-system_b200.build")
print(b.build(force=True, out="tools/librtmodt_b200_small.so",
              defines=("RTM_NMS_SMEM_CAND=1024", "RTM_TRACK_PREF_ROWS=128", "RTM_ZONE_PREF_VERTICES=512", "RTM_POST_CTAS_PER_SM=2")))
PY
  exit $?
fi
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 20 --no-e2e --no-cpu --no-extras > gpurun_out/room_$name.json 2> gpurun_out/room_$name.err
  echo "$name rc=$? $(python - <<PY
import json
d=json.load(open('gpurun_out/room_$name.json'))
print('us/step %.2f single %.2f parity %s decode %.2f post %.2f' % (d['ms_per_step']*1e3, d['single_stream']['ms_per_step']*1e3, d['parity']['ok'], d['kernels']['decode']['avg_us'], d['kernels']['post']['avg_us']))
PY
)"
}
SMALL="RTM_LIB_PATH=$PWD/tools/librtmodt_b200_small.so"
run base
run small $SMALL
run small_pad6 $SMALL RTM_TMA_SMEM_PAD_KB=6
run small_pad6_trigger $SMALL RTM_TMA_SMEM_PAD_KB=6 RTM_SCAN_TRIGGER=1
run small_pad6_trigger_s0 $SMALL RTM_TMA_SMEM_PAD_KB=6 RTM_SCAN_TRIGGER=1 RTM_TMA_STATIC_ROUNDS=0
