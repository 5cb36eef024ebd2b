#!/bin/bash
# round 2, closing session: host issue cost, the whole GPU suite, both bench arms with the driver's arguments,
# then the ncu evidence (launch list of the bench command, --set full of the step kernel, of the letterbox kernels
# and of the stand-alone tracker / zone kernels).  tools/ncu_summary.py <tag> condenses gpurun_out/ into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r2b}
mkdir -p gpurun_out
SHORT="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras --no-parity"
timeout 300 python tools/diag_dist.py --tag final --reps 3 2>&1 | grep ^diag | cut -c1-400
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
echo "bench reference rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_${TAG}.json
timeout 300 $SHORT > gpurun_out/plain_${TAG}.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launches_${TAG}.log 2>&1
  echo "ncu launches rc=$?"; grep -c step_kernel gpurun_out/launches_${TAG}.csv
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'step_kernel' -s 10 -c 3 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
fi
timeout 300 python tools/probe_kernels.py > gpurun_out/probe_kernels_${TAG}.log 2>&1; rc=$?; echo "probe_kernels rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'letterbox' -s 4 -c 4 -f -o gpurun_out/prof_${TAG}_letterbox python tools/probe_kernels.py > gpurun_out/ncu_lb_${TAG}.log 2>&1
  echo "ncu letterbox rc=$?"; tail -2 gpurun_out/ncu_lb_${TAG}.log | cut -c1-200
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'track_step_kernel|zone_step_kernel' -s 16 -c 4 -f -o gpurun_out/prof_${TAG}_side python tools/probe_kernels.py > gpurun_out/ncu_side_${TAG}.log 2>&1
  echo "ncu side rc=$?"; tail -2 gpurun_out/ncu_side_${TAG}.log | cut -c1-200
fi
