#!/bin/bash
# One GPU-box session: parity tests, the default bench line, and the ncu evidence the roofline
# numbers are read from.  Run under gpurun from the repo root:
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests] [bench] [launches] [full]'
#
# Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py turns the .ncu-rep / launch
# list into the committed summaries under profiles/.
set -u
TAG=${1:-r1}
shift || true
WANT=${*:-tests bench launches full}
mkdir -p gpurun_out
SHORT="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-extras"

has() { [[ " $WANT " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
fi
if has bench; then
  timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${TAG}.err; cut -c1-300 gpurun_out/bench_${TAG}.json
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
  echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_${TAG}.json
fi
if has launches || has full; then
  timeout 300 $SHORT > gpurun_out/plain_${TAG}.log 2>&1
  rc=$?
  echo "plain rc=$rc"
  if [ $rc -eq 0 ] && has launches; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_launches_${TAG}.log 2>&1
    echo "ncu launches rc=$?"
  fi
  if [ $rc -eq 0 ] && has full; then
    timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'decode_tma_kernel|post_kernel' -s 16 -c 4 -f -o gpurun_out/prof_${TAG} \
      $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
    echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
  fi
fi
