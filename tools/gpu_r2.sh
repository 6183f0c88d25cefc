#!/bin/bash
# Round-2 GPU pass: parity tests, perf probe with per-phase times, bench, optional ncu captures.
# Usage under gpurun: bash tools/gpu_r2.sh <tag> [ncu] [notest]
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc_$TAG.txt
if [[ " $* " != *" notest "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
  tail -15 $OUT/pytest_$TAG.log
fi
SPB_DEBUG=1 timeout 600 python tools/perf_probe.py > $OUT/perf_$TAG.log 2>&1; echo "perf rc=$?"
grep -E "spb_prefilter:|prefilter wall|align wall|queries/s" $OUT/perf_$TAG.log | tail -12
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
if [[ " $* " == *" ncu "* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'expand_kernel|lookup_kernel|bookkeep_kernel' -s 4 -c 4 \
      -o $OUT/prof_$TAG -f python tools/perf_probe.py > $OUT/prof_$TAG.log 2>&1; echo "ncu rc=$?"
fi
echo done
