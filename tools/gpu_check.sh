#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list and a full capture of the dominant kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh <tag> [quick]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc_$TAG.txt; lscpu | head -20 >> $OUT/nproc_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
if [ "${2:-}" != "quick" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
  cat $OUT/bench_ref_$TAG.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --exhaustive-spacers 50 > $OUT/launches_$TAG.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'prefilter_team_kernel|sw_fast_kernel' -s 3 -c 3 \
      -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --exhaustive-spacers 0 > $OUT/prof_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sw_exhaustive_kernel' -s 2 -c 2 \
      -o $OUT/prof_exh_$TAG -f python tools/exh_probe.py 8000 100 > $OUT/prof_exh_$TAG.log 2>&1
fi
echo done
