#!/bin/bash
# Quick GPU pass: a subset of the parity tests, then the perf probe under the SPB_TUNE variants given as arguments.
# Usage under gpurun: bash tools/gpu_quick.sh <tag> "<pytest -k expression or ''>" [tune values...]
TAG=$1; KEXPR=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
for t in "$@"; do
  echo "== SPB_TUNE=$t"
  SPB_TUNE=$t SPB_DEBUG=1 timeout 300 python tools/perf_probe.py > $OUT/quick_${TAG}_$t.log 2>&1
  grep "spb_prefilter:" $OUT/quick_${TAG}_$t.log | tail -2 | sed -e 's/.*handed to the CTA kernel[^;]*; //'
  grep "align wall" $OUT/quick_${TAG}_$t.log | tail -1 | sed -e 's/{.*}//'
done
