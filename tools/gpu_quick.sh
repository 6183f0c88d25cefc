#!/bin/bash
# Quick GPU pass: a subset of the parity tests, then the perf probe once per environment setting given as an argument.
# Usage under gpurun: bash tools/gpu_quick.sh <tag> "<pytest -k expression or ''>" ["VAR=value VAR2=value" ...]
#   e.g. bash tools/gpu_quick.sh r2x "prefilter or align" "SPB_EXPAND_HEAVY=65535" "SPB_EXPAND_HEAVY=2048" "SPB_LOOKUP_CTAS=5"
TAG=$1; KEXPR=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
i=0
for cfg in "${@:-}"; do
  echo "== ${cfg:-defaults}"
  env $cfg SPB_DEBUG=1 timeout 300 python tools/perf_probe.py > $OUT/quick_${TAG}_$i.log 2>&1
  grep "spb_prefilter:" $OUT/quick_${TAG}_$i.log | tail -2 | sed -e 's/.*handed to the CTA kernel[^;]*; //'
  grep "align wall" $OUT/quick_${TAG}_$i.log | tail -1 | sed -e 's/{.*}//'
  i=$((i + 1))
done
