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
if [[ " $* " == *" lanesweep "* ]]; then
  B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-parity --strong-spacers 0 --exhaustive-spacers 0"
  for cfg in "SPB_LOOKUP_CTAS=6" "SPB_LOOKUP_CTAS=4" "SPB_LOOKUP_CTAS=3" "SPB_LOOKUP_CTAS=4 SPB_SOLO_CTAS=6 SPB_TEAM_CTAS=3" "SPB_EXPAND_CTAS=6 SPB_LOOKUP_CTAS=4"; do
    for lanes in 3 4; do
      echo "$cfg lanes=$lanes: $(env $cfg timeout 300 $B --lanes $lanes 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value %.2fM e2e %.2fM ms %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]))')"
    done
  done
fi
if [[ " $* " == *" ncu3 "* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bookkeep_kernel|expand_kernel' -s 6 -c 4 \
      -o $OUT/prof3_$TAG -f python tools/perf_probe.py > $OUT/prof3_$TAG.log 2>&1; echo "ncu3 rc=$?"
fi
if [[ " $* " == *" ncu2 "* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bookkeep_kernel|sort_rows|sw_traceback|sw_reverse|align_profile|sw_fast' -s 8 -c 8 \
      -o $OUT/prof2_$TAG -f python tools/perf_probe.py > $OUT/prof2_$TAG.log 2>&1; echo "ncu2 rc=$?"
fi
if [[ " $* " == *" l2 "* ]]; then   # (r2c: cudaLimitMaxL2FetchGranularity 32/64/128 leaves the DRAM bytes unchanged)
  for g in 32 64 128; do
    SPB_L2_FETCH=$g timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'lookup_kernel|bookkeep_kernel|expand_kernel' -s 4 -c 4 --csv --log-file $OUT/l2fetch_${g}_$TAG.csv python tools/perf_probe.py > /dev/null 2>&1
    SPB_L2_FETCH=$g SPB_DEBUG=1 timeout 600 python tools/perf_probe.py 2>&1 | grep "spb_prefilter:" | tail -1 > $OUT/l2fetch_${g}_$TAG.txt
  done
fi
if [[ " $* " == *" sweep "* ]]; then
  for v in 1024 4096; do
    echo "SPB_HITS_SOLO=$v"; SPB_HITS_SOLO=$v SPB_DEBUG=1 timeout 600 python tools/perf_probe.py 2>&1 | grep "spb_prefilter:" | tail -1 | grep -o "expand.*"
  done

  for v in "4 2" "5 1" "6 1" "3 2"; do
    set -- $v
    echo "SPB_SOLO_CTAS=$1 SPB_TEAM_CTAS=$2"; SPB_SOLO_CTAS=$1 SPB_TEAM_CTAS=$2 SPB_DEBUG=1 timeout 600 python tools/perf_probe.py 2>&1 | grep "spb_prefilter:" | tail -1 | grep -o "expand.*"
  done
fi
echo done
