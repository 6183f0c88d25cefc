#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): NCCL / IPC parity tests and the N-GPU bench.  Usage: bash tools/gpu_multi.sh <tag> <N> [modes...]
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi_$TAG.txt 2>&1
nvidia-smi topo -m >> $OUT/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -q -x -k "two_gpus or index_split or target_split_config2 or byte_identical_to_the_reference_binary and not config4" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -6 $OUT/pytest_$TAG.log
for mode in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
      --split-mode $mode > $OUT/bench_${TAG}_$mode.json 2> $OUT/bench_${TAG}_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$mode.json"))
    print("$mode", "value", d["value"], "e2e", d["e2e"]["value"], "strong", d.get("strong_scaling",{}).get("value"), d.get("strong_scaling",{}).get("ms_total"), "parity", d.get("parity",{}).get("parity_ok"))
except Exception as e:
    print("no json", e)
PY
  tail -3 $OUT/bench_${TAG}_$mode.err
done
echo done
