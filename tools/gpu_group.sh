#!/bin/bash
# Index split with shard groups on an N-GPU box: value leg + strong leg only.  Usage: bash tools/gpu_group.sh <tag> <N> <group...>
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
for g in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 \
      --split-mode index --shard-group $g --no-cpu-baseline --exhaustive-spacers 0 > $OUT/bench_${TAG}_g$g.json 2> $OUT/bench_${TAG}_g$g.err; echo "group $g rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_g$g.json"))
    print("group $g", "value %.1fM" % (d["value"]/1e6), "e2e %.1fM" % (d["e2e"]["value"]/1e6), "strong %.1fM %.1f ms" % (d["strong_scaling"]["value"]/1e6, d["strong_scaling"]["ms_total"]),
          "lookup_ms %.3f" % d["roofline"]["kernels_ms"]["lookup_ms"], "replicas %.1fM" % (d.get("query_split",{}).get("value",0)/1e6), "parity", d.get("parity",{}).get("parity_ok"))
except Exception as e:
    print("no json", e)
PY
  tail -2 $OUT/bench_${TAG}_g$g.err
done
