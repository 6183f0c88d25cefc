#!/usr/bin/env python
"""Phase timing of the target-split step (development aid): torchrun --nproc-per-node 2 tools/split_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from spacepharer_b200 import api, multigpu, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tres, toff, _ = synth.make_targets(n_genomes=8000)
qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
eng = multigpu.ShardedSearcher(device=local, rank=rank, world=world, split_mode="target")
masked, _ = api.tantan_mask(eng.searcher.seed, tres, toff)
eng.load_targets(masked, tres, toff)
db_res = int(toff[-1])
eng.set_queries(qres, qoff)
s = eng.searcher


def sync():
    torch.cuda.synchronize()


orig = {}
acc = {}


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        sync(); t0 = time.perf_counter()
        r = f(*a, **k)
        sync(); acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    setattr(obj, name, g)


for n in ("prefilter_resident", "merge_shards", "pairs_from_prefilter", "align_resident"):
    wrap(s, n)
_ex = multigpu.exchange_hit_lists


def ex(*a, **k):
    sync(); t0 = time.perf_counter()
    r = _ex(*a, **k)
    sync(); acc.setdefault("exchange", []).append((time.perf_counter() - t0) * 1e3)
    return r


multigpu.exchange_hit_lists = ex
for it in range(8):
    sync(); dist.barrier(); sync()
    t0 = time.perf_counter()
    st = eng.step_resident(db_res)
    sync()
    acc.setdefault("step", []).append((time.perf_counter() - t0) * 1e3)
if rank == 0:
    for k, v in acc.items():
        print("%-22s" % k, " ".join("%8.3f" % x for x in v))
    print(st["prefilter"])
dist.destroy_process_group()
