#!/usr/bin/env python
"""Wall-clock phases of the pipelined end-to-end loop (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spacepharer_b200 import api, multigpu, synth

tres, toff, _ = synth.make_targets(n_genomes=8000)
qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
eng = multigpu.ShardedSearcher(device=0)
masked, _ = api.tantan_mask(eng.searcher.seed, tres, toff)
eng.load_targets(masked, tres, toff)
db_res = int(toff[-1])
q_pin = torch.from_numpy(np.ascontiguousarray(qres)).pin_memory().numpy()
o_pin = torch.from_numpy(np.ascontiguousarray(qoff).view(np.int64)).pin_memory().numpy().view(np.uint64)
s = eng.searcher
for mode in ("sync", "pipe", "pipe"):
    acc = {}
    def T(name, f):
        t0 = time.perf_counter(); r = f(); acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
    pend = None
    t_all = time.perf_counter()
    for k in range(12):
        T("set_queries", lambda: eng.set_queries(q_pin, o_pin))
        T("step", lambda: eng.step_resident(db_res))
        if mode == "sync":
            v = T("fetch", lambda: s.align_fetch_view())
            T("close", lambda: v.close())
        else:
            cur = T("begin", lambda: s.align_fetch_begin())
            if pend is not None:
                T("wait", lambda: pend.wait())
                T("close", lambda: pend.close())
            pend = cur
    if pend is not None:
        pend.wait(); pend.close()
    tot = (time.perf_counter() - t_all) * 1e3 / 12
    print(mode, "ms/step %.3f" % tot, {k: round(float(np.mean(v[2:])), 3) for k, v in acc.items()})
