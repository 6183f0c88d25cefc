#!/usr/bin/env python
"""Two (or more) contexts on one GPU, each searching its own stream of batches from its own host thread (development aid):
does kernel concurrency fill the tails / host-sync bubbles of a 10k-spacer batch?"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spacepharer_b200 import api, synth

tres, toff, _ = synth.make_targets(n_genomes=8000)
qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
db_res = int(toff[-1])
Q = len(qoff) - 1
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 2
STEPS = 30
lanes = []
masked = None
for i in range(NL):
    s = api.Searcher(0)
    if masked is None:
        masked, _ = api.tantan_mask(s.seed, tres, toff)
    s.build_index(masked, toff)
    s.set_align_targets(tres, toff)
    s.set_queries(qres, qoff)
    lanes.append(s)


def work(s, n, e2e):
    for _ in range(n):
        if e2e:
            s.set_queries(qres, qoff)
        s.prefilter_resident(); s.pairs_from_prefilter(); s.align_resident(db_res)
        if e2e:
            v = s.align_fetch_view(); v.close()


for e2e in (False, True):
    for nl in range(1, NL + 1):
        for s in lanes[:nl]:
            work(s, 3, e2e)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(s, STEPS, e2e)) for s in lanes[:nl]]
        [t.start() for t in th]; [t.join() for t in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("e2e" if e2e else "resident", "lanes", nl, "ms per batch %.3f" % (dt * 1e3 / (STEPS * nl)), "Mq/s %.2f" % (Q * STEPS * nl / dt / 1e6))
