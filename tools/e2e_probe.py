#!/usr/bin/env python
"""Wall-clock breakdown of the public end-to-end call at BASELINE config-2 scale (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacepharer_b200 import api, synth

tres, toff, _ = synth.make_targets(n_genomes=8000)
qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
s = api.Searcher(0)
masked, _ = api.tantan_mask(s.seed, tres, toff)
s.build_index(masked, toff)
s.set_align_targets(tres, toff)
db = int(toff[-1])
for it in range(4):
    t = [time.perf_counter()]
    s.set_queries(qres, qoff); t.append(time.perf_counter())
    pst = s.prefilter_resident(); t.append(time.perf_counter())
    s.pairs_from_prefilter(); t.append(time.perf_counter())
    ast = s.align_resident(db); t.append(time.perf_counter())
    off, rows, cig, _ = s.align_fetch(); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print("set_queries %.2f | prefilter %.2f (kernels %.2f, device total %.2f) | pairs %.2f | align %.2f (fwd %.2f, device total %.2f) | fetch %.2f | sum %.2f ms"
          % (d[0], d[1], pst["kernel_ms"], pst["total_ms"], d[2], d[3], ast["kernel_ms"], ast["total_ms"], d[4], sum(d)))
