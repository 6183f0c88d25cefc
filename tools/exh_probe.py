#!/usr/bin/env python
"""Exhaustive-search probe at BASELINE config-5 scale (development aid): GCUPS of the streaming kernel."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacepharer_b200 import api, synth

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tres, toff, _ = synth.make_targets(n_genomes=ng)
qres, qoff, _ = synth.make_query_frames(ns, tres, toff)
lens = np.diff(qoff).astype(int)
print("targets", len(toff) - 1, len(tres), "frames", len(qoff) - 1, "length histogram", np.bincount(lens).tolist())
s = api.Searcher(0)
s.set_align_targets(tres, toff)
s.set_queries(qres, qoff)
for it in range(3):
    t0 = time.time(); st = s.pairs_exhaustive(int(toff[-1])); w = time.time() - t0
    print("exhaustive wall %.3fs kernel %.1f ms GCUPS %.1f" % (w, st["kernel_ms"], st["gcups"]), st)
t0 = time.time(); ast = s.align_resident(int(toff[-1])); print("align tail wall %.3fs" % (time.time() - t0), ast)
