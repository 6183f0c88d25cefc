#!/usr/bin/env python
"""Quick performance probe at BASELINE config-2 scale (development aid, not the benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacepharer_b200 import api, synth

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
t0 = time.time(); tres, toff, _ = synth.make_targets(n_genomes=ng); print("targets", len(toff) - 1, len(tres), "gen %.1fs" % (time.time() - t0))
t0 = time.time(); qres, qoff, _ = synth.make_query_frames(ns, tres, toff); print("frames", len(qoff) - 1, "gen %.1fs" % (time.time() - t0))
s = api.Searcher(0)
t0 = time.time(); masked, nm = api.tantan_mask(s.seed, tres, toff); print("tantan %.2fs masked %d" % (time.time() - t0, nm))
t0 = time.time(); st = s.build_index(masked, toff); print("index build wall %.2fs" % (time.time() - t0), st)
for it in range(3):
    t0 = time.time(); off, hits, pst = s.prefilter(qres, qoff); w = time.time() - t0
    print("prefilter wall %.3fs kernel %.3f ms hits %d" % (w, pst["kernel_ms"], len(hits)), pst)
s.set_align_targets(tres, toff)
for it in range(3):
    t0 = time.time(); aoff, rows, cig, ast = s.align(qres, qoff, off, hits["id"], int(toff[-1])); w = time.time() - t0
    print("align wall %.3fs fwd kernel %.3f ms total %.3f ms rows %d" % (w, ast["kernel_ms"], ast["total_ms"], len(rows)), ast,
          "GCUPS %.2f" % (ast["n_cells"] / ast["kernel_ms"] / 1e6))
Q = len(qoff) - 1
print("queries/s (kernels only): %.0f" % (Q / ((pst["kernel_ms"] + ast["kernel_ms"]) / 1e3)))

if os.environ.get("SPB_PROBE_BANDS"):
    ql = rows["q_end"].astype(np.int64) - rows["q_start"] + 1
    tl = rows["t_end"].astype(np.int64) - rows["t_start"] + 1
    bw = np.abs(tl - ql) + 1
    print("traceback boxes: band classes", {int(b): int((bw == b).sum()) for b in np.unique(bw)[:12]}, "max band", int(bw.max()), "boxes wider than 4:", int((bw > 4).sum()),
          "max qLen", int(ql.max()), "max tLen", int(tl.max()), "cells of the widest", int((ql * (2 * bw + 1))[np.argmax(bw)]))
