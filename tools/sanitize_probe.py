#!/usr/bin/env python
"""Small end-to-end pass for compute-sanitizer (memcheck): index build, prefilter (team + fallback kernels), align, nucleotide
re-alignment, frames, exhaustive search, lanes.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spacepharer_b200 import api, multigpu, synth

tres, toff, set_of = synth.make_targets(n_genomes=30, seed=3)
qres, qoff, _, _, qnt = synth.make_spacer_arrays(40, tres, toff, set_of, seed=4, return_nucl=True)
eng = multigpu.ShardedSearcher(device=0, lanes=2)
s = eng.searcher
masked, _ = api.tantan_mask(s.seed, tres, toff)
eng.load_targets(masked, tres, toff)
db = int(toff[-1])
for max_seqs in (300, 5):
    s.set_queries(qres, qoff)
    s.prefilter_resident(max_seqs=max_seqs)
    s.pairs_from_prefilter()
    s.align_resident(db)
    off, rows, cig, _ = s.align_fetch()
    print("max_seqs", max_seqs, "rows", len(rows))
s.set_nucl_targets(synth.back_translate(tres, toff))
s.set_nucl_queries(qnt)
n, _ = s.nucl_rescore(3 * db)
print("nucl rows", len(n), "errors", int((n["err"] != 0).sum()))
v = s.align_fetch_begin(); v.wait(); v.close()
sts = eng.run_stream([(qres, qoff)] * 4, db)
print("lanes", sorted({st["lane"] for st in sts}))
s.set_queries(qres[: int(qoff[50])], qoff[:51])
print("exhaustive", s.pairs_exhaustive(db)["n_pairs"])
nt = np.concatenate([np.frombuffer(x, dtype=np.uint8) for x in qnt[:200]])
code = np.full(256, 4, dtype=np.uint8); code[[65, 67, 71, 84]] = [0, 1, 2, 3]
ntoff = np.zeros(201, dtype=np.uint64); np.cumsum([len(x) for x in qnt[:200]], out=ntoff[1:])
fr = s.spacer_frames(code[nt], ntoff)
print("frames", len(fr[1]) - 1 if isinstance(fr, tuple) else "ok")
eng.close()
print("sanitize probe done")
