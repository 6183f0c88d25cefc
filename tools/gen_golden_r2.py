#!/usr/bin/env python
"""Round-2 golden fixtures from the REFERENCE binary (oracle/_ref/spacepharer, `make -C oracle ref`); development time only.

  tests/golden/small/pref_s{1.0,4.0,7.5}.txt.gz   the `small` set at -s 1.0 / 4.0 / 7.5 (k-mer thresholds 154 / 127 / 96,
                                                  Prefiltering.cpp:987-1021): thr 127 and 154 exercise the index-side self-score
                                                  pruning (IndexTable.h:141-148), thr 96 doubles the similar-k-mer lists
  tests/golden/saturated/                         6 families x 40 diverged copies of 160-aa ORFs, queries of 60-140 aa cut out
                                                  of the family bases: ungapped diagonal scores saturate the 8-bit counters, so the
                                                  max-seqs cut falls among saturated candidates -> rescoreHits
                                                  (QueryMatcher.cpp:149-155,486-505); prefilter lists at --max-seqs 300 / 10 / 3
Usage: python tools/gen_golden_r2.py
"""
import gzip
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import dump_db, run, write_fasta  # noqa: E402
from oracle import refbin  # noqa: E402
from spacepharer_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def saturated_synthetic(seed=29):
    rng = np.random.default_rng(seed)
    p = synth.BG40 / synth.BG40.sum()
    T, Q, bases = [], [], []
    for f in range(6):
        base = rng.choice(20, size=160, p=p).astype(np.uint8)
        bases.append(base)
        for c in range(40):
            t = base.copy()
            for _ in range(int(len(t) * rng.uniform(0.0, 0.06))):
                t[int(rng.integers(0, len(t)))] = rng.choice(20, p=p)
            pre = rng.choice(20, size=int(rng.integers(0, 30)), p=p).astype(np.uint8)
            T.append(np.concatenate([pre, t]))
    order = rng.permutation(len(T))
    T = [T[i] for i in order]
    for i in range(90):
        b = bases[int(rng.integers(0, len(bases)))]
        L = int(rng.integers(60, 141))
        s = int(rng.integers(0, len(b) - L + 1))
        q = b[s:s + L].copy()
        for _ in range(int(rng.integers(0, 3))):
            q[int(rng.integers(0, L))] = rng.choice(20, p=p)
        Q.append(q)
    return [synth.decode(t) for t in T], [synth.decode(q) for q in Q]


def gunzip_to(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        g.write(f.read())


def main():
    assert refbin.available(), "build the reference first: make -C oracle ref"
    B = refbin.REF_BIN
    work = tempfile.mkdtemp(prefix="spgold2_")
    # ---- small set at other sensitivities
    w = os.path.join(work, "small")
    os.makedirs(w)
    for n in ("t", "q"):
        gunzip_to(os.path.join(GOLD, "small", n + ".fasta.gz"), os.path.join(w, n + ".fasta"))
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w, n, w, n))
    for sens in ("1.0", "4.0", "7.5"):
        name = "pref_s" + sens
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w, w, w, name, refbin.PREF_FLAGS.format(maxseqs=300, mask=1, threads=4, sens=sens)))
        dump_db(os.path.join(w, name), os.path.join(GOLD, "small", name + ".txt.gz"))
    # ---- saturated families
    out = os.path.join(GOLD, "saturated")
    os.makedirs(out, exist_ok=True)
    w = os.path.join(work, "sat")
    os.makedirs(w)
    T, Q = saturated_synthetic()
    write_fasta(os.path.join(w, "t.fasta"), T, "t")
    write_fasta(os.path.join(w, "q.fasta"), Q, "q")
    for n in ("t", "q"):
        with open(os.path.join(w, n + ".fasta"), "rb") as f, gzip.open(os.path.join(out, n + ".fasta.gz"), "wb", 9) as g:
            g.write(f.read())
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w, n, w, n))
    for maxseqs in (300, 10, 3):
        name = "pref_max%d" % maxseqs
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w, w, w, name, refbin.PREF_FLAGS.format(maxseqs=maxseqs, mask=1, threads=4, sens="5.7")))
        dump_db(os.path.join(w, name), os.path.join(out, name + ".txt.gz"))
    shutil.rmtree(work)
    print("round-2 fixtures written")


if __name__ == "__main__":
    main()
