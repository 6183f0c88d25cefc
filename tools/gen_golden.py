#!/usr/bin/env python
"""Generate tests/golden/ by running the REFERENCE binary (development-time only).

The reference (soedinglab/spacepharer @ d1e94c0) must have been built out of tree, e.g. as in
SURVEY.md 8c:

    mkdir /tmp/refbuild && cd /tmp/refbuild
    CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -DCMAKE_BUILD_TYPE=Release -DCMAKE_POLICY_VERSION_MINIMUM=3.5 \
        -DHAVE_AVX2=1 -DCMAKE_CXX_FLAGS="-include cstdint" /root/reference && make -j8

Usage:  python tools/gen_golden.py [--ref-bin /tmp/refbuild/src/spacepharer] [--ref-src /root/reference]

Nothing under tests/ or the product reads /root/reference or the reference binary at run time;
only the small fixtures written here are committed.
"""
import argparse
import glob
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spacepharer_b200 import synth  # noqa: E402

PREF_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out --seed-sub-mat nucl:nucleotide.out,aa:VTML80.out -k 6 "
              "--k-score 2147483647 --alph-size nucl:5,aa:21 --max-seq-len 65535 --max-seqs {maxseqs} --split 0 "
              "--split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 --comp-bias-corr 1 --diag-score 1 "
              "--exact-kmer-matching 0 --mask {mask} --mask-lower-case 0 --min-ungapped-score 15 --add-self-matches 0 "
              "--spaced-kmer-mode 1 --db-load-mode 0 --pca 1 --pcb 1.5 --spaced-kmer-pattern 11011101 --threads 4 "
              "--compressed 0 -v 3 -s 5.7")
ALN_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out -a 1 --alignment-mode 2 --alignment-output-mode 0 "
             "--wrapped-scoring 0 -e {evalue} --min-seq-id 0 --min-aln-len 0 --seq-id-mode 0 --alt-ali 0 -c 0 --cov-mode 0 "
             "--max-seq-len 65535 --comp-bias-corr 1 --max-rejected 2147483647 --max-accept 2147483647 "
             "--add-self-matches 0 --db-load-mode 0 --pca 1 --pcb 1.5 --score-bias 0 --realign 0 "
             "--realign-score-bias -0.2 --realign-max-seqs 2147483647 --gap-open nucl:10,aa:{go} --gap-extend {ge} "
             "--zdrop 40 --threads 4 --compressed 0 -v 3")


def run(cmd, **kw):
    print("+", cmd if isinstance(cmd, str) else " ".join(cmd), flush=True)
    subprocess.check_call(cmd, shell=isinstance(cmd, str), stdout=subprocess.DEVNULL, **kw)


def load_db(db):
    """MMseqs2 DB -> {key: bytes without the trailing NUL}."""
    if os.path.exists(db):
        data = open(db, "rb").read()
    else:
        parts = sorted(glob.glob(db + ".[0-9]*"), key=lambda x: int(x.rsplit(".", 1)[1]))
        data = b"".join(open(p, "rb").read() for p in parts)
    out = {}
    for line in open(db + ".index"):
        k, o, l = (int(x) for x in line.split())
        out[k] = data[o:o + l - 1]
    return out


def dump_db(db, path):
    d = load_db(db)
    with gzip.open(path, "wb", compresslevel=9) as f:
        for k in sorted(d):
            for line in d[k].splitlines():
                f.write(b"%d\t%s\n" % (k, line))


def copy_db(src, dst):
    """copy a (merged) DB triple; merges split data files"""
    d = load_db(src)
    idx = sorted((int(l.split()[0]), int(l.split()[1]), int(l.split()[2])) for l in open(src + ".index"))
    by_off = sorted(idx, key=lambda r: r[1])
    with open(dst, "wb") as f, open(dst + ".index", "w") as fi:
        new = {}
        pos = 0
        for k, o, l in by_off:
            f.write(d[k] + b"\0")
            new[k] = (pos, l)
            pos += l
        for k, o, l in idx:
            fi.write("%d\t%d\t%d\n" % (k, new[k][0], new[k][1]))
    shutil.copy(src + ".dbtype", dst + ".dbtype")


def write_fasta(path, seqs, prefix):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n%s\n" % (prefix, i, s))


def small_synthetic(seed=7):
    """~400 ORFs and ~700 query fragments exercising: planted substitutions, indels, low-complexity
    (tantan-masked) targets, X residues, diagonal-0 / 256 quirks, very short and longer queries."""
    rng = np.random.default_rng(seed)
    p = synth.BG40 / synth.BG40.sum()
    T = []
    for i in range(400):
        L = int(min(30 + rng.geometric(1 / 77.0), 1800))
        T.append(rng.choice(20, size=L, p=p).astype(np.uint8))
    # low-complexity / repeat targets
    for i in range(0, 400, 25):
        unit = rng.choice(20, size=int(rng.integers(1, 6)), p=p).astype(np.uint8)
        rep = np.tile(unit, 40)[:int(rng.integers(40, 120))]
        pos = int(rng.integers(0, max(1, len(T[i]) - 10)))
        T[i] = np.concatenate([T[i][:pos], rep, T[i][pos:]])
    # a few X residues
    for i in range(3, 400, 37):
        T[i][int(rng.integers(0, len(T[i])))] = 20
    # near-duplicate target families so that several targets pass for one query
    for i in range(10, 400, 40):
        src = T[i - 1].copy()
        for _ in range(max(1, len(src) // 15)):
            src[int(rng.integers(0, len(src)))] = rng.choice(20, p=p)
        T[i] = src
    Q = []
    for i in range(700):
        L = int(rng.integers(9, 24))
        r = rng.random()
        if r < 0.5:
            t = T[int(rng.integers(0, len(T)))]
            if len(t) <= L + 4:
                Q.append(rng.choice(20, size=L, p=p).astype(np.uint8))
                continue
            kind = rng.random()
            if kind < 0.15:
                s = 0
            elif kind < 0.25 and len(t) > 256 + L:
                s = 256
            else:
                s = int(rng.integers(0, len(t) - L - 3))
            q = t[s:s + L + 3].copy()
            nsub = int(rng.integers(0, 4))
            for _ in range(nsub):
                q[int(rng.integers(0, L))] = rng.choice(20, p=p)
            g = rng.random()
            if g < 0.15:      # deletion in the query
                c = int(rng.integers(3, L - 3))
                q = np.concatenate([q[:c], q[c + int(rng.integers(1, 3)):]])
            elif g < 0.3:     # insertion in the query
                c = int(rng.integers(3, L - 3))
                q = np.concatenate([q[:c], rng.choice(20, size=int(rng.integers(1, 3)), p=p).astype(np.uint8), q[c:]])
            Q.append(q[:L].astype(np.uint8))
        elif r < 0.55:
            q = rng.choice(20, size=L, p=p).astype(np.uint8)
            q[int(rng.integers(0, L))] = 20
            Q.append(q)
        elif r < 0.6:
            unit = rng.choice(20, size=2, p=p).astype(np.uint8)
            Q.append(np.tile(unit, 12)[:L])
        else:
            Q.append(rng.choice(20, size=L, p=p).astype(np.uint8))
    Q.append(rng.choice(20, size=7, p=p).astype(np.uint8))   # shorter than the k-mer span
    Q.append(rng.choice(20, size=8, p=p).astype(np.uint8))   # exactly one window
    return [synth.decode(t) for t in T], [synth.decode(q) for q in Q]


def family_synthetic(seed=13):
    """Target families (30 diverged copies each) so that one query reaches dozens of targets:
    exercises the top-max-seqs cut, its tie order and the E-value sort with many rows per query."""
    rng = np.random.default_rng(seed)
    p = synth.BG40 / synth.BG40.sum()
    T, Q = [], []
    bases = []
    for f in range(12):
        base = rng.choice(20, size=int(rng.integers(60, 200)), p=p).astype(np.uint8)
        bases.append(base)
        for c in range(30):
            t = base.copy()
            for _ in range(int(len(t) * rng.uniform(0.0, 0.12))):
                t[int(rng.integers(0, len(t)))] = rng.choice(20, p=p)
            pre = rng.choice(20, size=int(rng.integers(0, 40)), p=p).astype(np.uint8)
            T.append(np.concatenate([pre, t]))
    order = rng.permutation(len(T))
    T = [T[i] for i in order]
    for i in range(150):
        b = bases[int(rng.integers(0, len(bases)))]
        L = int(rng.integers(9, 24))
        s = int(rng.integers(0, len(b) - L))
        q = b[s:s + L].copy()
        for _ in range(int(rng.integers(0, 3))):
            q[int(rng.integers(0, L))] = rng.choice(20, p=p)
        Q.append(q)
    return [synth.decode(t) for t in T], [synth.decode(q) for q in Q]


def long_synthetic(seed=11):
    """Longer queries (up to 150 aa) against related targets: pins the striped-layout dependent
    recurrences of the SW (segLen > 1) and gapped tracebacks."""
    rng = np.random.default_rng(seed)
    p = synth.BG40 / synth.BG40.sum()
    T, Q = [], []
    for i in range(60):
        L = int(rng.integers(40, 400))
        T.append(rng.choice(20, size=L, p=p).astype(np.uint8))
    for i in range(80):
        t = T[int(rng.integers(0, len(T)))]
        L = int(rng.integers(17, min(150, len(t))))
        s = int(rng.integers(0, len(t) - L + 1))
        q = list(t[s:s + L])
        for _ in range(int(rng.integers(0, max(1, L // 6)))):
            q[int(rng.integers(0, len(q)))] = int(rng.choice(20, p=p))
        for _ in range(int(rng.integers(0, 4))):
            c = int(rng.integers(1, len(q) - 1))
            if rng.random() < 0.5:
                del q[c:c + int(rng.integers(1, 4))]
            else:
                q[c:c] = [int(x) for x in rng.choice(20, size=int(rng.integers(1, 4)), p=p)]
        if rng.random() < 0.2:   # tryptophan / cysteine rich: pushes the byte kernel into overflow
            q = [18 if rng.random() < 0.5 else x for x in q]
            T.append(np.array(q, dtype=np.uint8))
        Q.append(np.array(q, dtype=np.uint8))
    return [synth.decode(t) for t in T], [synth.decode(q) for q in Q]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
    ap.add_argument("--ref-src", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    B = a.ref_bin
    out = a.out
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="spgold_")
    ex = os.path.join(a.ref_src, "examples")

    # ---------------- config 1: the reference's own regression inputs -----------------------
    c1 = os.path.join(out, "config1")
    os.makedirs(c1, exist_ok=True)
    w1 = os.path.join(work, "c1")
    os.makedirs(w1)
    gca = " ".join(sorted(glob.glob(os.path.join(ex, "GCA*.fna.gz"))))
    fas = " ".join(sorted(glob.glob(os.path.join(ex, "*.fas"))))
    run("%s createsetdb %s %s/targets %s/tmp --threads 4" % (B, gca, w1, w1))
    run("%s createsetdb %s %s/targets_rev %s/tmp --reverse-fragments 1 --threads 4" % (B, gca, w1, w1))
    run("%s easy-predict %s %s/targets %s/result.tsv %s/tmp --threads 4 --remove-tmp-files 0" % (B, fas, w1, w1, w1))
    qsetdb = glob.glob(os.path.join(w1, "tmp", "*", "qsetdb"))[0]
    pred = os.path.dirname(glob.glob(os.path.join(w1, "tmp", "*", "predict", "*", "prot_result.index"))[0])
    copy_db(qsetdb, os.path.join(c1, "qsetdb"))
    copy_db(os.path.join(w1, "targets"), os.path.join(c1, "targets"))
    copy_db(os.path.join(w1, "targets_rev"), os.path.join(c1, "targets_rev"))
    pref = glob.glob(os.path.join(pred, "search", "*", "pref_0.index"))[0][:-6]
    pref_rev = glob.glob(os.path.join(pred, "search_rev", "*", "pref_0.index"))[0][:-6]
    copy_db(pref, os.path.join(c1, "pref_0"))
    copy_db(os.path.join(pred, "prot_result"), os.path.join(c1, "prot_result"))
    dump_db(pref, os.path.join(c1, "pref_0.txt.gz"))
    dump_db(pref_rev, os.path.join(c1, "pref_0_rev.txt.gz"))
    dump_db(os.path.join(pred, "prot_result"), os.path.join(c1, "prot_result.txt.gz"))
    dump_db(os.path.join(pred, "prot_result_rev"), os.path.join(c1, "prot_result_rev.txt.gz"))
    shutil.copy(os.path.join(w1, "result.tsv"), os.path.join(c1, "result.tsv"))

    # ---------------- small synthetic set ----------------------------------------------------
    s1 = os.path.join(out, "small")
    os.makedirs(s1, exist_ok=True)
    w2 = os.path.join(work, "small")
    os.makedirs(w2)
    T, Q = small_synthetic()
    write_fasta(os.path.join(w2, "t.fasta"), T, "t")
    write_fasta(os.path.join(w2, "q.fasta"), Q, "q")
    for n in ("t", "q"):
        with open(os.path.join(w2, n + ".fasta"), "rb") as f, gzip.open(os.path.join(s1, n + ".fasta.gz"), "wb", 9) as g:
            g.write(f.read())
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w2, n, w2, n))
    for name, mask, maxseqs in (("pref_mask1", 1, 300), ("pref_mask0", 0, 300), ("pref_mask1_max2", 1, 2),
                                ("pref_mask0_max5", 0, 5)):
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w2, w2, w2, name, PREF_FLAGS.format(mask=mask, maxseqs=maxseqs)))
        dump_db(os.path.join(w2, name), os.path.join(s1, name + ".txt.gz"))
    run("%s align %s/qdb %s/tdb %s/pref_mask1 %s/aln %s" % (B, w2, w2, w2, w2, ALN_FLAGS.format(evalue=200, go=16, ge=2)))
    dump_db(os.path.join(w2, "aln"), os.path.join(s1, "aln_mask1.txt.gz"))
    run("%s align %s/qdb %s/tdb %s/pref_mask1 %s/aln_e1 %s" % (B, w2, w2, w2, w2, ALN_FLAGS.format(evalue="1.000E-02", go=16, ge=2)))
    dump_db(os.path.join(w2, "aln_e1"), os.path.join(s1, "aln_mask1_e0.01.txt.gz"))
    # all-vs-all SW on a subset: prefilter DB listing every target for the first 40 queries
    nT = len(T)
    with open(os.path.join(w2, "allpairs"), "wb") as f, open(os.path.join(w2, "allpairs.index"), "w") as fi:
        pos = 0
        for q in range(len(Q)):
            body = b"".join(b"%d\t0\t0\n" % t for t in range(0, nT, 3)) if q < 40 else b""
            f.write(body + b"\0")
            fi.write("%d\t%d\t%d\n" % (q, pos, len(body) + 1))
            pos += len(body) + 1
    shutil.copy(os.path.join(w2, "pref_mask1.dbtype"), os.path.join(w2, "allpairs.dbtype"))
    run("%s align %s/qdb %s/tdb %s/allpairs %s/aln_all %s" % (B, w2, w2, w2, w2, ALN_FLAGS.format(evalue=200, go=16, ge=2)))
    dump_db(os.path.join(w2, "aln_all"), os.path.join(s1, "aln_allpairs.txt.gz"))

    # ---------------- long queries (striped layout, gapped tracebacks) ----------------------
    l1 = os.path.join(out, "long")
    os.makedirs(l1, exist_ok=True)
    w3 = os.path.join(work, "long")
    os.makedirs(w3)
    T, Q = long_synthetic()
    write_fasta(os.path.join(w3, "t.fasta"), T, "t")
    write_fasta(os.path.join(w3, "q.fasta"), Q, "q")
    for n in ("t", "q"):
        with open(os.path.join(w3, n + ".fasta"), "rb") as f, gzip.open(os.path.join(l1, n + ".fasta.gz"), "wb", 9) as g:
            g.write(f.read())
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w3, n, w3, n))
    nT = len(T)
    with open(os.path.join(w3, "allpairs"), "wb") as f, open(os.path.join(w3, "allpairs.index"), "w") as fi:
        pos = 0
        for q in range(len(Q)):
            body = b"".join(b"%d\t0\t0\n" % t for t in range(nT))
            f.write(body + b"\0")
            fi.write("%d\t%d\t%d\n" % (q, pos, len(body) + 1))
            pos += len(body) + 1
    shutil.copy(os.path.join(w2, "pref_mask1.dbtype"), os.path.join(w3, "allpairs.dbtype"))
    run("%s align %s/qdb %s/tdb %s/allpairs %s/aln_all %s" % (B, w3, w3, w3, w3, ALN_FLAGS.format(evalue="1.000E-01", go=16, ge=2)))
    dump_db(os.path.join(w3, "aln_all"), os.path.join(l1, "aln_allpairs_e0.1.txt.gz"))
    run("%s prefilter %s/qdb %s/tdb %s/pref %s" % (B, w3, w3, w3, PREF_FLAGS.format(mask=1, maxseqs=300)))
    dump_db(os.path.join(w3, "pref"), os.path.join(l1, "pref_mask1.txt.gz"))

    # ---------------- target families: truncation + many rows per query ----------------------
    f1 = os.path.join(out, "family")
    os.makedirs(f1, exist_ok=True)
    w4 = os.path.join(work, "family")
    os.makedirs(w4)
    T, Q = family_synthetic()
    write_fasta(os.path.join(w4, "t.fasta"), T, "t")
    write_fasta(os.path.join(w4, "q.fasta"), Q, "q")
    for n in ("t", "q"):
        with open(os.path.join(w4, n + ".fasta"), "rb") as f, gzip.open(os.path.join(f1, n + ".fasta.gz"), "wb", 9) as g:
            g.write(f.read())
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w4, n, w4, n))
    for name, mask, maxseqs in (("pref_max300", 1, 300), ("pref_max10", 1, 10), ("pref_max3", 1, 3)):
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w4, w4, w4, name, PREF_FLAGS.format(mask=mask, maxseqs=maxseqs)))
        dump_db(os.path.join(w4, name), os.path.join(f1, name + ".txt.gz"))
    run("%s align %s/qdb %s/tdb %s/pref_max300 %s/aln %s" % (B, w4, w4, w4, w4, ALN_FLAGS.format(evalue=200, go=16, ge=2)))
    dump_db(os.path.join(w4, "aln"), os.path.join(f1, "aln_max300.txt.gz"))

    shutil.rmtree(work)
    print("golden fixtures written to", out)


if __name__ == "__main__":
    main()
