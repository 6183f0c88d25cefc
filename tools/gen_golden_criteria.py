#!/usr/bin/env python
"""Golden fixtures for the acceptance criteria of `align` from the REFERENCE binary (oracle/_ref/spacepharer); development time only.

  tests/golden/criteria/<set>_<tag>.txt.gz   alignment DBs of the `small` (all-pairs subset, 9-15 aa queries) and `long` (all pairs,
                                             17-150 aa queries, gapped tracebacks) sets under -c / --cov-mode 0..5, --min-seq-id,
                                             --seq-id-mode, --min-aln-len (Alignment.cpp:353-404 canBeCovered + checkCriteria,
                                             StripedSmithWaterman.cpp:195-240 coverage exits, Util.cpp:533-567), and the
                                             10-column rows of `--alignment-mode 3` without `-a` (Matcher.cpp:355-366)
  tests/golden/criteria/pref_long_c0.5_m*    prefilter lists of the `long` set under -c 0.5 --cov-mode 0 / 1 / 2 / 5 (Prefiltering.cpp:834-842)
  tests/golden/criteria/pref_{small,long}_exact  prefilter lists under --exact-kmer-matching 1
  tests/golden/criteria/t_lower.fasta.gz, pref_small_lower_mask{0,1}  targets of `small` with lower-case stretches and the lists under
                                             --mask-lower-case 1 (IndexBuilder.cpp:151-160)
  tests/golden/criteria/pref_family_split3_max20, pref_small_split2_max100  lists of the reference's own target split (--split N --split-mode 0)
  tests/golden/criteria/family_acc* / family_rej*, limits.json  --max-accept / --max-rejected on the `family` set (Alignment.cpp:353-404)
  tests/golden/criteria/*_alt*, alt_ali.json  alternative alignments (--alt-ali N, Alignment.cpp:576-611)
  tests/golden/criteria/*_realign_*, realign.json  --realign 1 with --realign-score-bias / --realign-max-seqs (Alignment.cpp:415-447)
  tests/golden/criteria/cases.json           the flag set of every alignment fixture
Usage: python tools/gen_golden_criteria.py
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import dump_db, run  # noqa: E402
from oracle import refbin  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# tag -> (set, evalue, cov, cov_mode, min_seq_id, min_aln_len, backtrace, seq_id_mode); cov modes: Parameters.h:246-251
# (0 query and target, 1 target, 2 query, 3 / 4 / 5 length ratios)
CASES = {
    "small_c0.05_m0": ("small", "200", 0.05, 0, 0.0, 0, 1, 0),
    "small_c0.05_m1": ("small", "200", 0.05, 1, 0.0, 0, 1, 0),
    "small_c0.5_m2": ("small", "200", 0.5, 2, 0.0, 0, 1, 0),
    "small_c0.8_m2": ("small", "200", 0.8, 2, 0.0, 0, 1, 0),
    "small_c0.1_m3": ("small", "200", 0.1, 3, 0.0, 0, 1, 0),
    "small_c0.05_m4": ("small", "200", 0.05, 4, 0.0, 0, 1, 0),
    "small_c0.05_m5": ("small", "200", 0.05, 5, 0.0, 0, 1, 0),
    "small_id0.8": ("small", "200", 0.0, 0, 0.8, 0, 1, 0),
    "small_len9": ("small", "200", 0.0, 0, 0.0, 9, 1, 0),
    "small_noback": ("small", "200", 0.0, 0, 0.0, 0, 0, 0),
    "long_c0.5_m0": ("long", "1.000E-01", 0.5, 0, 0.0, 0, 1, 0),
    "long_c0.3_m1": ("long", "1.000E-01", 0.3, 1, 0.0, 0, 1, 0),
    "long_c0.7_m2": ("long", "1.000E-01", 0.7, 2, 0.0, 0, 1, 0),
    "long_c0.5_m3": ("long", "1.000E-01", 0.5, 3, 0.0, 0, 1, 0),
    "long_c0.5_m4": ("long", "1.000E-01", 0.5, 4, 0.0, 0, 1, 0),
    "long_c0.5_m5": ("long", "1.000E-01", 0.5, 5, 0.0, 0, 1, 0),
    "long_id0.6": ("long", "1.000E-01", 0.0, 0, 0.6, 0, 1, 0),
    "long_id0.3_mode1": ("long", "1.000E-01", 0.0, 0, 0.3, 0, 1, 1),
    "long_id0.3_mode2": ("long", "1.000E-01", 0.0, 0, 0.3, 0, 1, 2),
    "long_len40": ("long", "1.000E-01", 0.0, 0, 0.0, 40, 1, 0),
    "long_c0.6_m2_id0.5_len25": ("long", "1.000E-01", 0.6, 2, 0.5, 25, 1, 0),
    "long_c0.6_m2_id0.5_len25_noback": ("long", "1.000E-01", 0.6, 2, 0.5, 25, 0, 0),
}


# alignment modes without -a (Alignment::initSWMode, Alignment.cpp:174-196): 1 score only (no start positions, identity estimated
# from the score), 2 score + coverage (start positions, estimated identity), 0 picks by the thresholds; same tuple + alignment mode
MODE_CASES = {
    "small_mode0": ("small", "200", 0.0, 0, 0.0, 0, 0, 0, 0),
    "small_mode1": ("small", "200", 0.0, 0, 0.0, 0, 0, 0, 1),
    "small_mode2": ("small", "200", 0.0, 0, 0.0, 0, 0, 0, 2),
    "small_mode0_c0.3_m2": ("small", "200", 0.3, 2, 0.0, 0, 0, 0, 0),
    "small_mode0_c0.3_m2_id0.5": ("small", "200", 0.3, 2, 0.5, 0, 0, 0, 0),
    "small_mode1_id0.25_len12": ("small", "200", 0.0, 0, 0.25, 12, 0, 0, 1),
    "small_mode1_c0.3_m2": ("small", "200", 0.3, 2, 0.0, 0, 0, 0, 1),
    "long_mode1": ("long", "1.000E-01", 0.0, 0, 0.0, 0, 0, 0, 1),
    "long_mode2": ("long", "1.000E-01", 0.0, 0, 0.0, 0, 0, 0, 2),
    "long_mode2_c0.5_m0_id0.3": ("long", "1.000E-01", 0.5, 0, 0.3, 0, 0, 0, 2),
    "long_mode0_c0.6_m2": ("long", "1.000E-01", 0.6, 2, 0.0, 0, 0, 0, 0),
}


def aln_flags(evalue, cov, cov_mode, seq_id, aln_len, backtrace, id_mode, aln_mode=None):
    f = refbin.ALN_FLAGS.format(evalue=evalue, go=16, ge=2, threads=4)
    f = f.replace(" -c 0 --cov-mode 0 ", " -c %g --cov-mode %d " % (cov, cov_mode))
    f = f.replace("--min-seq-id 0 --min-aln-len 0", "--min-seq-id %g --min-aln-len %d" % (seq_id, aln_len))
    f = f.replace("--seq-id-mode 0", "--seq-id-mode %d" % id_mode)
    if not backtrace:
        f = f.replace("-a 1 --alignment-mode 2", "-a 0 --alignment-mode %d" % (3 if aln_mode is None else aln_mode))
    return f


def main():
    assert refbin.available(), "build the reference first: make -C oracle ref"
    B = refbin.REF_BIN
    out = os.path.join(GOLD, "criteria")
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="spgoldc_")
    n_seq = {}
    for name in ("small", "long"):
        w = os.path.join(work, name)
        os.makedirs(w)
        for n in ("t", "q"):
            with gzip.open(os.path.join(GOLD, name, n + ".fasta.gz"), "rb") as f, open(os.path.join(w, n + ".fasta"), "wb") as g:
                g.write(f.read())
            run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w, n, w, n))
        nT = sum(1 for _ in open(os.path.join(w, "tdb.index")))
        nQ = sum(1 for _ in open(os.path.join(w, "qdb.index")))
        n_seq[name] = (nT, nQ)
        # the all-pairs prefilter DBs of tools/gen_golden.py: every third target for the first 40 queries (small), all pairs (long)
        with open(os.path.join(w, "allpairs"), "wb") as f, open(os.path.join(w, "allpairs.index"), "w") as fi:
            pos = 0
            for q in range(nQ):
                if name == "small":
                    body = b"".join(b"%d\t0\t0\n" % t for t in range(0, nT, 3)) if q < 40 else b""
                else:
                    body = b"".join(b"%d\t0\t0\n" % t for t in range(nT))
                f.write(body + b"\0")
                fi.write("%d\t%d\t%d\n" % (q, pos, len(body) + 1))
                pos += len(body) + 1
        with open(os.path.join(w, "allpairs.dbtype"), "wb") as f:
            f.write((7).to_bytes(4, "little"))        # DBTYPE_PREFILTER_RES
    for tag, c in list(CASES.items()) + list(MODE_CASES.items()):
        name, ev, cov, mode, sid, alen, back, idm = c[:8]
        w = os.path.join(work, name)
        run("%s align %s/qdb %s/tdb %s/allpairs %s/aln_%s %s" % (B, w, w, w, w, tag, aln_flags(ev, cov, mode, sid, alen, back, idm,
                                                                                            c[8] if len(c) > 8 else None)))
        dump_db(os.path.join(w, "aln_" + tag), os.path.join(out, tag + ".txt.gz"))
    # prefilter lists under -c: hits whose lengths cannot reach the coverage are dropped after the max-seqs cut in the modes
    # 0 / 2 / 5 (Prefiltering.cpp:834-842); mode 1 leaves the list alone
    w = os.path.join(work, "long")
    for mode in (0, 1, 2, 5):
        tag = "pref_long_c0.5_m%d" % mode
        flags = refbin.PREF_FLAGS.format(maxseqs=300, mask=1, threads=4, sens="5.7").replace(" -c 0 --cov-mode 0 ", " -c 0.5 --cov-mode %d " % mode)
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w, w, w, tag, flags))
        dump_db(os.path.join(w, tag), os.path.join(out, tag + ".txt.gz"))
    # --exact-kmer-matching 1: every window looks up its own k-mer only (Prefiltering.cpp:177, QueryMatcher.cpp:221-231)
    for name in ("small", "long"):
        w = os.path.join(work, name)
        tag = "pref_%s_exact" % name
        flags = refbin.PREF_FLAGS.format(maxseqs=300, mask=1, threads=4, sens="5.7").replace("--exact-kmer-matching 0", "--exact-kmer-matching 1")
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w, w, w, tag, flags))
        dump_db(os.path.join(w, tag), os.path.join(out, tag + ".txt.gz"))
    # ... at -s 1.0 (k-mer threshold 154): the index of an exact-matching search is built WITHOUT the self-score pruning of its
    # k-mers (localKmerThr = 0, Prefiltering.cpp:509-512), so the lists hold hits the pruned index of pref_s1.0 cannot give
    w = os.path.join(work, "small")
    flags = refbin.PREF_FLAGS.format(maxseqs=300, mask=1, threads=4, sens="1.0").replace("--exact-kmer-matching 0", "--exact-kmer-matching 1")
    run("%s prefilter %s/qdb %s/tdb %s/pref_small_exact_s1.0 %s" % (B, w, w, w, flags))
    dump_db(os.path.join(w, "pref_small_exact_s1.0"), os.path.join(out, "pref_small_exact_s1.0.txt.gz"))
    # --mask-lower-case 1: lower-case letters of the target DB become X in the index, after tantan (IndexBuilder.cpp:151-160).
    # Targets = the `small` set with seeded lower-case stretches (createdb keeps the case).
    import numpy as np
    rng = np.random.default_rng(41)
    w = os.path.join(work, "small")
    with open(os.path.join(w, "t.fasta")) as f:
        recs = f.read().split(">")[1:]
    with open(os.path.join(w, "t_lower.fasta"), "w") as f:
        for r in recs:
            head, seq = r.split("\n", 1)
            seq = list(seq.replace("\n", ""))
            for _ in range(int(rng.integers(0, 4))):
                a = int(rng.integers(0, len(seq)))
                b = min(len(seq), a + int(rng.integers(1, 25)))
                seq[a:b] = [c.lower() for c in seq[a:b]]
            f.write(">%s\n%s\n" % (head, "".join(seq)))
    with open(os.path.join(w, "t_lower.fasta"), "rb") as f, gzip.open(os.path.join(out, "t_lower.fasta.gz"), "wb", 9) as g:
        g.write(f.read())
    run("%s createdb %s/t_lower.fasta %s/tdb_lower --dbtype 1 --shuffle 0" % (B, w, w))
    for mask in (0, 1):
        tag = "pref_small_lower_mask%d" % mask
        flags = refbin.PREF_FLAGS.format(maxseqs=300, mask=mask, threads=4, sens="5.7").replace("--mask-lower-case 0", "--mask-lower-case 1")
        run("%s prefilter %s/qdb %s/tdb_lower %s/%s %s" % (B, w, w, w, tag, flags))
        dump_db(os.path.join(w, tag), os.path.join(out, tag + ".txt.gz"))
    # --split N --split-mode 0: MMseqs2's own (lossy) target split -- targets cut by decomposeDomainByAminoAcid, every split searched
    # with max-seqs / N + 4 sqrt(max-seqs / N), lists concatenated and sorted (Prefiltering.cpp:343-346,364-481,712-733)
    w = os.path.join(work, "family")
    os.makedirs(w)
    for n in ("t", "q"):
        with gzip.open(os.path.join(GOLD, "family", n + ".fasta.gz"), "rb") as f, open(os.path.join(w, n + ".fasta"), "wb") as g:
            g.write(f.read())
        run("%s createdb %s/%s.fasta %s/%sdb --dbtype 1 --shuffle 0" % (B, w, n, w, n))
    for name, split, maxseqs in (("family", 3, 20), ("small", 2, 100)):
        w = os.path.join(work, name)
        tag = "pref_%s_split%d_max%d" % (name, split, maxseqs)
        flags = refbin.PREF_FLAGS.format(maxseqs=maxseqs, mask=1, threads=4, sens="5.7").replace("--split 0 --split-mode 2",
                                                                                               "--split %d --split-mode 0" % split)
        run("%s prefilter %s/qdb %s/tdb %s/%s %s" % (B, w, w, w, tag, flags))
        dump_db(os.path.join(w, tag), os.path.join(out, tag + ".txt.gz"))
    # --max-accept / --max-rejected (Alignment.cpp:353-404): the prefilter list is walked in order until that many pairs passed /
    # that many pairs in a row failed.  `family` set, its --max-seqs 300 prefilter lists (tests/golden/family/pref_max300).
    w = os.path.join(work, "family")
    ent = {}
    with gzip.open(os.path.join(GOLD, "family", "pref_max300.txt.gz"), "rt") as f:
        for line in f:
            k, row = line.rstrip("\n").split("\t", 1)
            ent.setdefault(int(k), []).append(row)
    nQ = sum(1 for _ in open(os.path.join(w, "qdb.index")))
    with open(os.path.join(w, "pref300"), "wb") as f, open(os.path.join(w, "pref300.index"), "w") as fi:
        pos = 0
        for q in range(nQ):
            body = "".join(r + "\n" for r in ent.get(q, [])).encode()
            f.write(body + b"\0")
            fi.write("%d\t%d\t%d\n" % (q, pos, len(body) + 1))
            pos += len(body) + 1
    with open(os.path.join(w, "pref300.dbtype"), "wb") as f:
        f.write((7).to_bytes(4, "little"))
    limits = {}
    for tag, acc, rej, sid, ev in (("family_acc2", 2, 2147483647, 0.0, "200"), ("family_rej1_id0.93", 2147483647, 1, 0.93, "200"),
                                   ("family_rej3_id0.95", 2147483647, 3, 0.95, "200"), ("family_acc3_rej2_id0.9", 3, 2, 0.9, "200"),
                                   ("family_rej2_e1e-20", 2147483647, 2, 0.0, "1.000E-20")):
        f = aln_flags(ev, 0.0, 0, sid, 0, 1, 0).replace("--max-rejected 2147483647 --max-accept 2147483647",
                                                       "--max-rejected %d --max-accept %d" % (rej, acc))
        run("%s align %s/qdb %s/tdb %s/pref300 %s/aln_%s %s" % (B, w, w, w, w, tag, f))
        dump_db(os.path.join(w, "aln_" + tag), os.path.join(out, tag + ".txt.gz"))
        limits[tag] = dict(max_accept=acc, max_rejected=rej, min_seq_id=sid, evalue=float(ev))
    # --alt-ali N (Alignment::computeAlternativeAlignment, Alignment.cpp:576-611)
    alts = {}
    for tag, name, n, ev, cov, mode in (("small_alt2", "small", 2, "200", 0.0, 0), ("long_alt3", "long", 3, "1.000E-01", 0.0, 0),
                                        ("long_alt2_c0.5_m2", "long", 2, "1.000E-01", 0.5, 2)):
        w = os.path.join(work, name)
        f = aln_flags(ev, cov, mode, 0.0, 0, 1, 0).replace("--alt-ali 0", "--alt-ali %d" % n)
        run("%s align %s/qdb %s/tdb %s/allpairs %s/aln_%s %s" % (B, w, w, w, w, tag, f))
        dump_db(os.path.join(w, "aln_" + tag), os.path.join(out, tag + ".txt.gz"))
        alts[tag] = dict(set=name, alt_ali=n, evalue=float(ev), cov=cov, cov_mode=mode)
    with open(os.path.join(out, "alt_ali.json"), "w") as f:
        json.dump(alts, f, indent=1)
    # --realign 1 (Alignment.cpp:47-56,415-447): score-only first pass, second pass with the realign matrix
    realign = {}
    for tag, name, bias, ev, cov, mode, mx in (("small_realign_b0.6", "small", -0.6, "200", 0.0, 0, 2147483647),
                                               ("long_realign_b1.0", "long", -1.0, "1.000E-01", 0.0, 0, 2147483647),
                                               ("long_realign_b1.0_c0.5_m2_max3", "long", -1.0, "1.000E-01", 0.5, 2, 3),
                                               ("long_realign_b0.2", "long", -0.2, "1.000E-01", 0.3, 0, 2147483647)):
        w = os.path.join(work, name)
        f = aln_flags(ev, cov, mode, 0.0, 0, 1, 0).replace("--realign 0 --realign-score-bias -0.2 --realign-max-seqs 2147483647",
                                                          "--realign 1 --realign-score-bias %g --realign-max-seqs %d" % (bias, mx))
        assert "--realign 1" in f
        run("%s align %s/qdb %s/tdb %s/allpairs %s/aln_%s %s" % (B, w, w, w, w, tag, f))
        dump_db(os.path.join(w, "aln_" + tag), os.path.join(out, tag + ".txt.gz"))
        realign[tag] = dict(set=name, realign_score_bias=bias, evalue=float(ev), cov=cov, cov_mode=mode, realign_max_seqs=mx)
    with open(os.path.join(out, "realign.json"), "w") as f:
        json.dump(realign, f, indent=1)
    with open(os.path.join(out, "limits.json"), "w") as f:
        json.dump(limits, f, indent=1)
    with open(os.path.join(out, "cases.json"), "w") as f:
        json.dump({t: dict(set=c[0], evalue=float(c[1]), cov=c[2], cov_mode=c[3], min_seq_id=c[4], min_aln_len=c[5], backtrace=c[6],
                        seq_id_mode=c[7], alignment_mode=(c[8] if len(c) > 8 else 3))
                   for t, c in list(CASES.items()) + list(MODE_CASES.items())}, f, indent=1)
    shutil.rmtree(work)
    print("criteria fixtures written")


if __name__ == "__main__":
    main()
