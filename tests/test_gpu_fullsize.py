"""Full-size GPU tests at BASELINE.json's sizes (config 2: 10 000 spacers vs ~0.5 M ORFs; config 4: 1 M spacers).

The oracle cannot run these sizes whole in seconds, so the CUDA path is checked
  * bit-exactly against the oracle on a seeded SAMPLE of the batch (the oracle indexes the full target set once), and
  * through size-independent properties on the WHOLE batch: batch-split independence (the result of a query does not
    depend on what else is in the batch, nor on the batch order), sortedness of the hit lists
    (QueryMatcher.h:40-50) and of the alignment rows (Matcher.h:156-167), the max-seqs / min-ungapped-score / E-value
    bounds, coordinate bounds and CIGAR consistency (Matcher.cpp:61-187)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
import common as Cm
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

EVAL_RTOL = 1e-6


@pytest.fixture(scope="module")
def world():
    """config-2 workload: targets (masked for the index, raw for alignment), frames, one resident searcher"""
    from spacepharer_b200 import api, synth
    tres, toff, _ = synth.make_targets(n_genomes=8000)
    qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
    s = api.Searcher(device=0)
    masked, _ = api.tantan_mask(s.seed, tres, toff)
    s.build_index(masked, toff)
    s.set_align_targets(tres, toff)
    w = dict(s=s, tres=tres, toff=toff, masked=masked, qres=qres, qoff=qoff, db_res=int(toff[-1]))
    yield w
    s.close()


def _take(qres, qoff, idx):
    """sub-batch of the queries idx (in that order)"""
    lens = (qoff[1:] - qoff[:-1]).astype(np.int64)[idx]
    off = np.zeros(len(idx) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    res = np.concatenate([qres[int(qoff[i]):int(qoff[i + 1])] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
    return res, off


def _rows_by_query(off, rows, cig):
    """per query: list of comparable tuples (CIGAR bytes included)"""
    out = []
    for i in range(len(off) - 1):
        rr = rows[int(off[i]):int(off[i + 1])]
        out.append([(int(r["target"]), int(r["score"]), int(r["bits"]), float(r["eval"]), int(r["q_start"]), int(r["q_end"]),
                     int(r["t_start"]), int(r["t_end"]), int(r["n_ident"]), int(r["aln_len"]),
                     cig[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["cigar_len"])]) for r in rr])
    return out


def _cigar_concat(ops, rows):
    """CIGAR ops of the rows back to back (the library's buffer may hold slack between rows)"""
    co, cl = rows["cigar_off"].astype(np.int64), rows["cigar_len"].astype(np.int64)
    if len(co) == 0:
        return ops[:0]
    start = np.concatenate([[0], np.cumsum(cl)[:-1]])
    return ops[np.repeat(co - start, cl) + np.arange(int(cl.sum()))]


@pytest.fixture(scope="module")
def full(world):
    s = world["s"]
    off, rows, cig, st, pref = s.search(world["qres"], world["qoff"], world["db_res"], fetch_prefilter=True)
    return dict(off=off, rows=rows, cig=cig, st=st, poff=pref[0], hits=pref[1])


def test_config2_prefilter_properties(world, full):
    poff, hits = full["poff"], full["hits"]
    Q = len(world["qoff"]) - 1
    T = len(world["toff"]) - 1
    assert len(poff) == Q + 1 and int(poff[-1]) == len(hits) and (np.diff(poff.astype(np.int64)) >= 0).all()
    cnt = np.diff(poff.astype(np.int64))
    assert cnt.max() <= 300                                   # --max-seqs
    assert (hits["score"] >= 15).all()                        # --min-ungapped-score
    assert (hits["id"] < T).all()
    qlen = (world["qoff"][1:] - world["qoff"][:-1]).astype(np.int64)
    assert (cnt[qlen < 9] == 0).all()                         # fewer than two k-mer windows: nothing can be counted (quirk A)
    # order inside a list: score desc, id asc (hit_t::compareHitsByScoreAndId); a target appears once per query
    qid = np.repeat(np.arange(Q), cnt)
    same = qid[1:] == qid[:-1]
    sc, ids = hits["score"].astype(np.int64), hits["id"].astype(np.int64)
    ok = (sc[:-1] > sc[1:]) | ((sc[:-1] == sc[1:]) & (ids[:-1] < ids[1:]))
    assert (ok | ~same).all()
    key = qid.astype(np.int64) * T + ids
    assert len(np.unique(key)) == len(key)
    # the diagonal of a hit lies inside the DP matrix of (query, target)
    tlen = (world["toff"][1:] - world["toff"][:-1]).astype(np.int64)
    d = hits["diag"].astype(np.int64)
    assert ((d < qlen[qid]) & (-d < tlen[ids])).all()


def test_config2_align_properties(world, full):
    off, rows, cig = full["off"], full["rows"], full["cig"]
    Q = len(world["qoff"]) - 1
    qlen = (world["qoff"][1:] - world["qoff"][:-1]).astype(np.int64)
    tlen = (world["toff"][1:] - world["toff"][:-1]).astype(np.int64)
    cnt = np.diff(off.astype(np.int64))
    qid = np.repeat(np.arange(Q), cnt)
    assert (rows["query"] == qid).all()
    assert (rows["eval"] <= 200.0).all() and (rows["eval"] >= 0).all()
    assert (rows["score"] > 0).all()
    tq, tt = qlen[qid], tlen[rows["target"]]
    assert ((rows["q_start"] >= 0) & (rows["q_start"] <= rows["q_end"]) & (rows["q_end"] < tq)).all()
    assert ((rows["t_start"] >= 0) & (rows["t_start"] <= rows["t_end"]) & (rows["t_end"] < tt)).all()
    # every row's target was a prefilter hit of that query
    T = len(tlen)
    pcnt = np.diff(full["poff"].astype(np.int64))
    pkey = np.repeat(np.arange(Q), pcnt).astype(np.int64) * T + full["hits"]["id"].astype(np.int64)
    rkey = qid.astype(np.int64) * T + rows["target"].astype(np.int64)
    assert np.isin(rkey, pkey).all() and len(np.unique(rkey)) == len(rkey)
    # order: E-value asc, score desc, target length asc, key asc (Matcher::compareHits)
    same = qid[1:] == qid[:-1]
    e, b, tl, k = rows["eval"], rows["bits"].astype(np.int64), tt, rows["target"].astype(np.int64)
    lt = (e[:-1] < e[1:]) | ((e[:-1] == e[1:]) & ((b[:-1] > b[1:]) | ((b[:-1] == b[1:]) & ((tl[:-1] < tl[1:]) | ((tl[:-1] == tl[1:]) & (k[:-1] < k[1:]))))))
    assert (lt | ~same).all()
    # CIGAR consistency: ops in {M, I, D}; M + I span the query interval, M + D the target interval; aln_len = #ops;
    # identities never exceed the matched columns
    ops = np.frombuffer(cig, dtype=np.uint8)
    co, cl = rows["cigar_off"].astype(np.int64), rows["cigar_len"].astype(np.int64)
    assert (co + cl <= len(ops)).all() and len(np.unique(co)) == len(co)      # (the buffer may hold slack between rows)
    cs = lambda ch: np.concatenate([[0], np.cumsum(ops == ord(ch))])   # noqa: E731
    cM, cI, cD = cs("M"), cs("I"), cs("D")
    nM, nI, nD = cM[co + cl] - cM[co], cI[co + cl] - cI[co], cD[co + cl] - cD[co]
    assert (nM + nI + nD == cl).all() and (cl == rows["aln_len"]).all()
    # the reference's CIGAR letters: I consumes the query?  fix the convention from the data of one gapped row, then demand it for all
    qspan = rows["q_end"].astype(np.int64) - rows["q_start"] + 1
    tspan = rows["t_end"].astype(np.int64) - rows["t_start"] + 1
    a = ((nM + nI == qspan) & (nM + nD == tspan)).all()
    b2 = ((nM + nD == qspan) & (nM + nI == tspan)).all()
    assert a or b2
    assert (rows["n_ident"] <= nM).all() and (rows["n_ident"] >= 0).all()
    assert np.allclose(rows["seq_id"], rows["n_ident"] / np.maximum(rows["aln_len"], 1), rtol=0, atol=1e-6)


def test_config2_sample_vs_oracle(world, full):
    """2 000 seeded queries of the full batch, against the oracle holding the FULL target index"""
    m80, m40u, m40 = Cm.mats()
    rng = np.random.default_rng(2024)
    Q = len(world["qoff"]) - 1
    idx = np.sort(rng.choice(Q, 2000, replace=False))
    sres, soff = _take(world["qres"], world["qoff"], idx)
    pf = O.Prefilter(world["masked"], world["toff"], m80, m40u)
    ooff, ohits = pf.batch(sres, soff)
    pf.close()
    poff, hits = full["poff"], full["hits"]
    for j, qi in enumerate(idx):
        g = hits[int(poff[qi]):int(poff[qi + 1])]
        o = ohits[int(ooff[j]):int(ooff[j + 1])]
        assert len(g) == len(o) and (g["id"] == o["id"]).all() and (g["score"] == o["score"]).all() \
            and (g["diag"] == o["diag"]).all(), "prefilter list of query %d differs" % qi
    # alignment of the sample's pairs
    p = O.align_params(world["db_res"])
    orows = O.align_batch(m40, p, sres, soff, world["tres"], world["toff"], ooff, ohits["id"])
    off, rows = full["off"], full["rows"]
    nrows = 0
    for j, qi in enumerate(idx):
        ids = ohits["id"][int(ooff[j]):int(ooff[j + 1])]
        oo = orows[int(ooff[j]):int(ooff[j + 1])]
        exp = {int(t): r for t, r in zip(ids, oo) if r["ok"] == 1}
        got = rows[int(off[qi]):int(off[qi + 1])]
        assert sorted(exp) == sorted(int(t) for t in got["target"]), "accepted targets of query %d differ" % qi
        for r in got:
            e = exp[int(r["target"])]
            assert (int(r["score"]), int(r["bits"]), int(r["q_start"]), int(r["q_end"]), int(r["t_start"]), int(r["t_end"]),
                    int(r["n_ident"]), int(r["aln_len"])) == \
                   (int(e["score"]), int(e["bits"]), int(e["q_start"]), int(e["q_end"]), int(e["t_start"]), int(e["t_end"]),
                    int(e["n_ident"]), int(e["aln_len"])), "row (%d, %d) differs" % (qi, int(r["target"]))
            assert abs(float(r["eval"]) - float(e["eval"])) <= EVAL_RTOL * abs(float(e["eval"]))
            nrows += 1
    assert nrows > 3000


def test_config2_batch_split_and_order_independence(world, full):
    """halves searched separately, and the whole batch in reversed order, give the same per-query rows"""
    s = world["s"]
    Q = len(world["qoff"]) - 1
    ref = _rows_by_query(full["off"], full["rows"], full["cig"])
    for idx in (np.arange(0, Q // 2), np.arange(Q // 2, Q), np.arange(Q)[::-1]):
        sres, soff = _take(world["qres"], world["qoff"], idx)
        off, rows, cig, _, _ = s.search(sres, soff, world["db_res"])
        got = _rows_by_query(off, rows, cig)
        bad = [int(qi) for j, qi in enumerate(idx) if got[j] != ref[qi]]
        assert not bad, bad[:5]


def test_config4_one_million_spacers_in_one_batch(world):
    """BASELINE config 4's query side on one GPU: 1 M spacers (~3.8 M frames) as ONE batch.  Checked against the same
    frames searched in 16 chunks (arrays equal after re-basing), plus the size-independent bounds."""
    from spacepharer_b200 import synth
    s = world["s"]
    qres, qoff, _ = synth.make_query_frames(1000000, world["tres"], world["toff"])
    Q = len(qoff) - 1
    assert Q > 3000000
    off, rows, cig, st, _ = s.search(qres, qoff, world["db_res"])
    assert len(off) == Q + 1 and int(off[-1]) == len(rows)
    assert st["prefilter"]["n_slow_path"] < 0.02 * Q
    assert (rows["eval"] <= 200.0).all()
    cnt = np.diff(off.astype(np.int64))
    assert (rows["query"] == np.repeat(np.arange(Q), cnt)).all()
    cigs = np.frombuffer(cig, dtype=np.uint8)
    fields = ("target", "score", "bits", "eval", "q_start", "q_end", "t_start", "t_end", "n_ident", "aln_len", "cigar_len")
    bounds = np.linspace(0, Q, 17).astype(np.int64)
    for c in range(16):
        lo, hi = int(bounds[c]), int(bounds[c + 1])
        r0, r1 = int(qoff[lo]), int(qoff[hi])
        coff, crows, ccig, _, _ = s.search(qres[r0:r1], (qoff[lo:hi + 1] - qoff[lo]).astype(np.uint64), world["db_res"])
        assert (coff.astype(np.int64) == off[lo:hi + 1].astype(np.int64) - int(off[lo])).all(), "row counts differ in chunk %d" % c
        big = rows[int(off[lo]):int(off[hi])]
        for f in fields:
            assert (big[f] == crows[f]).all(), "%s differs in chunk %d" % (f, c)
        assert (crows["query"] == big["query"] - lo).all()
        assert (_cigar_concat(cigs, big) == _cigar_concat(np.frombuffer(ccig, dtype=np.uint8), crows)).all(), "CIGARs differ in chunk %d" % c


def test_config2_byte_identical_to_the_reference_binary(tmp_path):
    """BASELINE config 2 at full size -- the workload bench.py measures -- through the spb200 binary: the prefilter DB and the
    alignment DB are byte-identical to the REFERENCE binary's (tests/golden/config2_digest.json holds the SHA-256 of the
    reference's DBs in canonical form, written by tools/gen_golden_config2.py; the inputs are regenerated from fixed seeds)"""
    import hashlib
    import json
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from gen_golden_config2 import db_digest, write_config2
    from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS
    want = json.load(open(os.path.join(GOLDEN, "config2_digest.json")))
    w = str(tmp_path)
    nq, nt = write_config2(w)
    assert (nq, nt) == (want["n_queries"], want["n_targets"])
    for args in (["prefilter", w + "/queries", w + "/targets", w + "/pref"] + PREF_FLAGS,
                 ["align", w + "/queries", w + "/targets", w + "/pref", w + "/aln"] + ALN_FLAGS):
        r = subprocess.run([BIN] + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert db_digest(w + "/pref") == (want["pref_sha256"], want["pref_rows"])
    assert db_digest(w + "/aln") == (want["aln_sha256"], want["aln_rows"])
    # the fused search (hit lists stay in HBM) writes the same alignment DB
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    r = subprocess.run([BIN, "search", w + "/queries", w + "/targets", w + "/fused"] + merged + ["--batch", "10000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert db_digest(w + "/fused") == (want["aln_sha256"], want["aln_rows"])
    # multi-GPU in the host binary: --gpus 2 = two contexts with one k-mer range of the table each (index split, the other
    # range read through peer-mapped pointers), query batches handed out to both.  On a one-GPU box the two contexts share
    # the device (--oversubscribe 1: same code path, same kernels); with two GPUs the ranges sit in different HBMs.
    import torch
    extra = [] if torch.cuda.device_count() >= 2 else ["--oversubscribe", "1"]
    r = subprocess.run([BIN, "search", w + "/queries", w + "/targets", w + "/fused2"] + merged + ["--gpus", "2"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sharded over 2 GPUs" in r.stdout
    assert db_digest(w + "/fused2") == (want["aln_sha256"], want["aln_rows"])
    # the flag surface beyond the predictmatch defaults at the same size (tools/gen_golden_config2_flags.py): acceptance criteria,
    # rows without backtrace, --seq-id-mode, exact k-mer matching, the reference's own target split
    flags = json.load(open(os.path.join(GOLDEN, "config2_flags_digest.json")))
    assert (flags["n_queries"], flags["n_targets"]) == (nq, nt)
    for name, c in sorted(flags["cases"].items()):
        argv = list(ALN_FLAGS if c["module"] == "align" else PREF_FLAGS)
        for k, v in c["flags"].items():
            argv[argv.index(k) + 1] = v
        dbs = [w + "/queries", w + "/targets"] + ([w + "/pref"] if c["module"] == "align" else []) + [w + "/" + name]
        r = subprocess.run([BIN, c["module"]] + dbs + argv, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert db_digest(w + "/" + name) == (c["sha256"], c["rows"]), name


def test_config4_one_million_spacers_byte_identical_to_the_reference_binary(tmp_path):
    """BASELINE config 4 on one GPU: 1 000 000 synthetic spacers (3 789 911 frames) against the 495 567-ORF set through the
    spb200 binary -- 42.8 M prefilter hits and 13.9 M alignment rows, byte-identical to the reference binary's DBs
    (tests/golden/config4_digest.json).  The fused search runs four batches with the formatter thread overlapping the GPU."""
    import json
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from gen_golden_config2 import db_digest, write_config2
    from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS
    want = json.load(open(os.path.join(GOLDEN, "config4_digest.json")))
    w = str(tmp_path)
    nq, nt = write_config2(w, want["spacers"])
    assert (nq, nt) == (want["n_queries"], want["n_targets"])
    threads = str(min(16, os.cpu_count() or 1))
    pf = [threads if PREF_FLAGS[i - 1] == "--threads" else x for i, x in enumerate(PREF_FLAGS)]
    r = subprocess.run([BIN, "prefilter", w + "/queries", w + "/targets", w + "/pref"] + pf, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert db_digest(w + "/pref") == (want["pref_sha256"], want["pref_rows"])
    os.remove(w + "/pref")
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], threads if lst[i] == "--threads" else lst[i + 1]]
    r = subprocess.run([BIN, "search", w + "/queries", w + "/targets", w + "/aln"] + merged, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert db_digest(w + "/aln") == (want["aln_sha256"], want["aln_rows"])
