"""Exhaustive search (BASELINE config 5): all query frames x all targets through spb_pairs_exhaustive + the alignment
stage, against (a) the reference's `align` rows on all-pairs prefilter DBs (golden fixtures), (b) the CPU oracle on
seeded sets, (c) the list path of this library.  Rows bit-exact, E-values within 1e-6 relative."""
import numpy as np
import pytest

from oracle import pyoracle as O
import common as Cm

pytestmark = pytest.mark.gpu

EVAL_RTOL = 1e-6


@pytest.fixture(scope="module")
def searcher():
    from spacepharer_b200 import api
    s = api.Searcher(device=0)
    yield s
    s.close()


def _exhaustive_rows(searcher, T, Q, db_res, tkeys=None, **kw):
    from spacepharer_b200 import api
    tres, toff = O.concat(T)
    searcher.set_align_targets(tres, toff, tkeys)
    qres, qoff = O.concat(Q)
    off, rows, cig, st = searcher.align_exhaustive(qres, qoff, db_res, **kw)
    qlens = [len(q) for q in Q]
    tlens = [len(t) for t in T]
    txt = [api.format_align_rows(rows[int(off[i]):int(off[i + 1])], cig, qlens, tlens, tkeys) for i in range(len(Q))]
    return txt, rows, off, st


def test_exhaustive_long_golden(searcher):
    """queries of 17..150 aa take every target unfiltered: rows equal the reference's all-pairs `align` output"""
    T, Q = Cm.load_fasta_set("long")
    txt, _, _, st = _exhaustive_rows(searcher, T, Q, Cm.db_res_count(T), evalue=0.1)
    ref = Cm.dump("long", "aln_allpairs_e0.1.txt.gz")
    bad = [(i, txt[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if txt[i] != ref.get(i, [])]
    assert not bad, bad[:4]
    ex = st["exhaustive"]
    assert ex["n_pairs"] == ex["n_kept"] + ex["n_unfiltered"]
    assert ex["n_unfiltered"] == sum(1 for q in Q if len(q) > 16) * len(T)


def test_exhaustive_small_golden_and_list_path(searcher):
    """spacer-length queries go through the streaming register kernel; its rows equal the list path over all pairs,
    and (restricted to the fixture's pair subset) the reference's rows"""
    T, Q = Cm.load_fasta_set("small")
    db_res = Cm.db_res_count(T)
    txt, rows, off, st = _exhaustive_rows(searcher, T, Q, db_res)
    ex = st["exhaustive"]
    nfast = sum(1 for q in Q if 1 <= len(q) <= 16)
    assert ex["n_scored"] == nfast * len(T)
    assert ex["n_cells"] == sum(len(q) for q in Q if 1 <= len(q) <= 16) * db_res
    assert 0 < ex["n_kept"] < ex["n_scored"]
    # (c) list path over the full cross product
    from spacepharer_b200 import api
    qres, qoff = O.concat(Q)
    poff = (np.arange(len(Q) + 1, dtype=np.uint64) * np.uint64(len(T)))
    pt = np.tile(np.arange(len(T), dtype=np.uint32), len(Q))
    off2, rows2, cig2, _ = searcher.align(qres, qoff, poff, pt, db_res)
    txt2 = api.format_align_rows(rows2, cig2, [len(q) for q in Q], [len(t) for t in T])
    assert (off == off2).all()
    assert [r for t in txt for r in t] == txt2
    # (a) the fixture: queries 0..39 against every third target
    ref = Cm.dump("small", "aln_allpairs.txt.gz")
    for i in range(40):
        sub = [r for r in txt[i] if int(r.split("\t")[0]) % 3 == 0]
        assert sub == ref.get(i, []), i


@pytest.mark.parametrize("evalue", [200.0, 1.0])
def test_exhaustive_random_vs_oracle(searcher, evalue):
    """seeded synthetic set, all lengths 1..20 (every row-count instantiation and the long-query path): the accepted set
    and every field against the oracle aligned over the full cross product"""
    from spacepharer_b200 import synth
    rng = np.random.default_rng(77)
    tres, toff, _ = synth.make_targets(n_genomes=3, seed=9)
    T = [tres[int(toff[i]):int(toff[i + 1])] for i in range(len(toff) - 1)][:150]
    Q = []
    for L in list(range(1, 21)) * 8:
        if rng.random() < 0.5:
            t = T[int(rng.integers(0, len(T)))]
            s = int(rng.integers(0, max(1, len(t) - L)))
            q = np.array(t[s:s + L], dtype=np.uint8)
            if len(q) < L:
                q = np.concatenate([q, rng.integers(0, 20, L - len(q)).astype(np.uint8)])
            for _ in range(int(rng.integers(0, 3))):
                q[int(rng.integers(0, L))] = int(rng.integers(0, 20))
        else:
            q = rng.integers(0, 20, L).astype(np.uint8)
        Q.append(q)
    db_res = 53_000_000     # whole-DB residue count of the named workload: the threshold the bench runs at
    _, m40u, m40 = Cm.mats()
    txt, raw, off, st = _exhaustive_rows(searcher, T, Q, db_res, evalue=evalue)
    p = O.align_params(db_res, eval_thr=evalue)
    n = 0
    for i, q in enumerate(Q):
        exp = {}
        for t in range(len(T)):
            a = O.align_pair(m40, p, q, T[t])
            if a.ok == 1:
                exp[t] = a
        got = {int(r["target"]): r for r in raw[int(off[i]):int(off[i + 1])]}
        assert set(got) == set(exp), (i, len(q), sorted(got)[:5], sorted(exp)[:5])
        for t, a in exp.items():
            r = got[t]
            assert (int(r["score"]), int(r["bits"]), int(r["q_start"]), int(r["q_end"]), int(r["t_start"]), int(r["t_end"]),
                    int(r["n_ident"]), int(r["aln_len"])) == (a.score, a.bits, a.qStart, a.qEnd, a.tStart, a.tEnd, a.nIdent, a.alnLen)
            assert abs(float(r["eval"]) - a.eval) <= EVAL_RTOL * abs(a.eval)
            n += 1
    assert n > 50


def test_exhaustive_edge_cases(searcher):
    """empty batch, empty query, one target, query batch smaller than a tile"""
    T, Q = Cm.load_fasta_set("small")
    db_res = Cm.db_res_count(T)
    txt, rows, off, st = _exhaustive_rows(searcher, T[:1], [Q[0], np.zeros(0, np.uint8), Q[1]], db_res)
    assert len(off) == 4 and off[2] == off[1]
    tres, toff = O.concat(T)
    searcher.set_align_targets(tres, toff)
    off, rows, cig, st = searcher.align_exhaustive(np.zeros(0, np.uint8), np.zeros(1, np.uint64), db_res)
    assert len(off) == 1 and len(rows) == 0


def test_exhaustive_superset_of_prefilter_path():
    """size-independent property at the bench's target DB: every row of the prefilter+align path is a row of the
    exhaustive path (same fields), and rows only found exhaustively are pairs the k-mer stage did not propose"""
    from spacepharer_b200 import api, synth
    tres, toff, _ = synth.make_targets(n_genomes=400, seed=20180901)
    qres, qoff, _ = synth.make_query_frames(300, tres, toff, seed=5)
    s = api.Searcher(device=0)
    try:
        masked, _ = api.tantan_mask(s.seed, tres, toff)
        s.build_index(masked, toff)
        s.set_align_targets(tres, toff)
        db_res = int(toff[-1])
        off, rows, cig, st, pref = s.search(qres, qoff, db_res, fetch_prefilter=True)
        eoff, erows, ecig, est = s.align_exhaustive(qres, qoff, db_res)
    finally:
        s.close()
    fields = ("score", "bits", "q_start", "q_end", "t_start", "t_end", "n_ident", "aln_len", "eval")
    poff, phits = pref[0], pref[1]
    extra = 0
    for i in range(len(qoff) - 1):
        a = {int(r["target"]): tuple(r[f] for f in fields) for r in rows[int(off[i]):int(off[i + 1])]}
        b = {int(r["target"]): tuple(r[f] for f in fields) for r in erows[int(eoff[i]):int(eoff[i + 1])]}
        assert set(a) <= set(b), i
        assert all(a[t] == b[t] for t in a), i
        proposed = set(int(x) for x in phits["id"][int(poff[i]):int(poff[i + 1])])
        assert not (set(b) - set(a)) & proposed, i
        extra += len(set(b) - set(a))
    assert est["exhaustive"]["n_kept"] >= len(erows)
    assert extra > 0


def test_exhaustive_search_at_scale_matches_reference_digest(tmp_path):
    """BASELINE config 5 at medium scale: 593 frames x 62 177 ORFs = 36.9 M pairs, no k-mer stage.  The reference side is its `align`
    over an all-pairs prefilter DB; `spb200 search --exhaustive-search 1` writes the same alignment DB byte for byte (68 005 rows;
    tests/golden/config5_digest.json, tools/gen_golden_config5.py)"""
    import json
    import os
    import subprocess
    import sys
    from conftest import GOLDEN, ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from gen_golden_config2 import db_digest
    from gen_golden_config5 import write_config5
    from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS
    want = json.load(open(os.path.join(GOLDEN, "config5_digest.json")))
    w = str(tmp_path)
    assert write_config5(w, **want["shape"]) == (want["n_queries"], want["n_targets"])
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    r = subprocess.run([BIN, "search", w + "/queries", w + "/targets", w + "/aln", "--exhaustive-search", "1"] + merged, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Exhaustive search" in r.stdout
    assert db_digest(w + "/aln") == (want["aln_sha256"], want["aln_rows"])
