"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle and the golden fixtures.
Bit-exact for hit lists, scores, coordinates and CIGARs; E-values within 1e-6 relative (BASELINE.json)."""
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


def _product_mask(searcher, T):
    from spacepharer_b200 import api
    tres, toff = O.concat(T)
    m, _ = api.tantan_mask(searcher.seed, tres, toff)
    return m, toff


def test_matrices_match_oracle(searcher):
    m80, m40u, m40 = Cm.mats()
    assert (searcher.seed.sub == O.matrix_sub(m80)).all()
    assert (searcher.ungapped.sub == O.matrix_sub(m40u)).all()
    assert (searcher.align_mat.sub == O.matrix_sub(m40)).all()


def test_index_matches_oracle(searcher):
    T, _ = Cm.load_fasta_set("small")
    tres, toff = _product_mask(searcher, T)
    m80, m40u, _ = Cm.mats()
    st = searcher.build_index(tres, toff)
    pf = O.Prefilter(tres, toff, m80, m40u)
    k, i, p = pf.postings()
    gk, gi, gp = searcher.export_index()
    assert st["n_postings"] == len(k)
    assert (gk == k).all() and (gi == i).all() and (gp == p).all()
    pf.close()


def test_similar_kmers_order(searcher):
    """KmerGenerator parity: same list, same order (the order fixes the hit stream)"""
    T, Q = Cm.load_fasta_set("small")
    tres, toff = _product_mask(searcher, T)
    m80, m40u, _ = Cm.mats()
    searcher.build_index(tres, toff)
    pf = O.Prefilter(tres, toff, m80, m40u)
    rng = np.random.default_rng(3)
    for thr in (90, 100, 112, 120, 135, 160):
        for _ in range(6):
            k6 = rng.integers(0, 20, 6).astype(np.uint8)
            a = pf.similar(k6, thr)
            b = searcher.similar_kmers(k6, thr)
            assert len(a) == len(b) and (a == b).all(), (thr, k6)
    pf.close()


def _gpu_prefilter_rows(searcher, T_masked, toff, Q, tkeys=None, **kw):
    from spacepharer_b200 import api
    searcher.build_index(T_masked, toff)
    qres, qoff = O.concat(Q)
    off, hits, st = searcher.prefilter(qres, qoff, **kw)
    return [api.format_prefilter_rows(hits[int(off[i]):int(off[i + 1])], tkeys) for i in range(len(Q))], st


@pytest.mark.parametrize("fname,mask,maxseqs", [
    ("pref_mask1.txt.gz", 1, 300), ("pref_mask0.txt.gz", 0, 300),
    ("pref_mask1_max2.txt.gz", 1, 2), ("pref_mask0_max5.txt.gz", 0, 5)])
def test_prefilter_small_golden(searcher, fname, mask, maxseqs):
    T, Q = Cm.load_fasta_set("small")
    tres, toff = _product_mask(searcher, T) if mask else O.concat(T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q, max_seqs=maxseqs)
    ref = Cm.dump("small", fname)
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("fname,maxseqs", [("pref_max300.txt.gz", 300), ("pref_max10.txt.gz", 10),
                                           ("pref_max3.txt.gz", 3)])
def test_prefilter_family_truncation_golden(searcher, fname, maxseqs):
    T, Q = Cm.load_fasta_set("family")
    tres, toff = _product_mask(searcher, T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q, max_seqs=maxseqs)
    ref = Cm.dump("family", fname)
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("sens,thr", [("1.0", 154), ("4.0", 127), ("7.5", 96)])
def test_prefilter_small_sensitivities_golden(searcher, sens, thr):
    """-s 1.0 / 4.0 / 7.5: index-side self-score pruning (thr 127, 154: IndexTable.h:141-148) and doubled similar-k-mer
    lists (thr 96), against the reference binary's lists and the oracle"""
    from spacepharer_b200 import api
    assert api.kmer_threshold(float(sens)) == thr
    T, Q = Cm.load_fasta_set("small")
    tres, toff = _product_mask(searcher, T)
    searcher.build_index(tres, toff, kmer_thr=thr)
    qres, qoff = O.concat(Q)
    off, hits, st = searcher.prefilter(qres, qoff, kmer_thr=thr)
    rows = [api.format_prefilter_rows(hits[int(off[i]):int(off[i + 1])]) for i in range(len(Q))]
    ref = Cm.dump("small", "pref_s%s.txt.gz" % sens)
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:5]
    m80, m40u, _ = Cm.mats()
    pf = O.Prefilter(tres, toff, m80, m40u, kmer_thr=thr)
    ooff, ohits = pf.batch(qres, qoff)
    pf.close()
    assert (off == ooff).all() and (hits["id"] == ohits["id"]).all() and (hits["score"] == ohits["score"]).all()
    assert st["n_kmers"] > 0


@pytest.mark.parametrize("maxseqs", [300, 10, 3])
def test_prefilter_saturated_families_golden(searcher, maxseqs):
    """the all-saturated cut (rescoreHits, QueryMatcher.cpp:149-155,486-505): 60-140-aa queries against 40-member families"""
    T, Q = Cm.load_fasta_set("saturated")
    tres, toff = _product_mask(searcher, T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q, max_seqs=maxseqs)
    ref = Cm.dump("saturated", "pref_max%d.txt.gz" % maxseqs)
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("name", ["small", "long"])
def test_prefilter_exact_kmer_matching_golden(searcher, name):
    """--exact-kmer-matching 1 against the reference binary's lists: one k-mer per window, served by the CTA-per-query kernel"""
    T, Q = Cm.load_fasta_set(name)
    tres, toff = _product_mask(searcher, T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q, exact_kmer=1)
    ref = Cm.dump("criteria", "pref_%s_exact.txt.gz" % name)
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]
    assert st["n_slow_path"] == len(Q)
    if name == "small":
        # -s 1.0: the threshold (154) neither selects k-mers nor prunes the index of an exact-matching search
        # (localKmerThr = 0, Prefiltering.cpp:509-512) -- same lists
        from spacepharer_b200 import api
        assert Cm.dump("criteria", "pref_small_exact_s1.0.txt.gz") == ref
        searcher.build_index(tres, toff, kmer_thr=0)
        qres, qoff = O.concat(Q)
        off, hits, _ = searcher.prefilter(qres, qoff, kmer_thr=154, exact_kmer=1)
        got = [api.format_prefilter_rows(hits[int(off[i]):int(off[i + 1])], None) for i in range(len(Q))]
        assert all(got[i] == ref.get(i, []) for i in range(len(Q)))
    # the default mode on the same context afterwards is untouched
    rows, _ = _gpu_prefilter_rows(searcher, tres, toff, Q)
    ref = Cm.dump(name, "pref_mask1.txt.gz")
    assert all(rows[i] == ref.get(i, []) for i in range(len(Q)))


def test_prefilter_long_golden(searcher):
    T, Q = Cm.load_fasta_set("long")
    tres, toff = _product_mask(searcher, T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q)
    ref = Cm.dump("long", "pref_mask1.txt.gz")
    bad = [(i, rows[i][:3], ref.get(i, [])[:3]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("target,fname", [("targets", "pref_0.txt.gz"), ("targets_rev", "pref_0_rev.txt.gz")])
def test_prefilter_config1_golden(searcher, target, fname):
    tkeys, T, qkeys, Q = Cm.load_config1(target)
    tres, toff = _product_mask(searcher, T)
    rows, st = _gpu_prefilter_rows(searcher, tres, toff, Q, tkeys=tkeys)
    ref = Cm.dump("config1", fname)
    bad = [(int(qkeys[i]), rows[i], ref.get(int(qkeys[i]), [])) for i in range(len(Q)) if rows[i] != ref.get(int(qkeys[i]), [])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("bin_size,maxseqs", [(2, 300), (2, 4), (8, 4), (2048, 7)])
def test_prefilter_random_vs_oracle(searcher, bin_size, maxseqs):
    """seeded synthetic set of the benchmark's shape (scaled down), oracle finishes in seconds"""
    from spacepharer_b200 import synth
    tres, toff, _ = synth.make_targets(n_genomes=60, seed=5)
    qres, qoff, _ = synth.make_query_frames(3000, tres, toff, seed=9)
    m80, m40u, _ = Cm.mats()
    searcher.build_index(tres, toff)
    off, hits, st = searcher.prefilter(qres, qoff, max_seqs=maxseqs, bin_size=bin_size)
    pf = O.Prefilter(tres, toff, m80, m40u, max_seqs=maxseqs, bin_size=bin_size)
    ooff, ohits = pf.batch(qres, qoff)
    pf.close()
    assert (off == ooff).all()
    assert (hits["id"] == ohits["id"]).all() and (hits["score"] == ohits["score"]).all() and (hits["diag"] == ohits["diag"]).all()
    if maxseqs == 300:
        # the warp-per-query kernel must carry the bulk; the CTA kernel only takes what it hands over
        assert st["n_slow_path"] <= 0.1 * (len(qoff) - 1)


@pytest.mark.parametrize("env", [
    {"SPB_PF_CHUNKS": "300", "SPB_PF_HITS_PER_CHUNK": "24"},                 # slot / hit arenas overflow: retry rounds with doubled arenas
    {"SPB_EXPAND_HEAVY": "64"},                                               # almost every query is filled by expand_heavy_kernel
    {"SPB_EXPAND_HEAVY": "65535"},                                            # ... none is
    {"SPB_HITS_SOLO": "32", "SPB_PF_TILEQ": "256"},                           # everything handed to the CTA-per-query launch, many tiles
    {"SPB_PF_TILEQ": "300", "SPB_PF_CHUNKS": "700", "SPB_EXPAND_HEAVY": "300"},
])
def test_prefilter_is_independent_of_arena_sizes_and_hand_over_points(searcher, env):
    """arena capacities, tile size and the hand-over thresholds between the kernels are tuning knobs: every setting gives the
    list of the default configuration (which the other tests pin on the oracle and the reference)"""
    import os
    from spacepharer_b200 import synth
    tres, toff, _ = synth.make_targets(n_genomes=60, seed=5)
    qres, qoff, _ = synth.make_query_frames(2500, tres, toff, seed=9)
    searcher.build_index(tres, toff)
    off0, hits0, _ = searcher.prefilter(qres, qoff)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        off, hits, _ = searcher.prefilter(qres, qoff)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert (off == off0).all()
    assert (hits["id"] == hits0["id"]).all() and (hits["score"] == hits0["score"]).all() and (hits["diag"] == hits0["diag"]).all()


def test_prefilter_edge_cases(searcher):
    """empty batch, queries shorter than the k-mer span, one window only, X-only query"""
    T, _ = Cm.load_fasta_set("small")
    tres, toff = O.concat(T)
    searcher.build_index(tres, toff)
    off, hits, _ = searcher.prefilter(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(off) == 1 and len(hits) == 0
    Q = [T[5][:7], T[5][:8], np.full(12, 20, np.uint8), T[5][:9], T[7][10:22]]
    qres, qoff = O.concat(Q)
    off, hits, _ = searcher.prefilter(qres, qoff)
    m80, m40u, _ = Cm.mats()
    pf = O.Prefilter(tres, toff, m80, m40u)
    ooff, ohits = pf.batch(qres, qoff)
    pf.close()
    assert (off == ooff).all() and (hits["id"] == ohits["id"]).all() and (hits["score"] == ohits["score"]).all()
    assert off[1] == 0 and off[2] == off[1] and off[3] == off[2]   # no windows / single window / X never hit


def test_prefilter_slow_path_matches(searcher):
    """a hit list larger than the shared-memory capacity takes the global-memory variant: same result"""
    rng = np.random.default_rng(17)
    base = rng.integers(0, 20, 60).astype(np.uint8)
    T = []
    for i in range(3000):
        t = base.copy()
        for _ in range(3):
            t[int(rng.integers(0, 60))] = rng.integers(0, 20)
        T.append(np.concatenate([rng.integers(0, 20, int(rng.integers(0, 30))).astype(np.uint8), t]))
    Q = [base[5:28].copy(), base[20:40].copy(), rng.integers(0, 20, 15).astype(np.uint8)]
    tres, toff = O.concat(T)
    qres, qoff = O.concat(Q)
    searcher.build_index(tres, toff)
    m80, m40u, _ = Cm.mats()
    for maxseqs in (300, 50):
        off, hits, st = searcher.prefilter(qres, qoff, max_seqs=maxseqs)
        pf = O.Prefilter(tres, toff, m80, m40u, max_seqs=maxseqs)
        ooff, ohits = pf.batch(qres, qoff)
        pf.close()
        assert st["n_slow_path"] >= 1
        assert (off == ooff).all()
        assert (hits["id"] == ohits["id"]).all() and (hits["score"] == ohits["score"]).all() and (hits["diag"] == ohits["diag"]).all()


# ---------------------------------------------------------------------------------------------- align
def _gpu_align_rows(searcher, T, Q, pairs, db_res, tkeys=None, **kw):
    """pairs: list (per query) of target indices"""
    from spacepharer_b200 import api
    tres, toff = O.concat(T)
    searcher.set_align_targets(tres, toff, tkeys)
    qres, qoff = O.concat(Q)
    poff = np.zeros(len(Q) + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in pairs], out=poff[1:])
    pt = np.concatenate([np.asarray(p, dtype=np.uint32) for p in pairs]) if poff[-1] else np.zeros(0, np.uint32)
    off, rows, cig, st = searcher.align(qres, qoff, poff, pt, db_res, **kw)
    qlens = [len(q) for q in Q]
    tlens = [len(t) for t in T]
    out = [api.format_align_rows(rows[int(off[i]):int(off[i + 1])], cig, qlens, tlens, tkeys) for i in range(len(Q))]
    return out, rows, st


def test_align_small_golden(searcher):
    T, Q = Cm.load_fasta_set("small")
    pref = Cm.dump("small", "pref_mask1.txt.gz")
    pairs = [[int(r.split("\t")[0]) for r in pref.get(i, [])] for i in range(len(Q))]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T))
    ref = Cm.dump("small", "aln_mask1.txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]


def test_align_small_evalue_cut_golden(searcher):
    T, Q = Cm.load_fasta_set("small")
    pref = Cm.dump("small", "pref_mask1.txt.gz")
    pairs = [[int(r.split("\t")[0]) for r in pref.get(i, [])] for i in range(len(Q))]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T), evalue=1e-2)
    ref = Cm.dump("small", "aln_mask1_e0.01.txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]


def test_align_small_allpairs_golden(searcher):
    T, Q = Cm.load_fasta_set("small")
    pairs = [list(range(0, len(T), 3)) if i < 40 else [] for i in range(len(Q))]
    rows, _, st = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T))
    ref = Cm.dump("small", "aln_allpairs.txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]
    assert st["n_cells"] == sum(len(Q[i]) * len(T[t]) for i in range(40) for t in pairs[i])


def test_align_family_golden(searcher):
    T, Q = Cm.load_fasta_set("family")
    pref = Cm.dump("family", "pref_max300.txt.gz")
    pairs = [[int(r.split("\t")[0]) for r in pref.get(i, [])] for i in range(len(Q))]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T))
    ref = Cm.dump("family", "aln_max300.txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]


@pytest.mark.parametrize("tag", sorted(Cm.criteria_cases()))
def test_align_acceptance_criteria_golden(searcher, tag):
    """-c / --cov-mode 0..5, --min-seq-id, --seq-id-mode, --min-aln-len and the rows without backtrace against the reference
    binary's alignment DBs (tests/golden/criteria; Alignment.cpp:378,396,555-574)"""
    from spacepharer_b200 import api
    T, Q, pref, ref, kw = Cm.criteria_case(tag)
    pairs = [[int(r.split("\t")[0]) for r in pref.get(i, [])] for i in range(len(Q))]
    backtrace = kw.pop("backtrace")
    kw["evalue"] = kw.pop("eval_thr")
    tres, toff = O.concat(T)
    searcher.set_align_targets(tres, toff, None)
    qres, qoff = O.concat(Q)
    poff = np.zeros(len(Q) + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in pairs], out=poff[1:])
    pt = np.concatenate([np.asarray(p, dtype=np.uint32) for p in pairs])
    off, rows, cig, st = searcher.align(qres, qoff, poff, pt, Cm.db_res_count(T), **kw)
    qlens, tlens = [len(q) for q in Q], [len(t) for t in T]
    got = [api.format_align_rows(rows[int(off[i]):int(off[i + 1])], cig, qlens, tlens, None, backtrace) for i in range(len(Q))]
    bad = [(i, got[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if got[i] != ref.get(i, [])]
    assert not bad, bad[:4]
    assert st["n_rows"] == sum(len(v) for v in ref.values()) == int(off[-1])


@pytest.mark.parametrize("tag", sorted(Cm.limit_cases()))
def test_align_max_accept_max_rejected_golden(searcher, tag):
    """--max-accept / --max-rejected (Alignment.cpp:353-404) against the reference binary: the verdicts of all pairs are replayed in
    the order of the caller's list (the work order of the kernels is by target length)"""
    c = Cm.limit_cases()[tag]
    T, Q = Cm.load_fasta_set("family")
    pref = Cm.dump("family", "pref_max300.txt.gz")
    pairs = [[int(r.split("\t")[0]) for r in pref.get(i, [])] for i in range(len(Q))]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T), evalue=c["evalue"], min_seq_id=c["min_seq_id"],
                                 max_accept=c["max_accept"], max_rejected=c["max_rejected"])
    ref = Cm.dump("criteria", tag + ".txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]
    # the same through the fused search: the resident hit lists are walked in their own order
    if tag == "family_acc3_rej2_id0.9":
        from spacepharer_b200 import api
        tres, toff = _product_mask(searcher, T)
        searcher.build_index(tres, toff)
        qres, qoff = O.concat(Q)
        off, r2, cig, _, _ = searcher.search(qres, qoff, Cm.db_res_count(T), evalue=c["evalue"], min_seq_id=c["min_seq_id"],
                                             max_accept=c["max_accept"], max_rejected=c["max_rejected"])
        got = [api.format_align_rows(r2[int(off[i]):int(off[i + 1])], cig, [len(q) for q in Q], [len(t) for t in T]) for i in range(len(Q))]
        assert all(got[i] == ref.get(i, []) for i in range(len(Q)))


def test_align_long_striped_golden(searcher):
    """queries of 17..150 aa: stripe-dependent E update, byte->word overflow, gapped tracebacks"""
    T, Q = Cm.load_fasta_set("long")
    pairs = [list(range(len(T))) for _ in Q]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T), evalue=0.1)
    ref = Cm.dump("long", "aln_allpairs_e0.1.txt.gz")
    bad = [(i, rows[i][:2], ref.get(i, [])[:2]) for i in range(len(Q)) if rows[i] != ref.get(i, [])]
    assert not bad, bad[:4]


@pytest.mark.parametrize("target,pref,aln", [("targets", "pref_0.txt.gz", "prot_result.txt.gz"),
                                             ("targets_rev", "pref_0_rev.txt.gz", "prot_result_rev.txt.gz")])
def test_align_config1_golden(searcher, target, pref, aln):
    tkeys, T, qkeys, Q = Cm.load_config1(target)
    key2t = {int(k): i for i, k in enumerate(tkeys)}
    p = Cm.dump("config1", pref)
    pairs = [[key2t[int(r.split("\t")[0])] for r in p.get(int(qkeys[i]), [])] for i in range(len(Q))]
    rows, _, _ = _gpu_align_rows(searcher, T, Q, pairs, Cm.db_res_count(T), tkeys=tkeys)
    ref = Cm.dump("config1", aln)
    bad = [(int(qkeys[i]), rows[i], ref.get(int(qkeys[i]), [])) for i in range(len(Q)) if rows[i] != ref.get(int(qkeys[i]), [])]
    assert not bad, bad[:4]


@pytest.mark.parametrize("go,ge", [(16, 2), (5, 1), (3, 1)])
def test_align_random_vs_oracle(searcher, go, ge):
    """integer fields bit-exact, E-values within 1e-6 relative; low gap costs exercise the
    'E never opens from an F-derived cell' rule and gapped tracebacks"""
    from spacepharer_b200 import synth
    rng = np.random.default_rng(23)
    tres, toff, _ = synth.make_targets(n_genomes=8, seed=6)
    T = [tres[int(toff[i]):int(toff[i + 1])] for i in range(len(toff) - 1)]
    Q, pairs = [], []
    for i in range(120):
        t = int(rng.integers(0, len(T)))
        L = int(rng.integers(9, 33))
        if len(T[t]) <= L + 6:
            continue
        s = int(rng.integers(0, len(T[t]) - L - 4))
        q = list(T[t][s:s + L + 4])
        for _ in range(int(rng.integers(0, 4))):
            q[int(rng.integers(0, len(q)))] = int(rng.integers(0, 20))
        for _ in range(int(rng.integers(0, 3))):
            c = int(rng.integers(2, len(q) - 2))
            if rng.random() < 0.5:
                del q[c]
            else:
                q.insert(c, int(rng.integers(0, 20)))
        Q.append(np.array(q[:L], dtype=np.uint8))
        pairs.append([t] + [int(x) for x in rng.integers(0, len(T), 6)])
    _, _, m40 = Cm.mats()
    db_res = Cm.db_res_count(T)
    rows, raw, _ = _gpu_align_rows(searcher, T, Q, pairs, db_res, gap_open=go, gap_extend=ge,
                                   gumbel=tuple(O.GUMBEL_VTML40_16_2.values()))
    p = O.align_params(db_res, gap_open=go, gap_extend=ge)
    nrow = 0
    for i, q in enumerate(Q):
        exp = {}
        for t in pairs[i]:
            a = O.align_pair(m40, p, q, T[t])
            if a.ok == 1:
                exp[t] = a
        got = {int(r["target"]): r for r in raw if int(r["query"]) == i}
        assert set(got) == set(exp), (i, sorted(got), sorted(exp))
        for t, a in exp.items():
            r = got[t]
            assert (int(r["score"]), int(r["bits"]), int(r["q_start"]), int(r["q_end"]), int(r["t_start"]), int(r["t_end"]),
                    int(r["n_ident"]), int(r["aln_len"])) == (a.score, a.bits, a.qStart, a.qEnd, a.tStart, a.tEnd, a.nIdent, a.alnLen)
            assert abs(float(r["eval"]) - a.eval) <= EVAL_RTOL * abs(a.eval)
            nrow += 1
    assert nrow > 100


def test_fused_search_equals_two_step(searcher):
    """resident prefilter -> align (hit lists stay in HBM) gives the same rows as the two host-level calls"""
    from spacepharer_b200 import api
    T, Q = Cm.load_fasta_set("family")
    tres, toff = _product_mask(searcher, T)
    searcher.build_index(tres, toff)
    ures, utoff = O.concat(T)
    searcher.set_align_targets(ures, utoff)
    qres, qoff = O.concat(Q)
    db_res = Cm.db_res_count(T)
    off, rows, cig, st, pref = searcher.search(qres, qoff, db_res, fetch_prefilter=True)
    poff, phits, _ = searcher.prefilter(qres, qoff)
    assert (pref[0] == poff).all() and (pref[1]["id"] == phits["id"]).all()
    off2, rows2, cig2, _ = searcher.align(qres, qoff, poff, phits["id"], db_res)
    qlens = [len(q) for q in Q]
    tlens = [len(t) for t in T]
    a = api.format_align_rows(rows, cig, qlens, tlens)
    b = api.format_align_rows(rows2, cig2, qlens, tlens)
    assert (off == off2).all() and a == b
    ref = Cm.dump("family", "aln_max300.txt.gz")
    got = {i: a[int(off[i]):int(off[i + 1])] for i in range(len(Q)) if off[i + 1] > off[i]}
    assert got == ref


def row_fields(rows):
    """rows without the struct's padding bytes and without the CIGAR slack offsets (both may differ between runs and contexts)"""
    return tuple(np.asarray(rows[n]).tobytes() for n in rows.dtype.names if n != "cigar_off")


def row_cigars(rows, cig):
    """the CIGAR buffer has unused slack between the rows' slices: compare what the rows point at"""
    return [cig[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])] for r in rows]


def test_pipelined_fetch_matches_synchronous_fetch():
    """spb_align_fetch_begin / _wait (result copied out on a second stream while the next batch runs): two batches in
    flight, each result byte-identical to the synchronous spb_align_fetch of the same batch"""
    from spacepharer_b200 import api, synth
    tres, toff, _ = synth.make_targets(n_genomes=50, seed=12)
    batches = [synth.make_query_frames(800, tres, toff, seed=s) for s in (21, 22, 23)]
    s = api.Searcher(device=0)
    masked, _ = api.tantan_mask(s.seed, tres, toff)
    s.build_index(masked, toff)
    s.set_align_targets(tres, toff)
    db_res = int(toff[-1])
    want = []
    for qres, qoff, _ in batches:
        s.set_queries(qres, qoff)
        s.prefilter_resident()
        s.pairs_from_prefilter()
        s.align_resident(db_res)
        off, rows, cig, _ = s.align_fetch()
        want.append((off.tobytes(), row_fields(rows), row_cigars(rows, bytes(cig))))
    pend = []
    for qres, qoff, _ in batches:
        s.set_queries(qres, qoff)
        s.prefilter_resident()
        s.pairs_from_prefilter()
        s.align_resident(db_res)
        pend.append(s.align_fetch_begin())       # all three stay outstanding
    for v, w in zip(pend, want):
        v.wait()
        assert (v.offsets.tobytes(), row_fields(v.rows), row_cigars(v.rows, v.cigar.tobytes())) == w
        v.close()
    assert len(want[0][1]) > 0 and want[0] != want[1]
    s.close()


def test_lanes_stream_matches_single_context():
    """spb_ctx_create_lane: three contexts sharing one index search a stream of different batches concurrently (one host thread
    per lane); every batch's rows are byte-identical to the single-context search, and a lane refuses to rebuild what it borrows"""
    from spacepharer_b200 import api, capi, multigpu, synth
    tres, toff, _ = synth.make_targets(n_genomes=80, seed=31)
    batches = [synth.make_query_frames(600 + 50 * i, tres, toff, seed=40 + i)[:2] for i in range(7)]
    eng = multigpu.ShardedSearcher(device=0, lanes=3)
    masked, _ = api.tantan_mask(eng.searcher.seed, tres, toff)
    eng.load_targets(masked, tres, toff)
    db_res = int(toff[-1])
    want = []
    for qres, qoff in batches:
        r = eng.search_host(qres, qoff, db_res)
        want.append((r["offsets"].tobytes(), row_fields(r["rows"]), row_cigars(r["rows"], bytes(r["cigar"]))))
    got = {}

    def consume(i, q_lo, v):
        got[i] = (v.offsets.tobytes(), row_fields(v.rows), row_cigars(v.rows, v.cigar.tobytes()))

    for _ in range(2):      # twice: the second pass runs on warm lanes
        got.clear()
        sts = eng.run_stream(batches, db_res, consume=consume)
        assert sorted(got) == list(range(len(batches)))
        for i in range(len(batches)):
            assert got[i] == want[i], i
    assert len({st["lane"] for st in sts}) > 1, "the stream never used a second lane"
    assert len(want[0][1]) > 0
    # resident mode (no upload / fetch): same counters as the single context
    sts = eng.run_stream(batches[:3], db_res, steps=6, resident=True)
    assert all(st["align"]["n_rows"] > 0 for st in sts)
    ln = eng.lanes()[1]
    with pytest.raises(capi.SpbError, match="lane"):
        ln.build_index(masked, toff)
    with pytest.raises(capi.SpbError, match="lane"):
        ln.set_align_targets(tres, toff)
    eng.close()
