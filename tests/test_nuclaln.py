"""Nucleotide re-alignment branch of predictmatch (--perform-nucl-aln 1 is the reference's default):
proteinaln2nucl (lib/mmseqs/src/util/proteinaln2nucl.cpp) + combineprotnuclaln (src/util/CombineProtNuclAln.cpp).
Golden outputs come from the reference binary (tools/gen_golden_nucl.py): the reference's example (tests/golden/config1,
nucl_*) and the synthetic CRISPR-array fixture (tests/golden/sets, nucl_*; 2556 rows incl. gapped and negative-score ones).

CPU part: the oracle restatement against the golden rows, the host-side combineprotnuclaln module byte for byte.
GPU part: the device kernel behind the C ABI (module CLI, fused predictmatch, Python mirror) against the same rows."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, read_dump
from oracle import pyoracle as O
from spacepharer_b200 import dbio, synth
from test_host_cli import BIN, run
from test_setscore_host import db_from_dump, rows_of

SETS = os.path.join(GOLDEN, "sets")
C1 = os.path.join(GOLDEN, "config1")
P2N_FLAGS = "--sub-mat nucl:nucleotide.out,aa:VTML40.out --gap-open nucl:10,aa:16 --gap-extend 2 --threads 4 --compressed 0 -v 3".split()


@pytest.fixture(scope="module")
def sets_dbs(tmp_path_factory):
    w = str(tmp_path_factory.mktemp("nsets"))
    digest = synth.write_array_search_dbs(w, nucl=True)
    assert digest == open(os.path.join(SETS, "inputs.sha256")).read().strip(), "synthetic inputs drifted from the pinned fixture"
    return w


@pytest.fixture(scope="module")
def c1_dbs(tmp_path_factory):
    """config1 with its <db>_nucl_orf DBs unpacked next to links of the amino-acid DBs"""
    w = str(tmp_path_factory.mktemp("c1n"))
    for f in os.listdir(C1):
        src = os.path.join(C1, f)
        if f.endswith("_nucl_orf.gz"):
            with gzip.open(src, "rb") as g, open(os.path.join(w, f[:-3]), "wb") as o:
                shutil.copyfileobj(g, o)
        elif not f.endswith(".txt.gz") and not f.endswith(".tsv"):
            shutil.copyfile(src, os.path.join(w, f))
    return w


def oracle_entry(p, qn, tn, tkeys_pos, prot_rows):
    out = []
    for r in prot_rows:
        f = r.split("\t")
        rr = O.nucl_rescore(p, qn, tn[tkeys_pos[int(f[0])]], int(f[4]), int(f[5]), int(f[7]), int(f[8]), f[10])
        assert rr.err == 0
        out.append(O.format_nucl_row(int(f[0]), rr).rstrip("\n"))
    return out


def check_oracle(w, q, t, prot_dump, nucl_dump, comb_dump):
    qk, qn = dbio.read_sequence_db(os.path.join(w, q + "_nucl_orf"))
    tk, tn = dbio.read_sequence_db(os.path.join(w, t + "_nucl_orf"))
    qpos = {int(k): i for i, k in enumerate(qk)}
    tpos = {int(k): i for i, k in enumerate(tk)}
    p = O.nucl_params(dbio.amino_acid_db_size(os.path.join(w, t + "_nucl_orf")))
    prot, nucl, comb = read_dump(prot_dump), read_dump(nucl_dump), read_dump(comb_dump)
    n = 0
    for k, rows in prot.items():
        got = oracle_entry(p, qn[qpos[k]], tn, tpos, rows)
        assert got == nucl[k], k
        assert O.combine_prot_nucl(rows, got) == comb[k], k
        n += len(rows)
    return n


def test_oracle_matches_reference_synthetic(sets_dbs):
    n = check_oracle(sets_dbs, "qsetdb", "targets", os.path.join(SETS, "result.txt.gz"), os.path.join(SETS, "nucl_result.txt.gz"),
                     os.path.join(SETS, "nucl_combined.txt.gz"))
    n += check_oracle(sets_dbs, "qsetdb", "targets_rev", os.path.join(SETS, "result_rev.txt.gz"),
                      os.path.join(SETS, "nucl_result_rev.txt.gz"), os.path.join(SETS, "nucl_combined_rev.txt.gz"))
    assert n > 2000
    # the fixture covers gapped backtraces and negative nucleotide scores
    rows = [r for v in read_dump(os.path.join(SETS, "nucl_result.txt.gz")).values() for r in v]
    assert any("D" in r.split("\t")[10] or "I" in r.split("\t")[10] for r in rows)
    assert any(int(r.split("\t")[1]) < 0 for r in rows)


def test_oracle_matches_reference_example(c1_dbs):
    n = check_oracle(c1_dbs, "qsetdb", "targets", os.path.join(C1, "prot_result.txt.gz"), os.path.join(C1, "nucl_result.txt.gz"),
                     os.path.join(C1, "nucl_combined.txt.gz"))
    n += check_oracle(c1_dbs, "qsetdb", "targets_rev", os.path.join(C1, "prot_result_rev.txt.gz"),
                      os.path.join(C1, "nucl_result_rev.txt.gz"), os.path.join(C1, "nucl_combined_rev.txt.gz"))
    assert n >= 15


def test_combineprotnuclaln_module(sets_dbs, tmp_path):
    """host module, reference argv (data/predictmatch.sh:41-45): byte-identical entries; then the set stages on them"""
    w, out = sets_dbs, str(tmp_path)
    qkeys = sorted(dbio.read_entries(w + "/qsetdb"))
    for suffix in ("", "_rev"):
        db_from_dump(out + "/prot" + suffix, os.path.join(SETS, "result%s.txt.gz" % suffix), qkeys)
        db_from_dump(out + "/nucl" + suffix, os.path.join(SETS, "nucl_result%s.txt.gz" % suffix), qkeys)
        run("combineprotnuclaln", out + "/prot" + suffix, out + "/nucl" + suffix, out + "/comb" + suffix, "--threads", "3", "-v", "3")
        assert rows_of(out + "/comb" + suffix) == read_dump(os.path.join(SETS, "nucl_combined%s.txt.gz" % suffix))
        assert sorted(dbio.read_entries(out + "/comb" + suffix)) == qkeys
        assert dbio.read_dbtype(out + "/comb" + suffix) == dbio.DBTYPE_ALIGNMENT_RES
    from test_setscore_host import stages
    stages(w + "/qsetdb", w + "/targets", out + "/comb", out, "")
    stages(w + "/qsetdb", w + "/targets_rev", out + "/comb_rev", out, "_rev")
    assert rows_of(out + "/aggregate") == read_dump(os.path.join(SETS, "nucl_aggregate.txt.gz"))
    assert rows_of(out + "/cScore") == read_dump(os.path.join(SETS, "nucl_cScore.txt.gz"))
    assert rows_of(out + "/cScore_rev") == read_dump(os.path.join(SETS, "nucl_cScore_rev.txt.gz"))
    run("filtermatchbyfdr", out + "/cScore", out + "/cScore_rev", out + "/match", "--fdr", "0.05", "--report-fdr", "0", "--threads", "2")
    assert rows_of(out + "/match") == read_dump(os.path.join(SETS, "nucl_match.txt.gz"))


def test_combineprotnuclaln_missing_entry_is_dropped(tmp_path):
    """CombineProtNuclAln.cpp:46-50: a protein entry without a nucleotide entry is skipped with a warning"""
    out = str(tmp_path)
    row = "7\t40\t0.909\t7.352E-09\t0\t10\t11\t524\t534\t614\t11M\n"
    dbio.write_db(out + "/prot", [(1, row.encode()), (2, row.encode())], dbio.DBTYPE_ALIGNMENT_RES)
    dbio.write_db(out + "/nucl", [(2, b"9\t29\t0.787\t3.170E-03\t0\t32\t33\t1572\t1604\t1842\t33M\n")], dbio.DBTYPE_ALIGNMENT_RES)
    r = run("combineprotnuclaln", out + "/prot", out + "/nucl", out + "/comb")
    assert "Missing nucleotide alignment result for key 1" in r.stderr
    # the only nucleotide row belongs to another target: identity 0, E-value FLT_MAX
    assert rows_of(out + "/comb") == {2: ["7\t40\t0.000E+00\t3.403E+38\t0\t10\t11\t524\t534\t614\t11M"]}


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_proteinaln2nucl_module(sets_dbs, c1_dbs, tmp_path):
    """spb200 proteinaln2nucl with the reference's argv on the reference's protein alignment DBs"""
    out = str(tmp_path)
    w = sets_dbs
    qkeys = sorted(dbio.read_entries(w + "/qsetdb"))
    for suffix in ("", "_rev"):
        db_from_dump(out + "/prot" + suffix, os.path.join(SETS, "result%s.txt.gz" % suffix), qkeys)
        run("proteinaln2nucl", w + "/qsetdb_nucl_orf", w + "/targets%s_nucl_orf" % suffix, w + "/qsetdb", w + "/targets" + suffix,
            out + "/prot" + suffix, out + "/nucl" + suffix, *P2N_FLAGS)
        assert rows_of(out + "/nucl" + suffix) == read_dump(os.path.join(SETS, "nucl_result%s.txt.gz" % suffix))
        assert sorted(dbio.read_entries(out + "/nucl" + suffix)) == qkeys
    w = c1_dbs
    for suffix in ("", "_rev"):
        run("proteinaln2nucl", w + "/qsetdb_nucl_orf", w + "/targets%s_nucl_orf" % suffix, w + "/qsetdb", w + "/targets" + suffix,
            w + "/prot_result" if not suffix else _db(out, w, suffix), out + "/c1nucl" + suffix, *P2N_FLAGS)
        assert rows_of(out + "/c1nucl" + suffix) == read_dump(os.path.join(C1, "nucl_result%s.txt.gz" % suffix))


def _db(out, w, suffix):
    qkeys = sorted(dbio.read_entries(w + "/qsetdb"))
    db_from_dump(out + "/c1prot" + suffix, os.path.join(C1, "prot_result%s.txt.gz" % suffix), qkeys)
    return out + "/c1prot" + suffix


@pytest.mark.gpu
def test_predictmatch_default_reference_example(c1_dbs, tmp_path):
    """BASELINE config 1 with the reference's DEFAULT predictmatch (nucleotide re-alignment on): every intermediate and the
    match DB byte-identical to the reference run"""
    w, out = c1_dbs, str(tmp_path)
    run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/match", out, "--threads", "2", "--keep-intermediate", "1")
    for n, g in (("prot_result", "prot_result"), ("nucl_result", "nucl_result"), ("nucl_result_rev", "nucl_result_rev"),
                 ("result", "nucl_combined"), ("result_rev", "nucl_combined_rev"), ("aggregate", "nucl_sets_aggregate"),
                 ("cScore", "nucl_sets_cScore"), ("cScore_rev", "nucl_sets_cScore_rev"), ("match", "nucl_sets_match")):
        assert rows_of(out + "/" + n) == read_dump(os.path.join(C1, g + ".txt.gz")), n


@pytest.mark.gpu
def test_final_tsv_identical_to_reference(c1_dbs, tmp_path):
    """BASELINE config 1 down to the final match TSV: `spb200 predictmatch` (default pipeline) writes `aggregate` and `match`;
    the reference's own six post-processing modules (data/predictmatch.sh:127-168, run by the reference binary as the checker)
    turn them into result.tsv, which must equal the reference's easy-predict output byte for byte."""
    import common as Cm
    from oracle import refbin
    if not refbin.available():
        pytest.skip("reference build (oracle/_ref/spacepharer, make -C oracle ref) not in this snapshot")
    w, out = c1_dbs, str(tmp_path / "pm")
    os.makedirs(out)
    run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/match", out, "--threads", "2", "--keep-intermediate", "1")
    post = str(tmp_path / "post")
    os.makedirs(post)
    tsv = Cm.reference_postprocess(post, out + "/aggregate", out + "/match")
    gold = open(os.path.join(C1, "result.tsv"), "rb").read()
    assert tsv == gold
    assert tsv.count(b"\n") == 8 and b"AY369265.2\t1.224E+01" in tsv      # 1 + 5 hits under two set lines


def test_reference_postprocess_harness_reproduces_the_golden_tsv(tmp_path):
    """CPU: the checker itself -- the reference's post-processing on the reference's own aggregate / match rows (golden dumps)
    gives the committed result.tsv, so a difference in the GPU test above can only come from our DBs"""
    import common as Cm
    from oracle import refbin
    if not refbin.available():
        pytest.skip("reference build (oracle/_ref/spacepharer, make -C oracle ref) not in this snapshot")
    w = str(tmp_path)
    qkeys = sorted(dbio.read_entries(os.path.join(C1, "qsetdb")))
    # aggregate is keyed by query sequence, match by query set
    db_from_dump(w + "/agg", os.path.join(C1, "nucl_sets_aggregate.txt.gz"), qkeys)
    skeys = sorted(dbio.read_entries(os.path.join(C1, "qsetdb_set_size")))
    db_from_dump(w + "/match", os.path.join(C1, "nucl_sets_match.txt.gz"), skeys, dbtype=dbio.DBTYPE_GENERIC_DB)
    post = w + "/post"
    os.makedirs(post)
    assert Cm.reference_postprocess(post, w + "/agg", w + "/match") == open(os.path.join(C1, "result.tsv"), "rb").read()


@pytest.mark.gpu
def test_predictmatch_default_synthetic_arrays(sets_dbs, tmp_path):
    w, out = sets_dbs, str(tmp_path)
    r = run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/match", out, "--threads", "4", "--keep-intermediate", "1",
            "--fdr", "0.05")
    assert "Nucleotide re-alignment" in r.stdout
    for n, g in (("prot_result", "result"), ("prot_result_rev", "result_rev"), ("nucl_result", "nucl_result"),
                 ("nucl_result_rev", "nucl_result_rev"), ("result", "nucl_combined"), ("result_rev", "nucl_combined_rev"),
                 ("aggregate", "nucl_aggregate"), ("cScore", "nucl_cScore"), ("cScore_rev", "nucl_cScore_rev"), ("match", "nucl_match")):
        assert rows_of(out + "/" + n) == read_dump(os.path.join(SETS, g + ".txt.gz")), n
    assert len(read_dump(os.path.join(SETS, "nucl_match.txt.gz"))) > 10
    # several batches: the formatter thread of batch k runs while batch k+1 is on the GPU -- same DBs
    out2 = out + "/b"
    os.makedirs(out2)
    r = run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out2 + "/match", out2, "--threads", "3", "--keep-intermediate", "1",
            "--fdr", "0.05", "--batch", "1500")
    assert r.stdout.count("Nucleotide re-alignment") >= 2 * 4
    for n in ("prot_result", "nucl_result_rev", "result", "result_rev", "cScore", "match"):
        assert dbio.read_entries(out2 + "/" + n) == dbio.read_entries(out + "/" + n), n
    # without the nucleotide DBs the default branch fails loudly instead of silently searching proteins only
    for f in os.listdir(w):
        if "_nucl_orf" in f:
            os.remove(os.path.join(w, f))
    r = subprocess.run([BIN, "predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/m2", out], capture_output=True, text=True)
    assert r.returncode == 1 and "_nucl_orf" in r.stderr


@pytest.mark.gpu
def test_python_mirror_resident_rescore_vs_oracle():
    """Searcher.nucl_rescore on the resident alignment rows of a seeded synthetic search vs the oracle, row by row
    (integers exact, E-values within 1e-6 relative)"""
    from spacepharer_b200 import api
    tres, toff, set_of = synth.make_targets(n_genomes=60, seed=5)
    qres, qoff, _, _, qnt = synth.make_spacer_arrays(120, tres, toff, set_of, seed=9, return_nucl=True)
    tnt = synth.back_translate(tres, toff)
    s = api.Searcher(device=0)
    masked, _ = api.tantan_mask(s.seed, tres, toff)
    s.build_index(masked, toff)
    s.set_align_targets(tres, toff)
    s.set_queries(qres, qoff)
    s.prefilter_resident()
    s.pairs_from_prefilter()
    db_res = int(toff[-1])
    s.align_resident(db_res)
    s.set_nucl_targets(tnt)
    s.set_nucl_queries(qnt)
    nt_db = 3 * db_res
    nrows, ms = s.nucl_rescore(nt_db)
    off, rows, cig, _ = s.align_fetch()
    assert len(nrows) == len(rows) > 200 and ms > 0
    p = O.nucl_params(nt_db)
    for n, r in zip(nrows, rows):
        ops = cig[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])].decode()
        o = O.nucl_rescore(p, qnt[int(r["query"])], tnt[int(r["target"])], int(r["q_start"]), int(r["q_end"]), int(r["t_start"]),
                           int(r["t_end"]), api.compress_cigar(ops))
        assert (o.rawScore, o.bits, o.qStart, o.qEnd, o.tStart, o.tEnd, o.qLen, o.tLen, o.idCnt, o.alnLen) == \
            (int(n["raw_score"]), int(n["bits"]), int(n["q_start"]), int(n["q_end"]), int(n["t_start"]), int(n["t_end"]),
             int(n["q_len"]), int(n["t_len"]), int(n["n_ident"]), int(n["aln_len"]))
        assert abs(o.eval - float(n["eval"])) <= 1e-6 * abs(o.eval)
        assert np.float32(o.seqId) == np.float32(n["seq_id"])
    # the explicit-row entry point gives the same rows, and the text rows match the oracle's formatting
    nrows2, _ = s.nucl_rescore(nt_db, rows=rows, cigar=cig)
    assert nrows2.tobytes() == nrows.tobytes()
    txt = api.format_nucl_rows(nrows[:50], rows[:50], cig)
    for line, n, r in zip(txt, nrows[:50], rows[:50]):
        ops = cig[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])].decode()
        o = O.nucl_rescore(p, qnt[int(r["query"])], tnt[int(r["target"])], int(r["q_start"]), int(r["q_end"]), int(r["t_start"]),
                           int(r["t_end"]), api.compress_cigar(ops))
        assert line == O.format_nucl_row(int(r["target"]), o).rstrip("\n")
    # a start codon '*' inside the alignment is an error, as in the reference
    sc = np.zeros(len(qnt), dtype=np.uint8)
    q0 = int(rows[0]["query"])
    if int(rows[0]["q_start"]) == 0:
        sc[q0] = 1
        s.set_nucl_queries(qnt, start_codon=sc)
        bad, _ = s.nucl_rescore(nt_db)
        assert int(bad[0]["err"]) == 1
        with pytest.raises(Exception):
            api.format_nucl_rows(bad[:1], rows[:1], cig)
    s.close()


@pytest.mark.gpu
def test_nucl_rescore_error_codes_and_lanes():
    """rows that do not fit their nucleotide entries are flagged, never read out of bounds; a lane re-scores on the parent's
    resident nucleotide targets"""
    from spacepharer_b200 import api, capi
    tres, toff, set_of = synth.make_targets(n_genomes=20, seed=6)
    tnt = synth.back_translate(tres, toff)
    s = api.Searcher(device=0)
    masked, _ = api.tantan_mask(s.seed, tres, toff)
    s.build_index(masked, toff)
    s.set_align_targets(tres, toff)
    s.set_nucl_targets(tnt)
    q_aa = tres[int(toff[3]):int(toff[3]) + 12].copy()
    qnt = [synth.back_translate(q_aa, np.array([0, 12], dtype=np.uint64))[0]]
    s.set_nucl_queries(qnt)
    rows = np.zeros(4, dtype=np.dtype(capi.Aln))
    cig = b"M" * 12
    for i in range(4):
        rows[i]["query"], rows[i]["target"] = 0, 3
        rows[i]["q_start"], rows[i]["q_end"], rows[i]["t_start"], rows[i]["t_end"] = 0, 11, 0, 11
        rows[i]["cigar_off"], rows[i]["cigar_len"] = 0, 12
    rows[1]["q_start"] = 5                      # 12 aligned columns from codon 5 leave a 12-codon query
    rows[2]["target"] = len(toff) + 7           # no such target
    rows[3]["cigar_off"] = 100                  # backtrace outside the buffer
    n, _ = s.nucl_rescore(3 * int(toff[-1]), rows=rows, cigar=cig)
    assert [int(x) for x in n["err"]] == [0, 2, 3, 3]
    assert int(n[0]["n_ident"]) == 36 and int(n[0]["raw_score"]) == 72 and int(n[0]["aln_len"]) == 36
    ln = s.lane()
    ln.set_nucl_queries(qnt)
    n2, _ = ln.nucl_rescore(3 * int(toff[-1]), rows=rows[:1], cigar=cig)
    assert n2[0].tobytes() == n[0].tobytes()
    with pytest.raises(capi.SpbError, match="lane"):
        ln.set_nucl_targets(tnt)
    ln.close()
    s.close()


@pytest.mark.gpu
def test_predictmatch_default_at_scale_matches_reference_digests(tmp_path):
    """BASELINE config 3 at a larger scale (2000 genomes = 124 k ORFs + reversed control, 3000 CRISPR arrays = 153 k frames): the
    reference's DEFAULT predictmatch (search, nucleotide re-alignment, set stages, FDR filter) run module by module with the
    reference binary vs the fused `spb200 predictmatch`; every result DB has the reference's SHA-256 (207 k alignment rows,
    198 k combined set scores, 1160 reported matches; tests/golden/config3_digest.json, tools/gen_golden_config3.py)"""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from gen_golden_config2 import db_digest
    want = json.load(open(os.path.join(GOLDEN, "config3_digest.json")))
    w, out = str(tmp_path / "db"), str(tmp_path / "out")
    os.makedirs(w)
    os.makedirs(out)
    assert synth.write_array_search_dbs(w, nucl=True, **want["shape"]) == want["inputs_sha256"]
    run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/match", out, "--threads", "8", "--keep-intermediate", "1",
        "--fdr", "0.05", "--batch", "50000")
    for n in ("prot_result", "nucl_result", "result", "result_rev", "aggregate", "cScore", "cScore_rev", "match"):
        assert db_digest(out + "/" + n) == (want[n]["sha256"], want[n]["rows"]), n
