"""The spb200 binary (host C++ over the C ABI) on the reference's own example DBs: the prefilter and alignment DBs it
writes hold, per query key, exactly the bytes of the reference's pref_0 / prot_result (tests/golden/config1)."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, read_dump
from spacepharer_b200 import dbio
from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS

pytestmark = pytest.mark.gpu
C1 = os.path.join(GOLDEN, "config1")


def run(*args):
    r = subprocess.run([BIN] + list(args), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    return r


def entries_as_rows(db):
    return {k: [l for l in v.decode().split("\n") if l] for k, v in dbio.read_entries(db).items()}


def test_prefilter_and_align_dbs_match_reference(tmp_path):
    pref, aln = str(tmp_path / "pref_0"), str(tmp_path / "aln")
    run("prefilter", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), pref, *PREF_FLAGS)
    assert dbio.read_dbtype(pref) == dbio.DBTYPE_PREFILTER_RES
    assert dbio.read_entries(pref) == dbio.read_entries(os.path.join(C1, "pref_0"))
    run("align", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), pref, aln, *ALN_FLAGS)
    assert dbio.read_dbtype(aln) == dbio.DBTYPE_ALIGNMENT_RES
    assert dbio.read_entries(aln) == dbio.read_entries(os.path.join(C1, "prot_result"))


def test_align_reads_the_reference_prefilter_db(tmp_path):
    aln = str(tmp_path / "aln")
    run("align", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), os.path.join(C1, "pref_0"), aln, *ALN_FLAGS)
    assert dbio.read_entries(aln) == dbio.read_entries(os.path.join(C1, "prot_result"))


def test_control_db_and_fused_search(tmp_path):
    """reversed-ORF control setDB (second search of predictmatch) + the fused search command"""
    pref, aln, fused = str(tmp_path / "pref_rev"), str(tmp_path / "aln_rev"), str(tmp_path / "fused")
    run("prefilter", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets_rev"), pref, *PREF_FLAGS)
    got = {k: v for k, v in entries_as_rows(pref).items() if v}
    assert got == read_dump(os.path.join(C1, "pref_0_rev.txt.gz"))
    run("align", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets_rev"), pref, aln, *ALN_FLAGS)
    got = {k: v for k, v in entries_as_rows(aln).items() if v}
    assert got == read_dump(os.path.join(C1, "prot_result_rev.txt.gz"))
    # (flag, value) pairs must stay together: rebuild from both lists without duplicates
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    run("search", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets_rev"), fused, *merged)
    assert {k: v for k, v in entries_as_rows(fused).items() if v} == read_dump(os.path.join(C1, "prot_result_rev.txt.gz"))


def test_predictmatch_reference_example(tmp_path):
    """BASELINE config 1 + control through the set-level stages: the `match` DB of predictmatch (--perform-nucl-aln 0)"""
    out = str(tmp_path)
    r = run("predictmatch", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), os.path.join(C1, "targets_rev"), out + "/match", out,
            "--perform-nucl-aln", "0", "--threads", "2", "--keep-intermediate", "1")
    assert dbio.read_dbtype(out + "/match") == dbio.DBTYPE_GENERIC_DB
    for n, g in (("result", "prot_result"), ("result_rev", "prot_result_rev"), ("aggregate", "sets_aggregate"), ("cScore", "sets_cScore"),
                 ("cScore_rev", "sets_cScore_rev"), ("match", "sets_match")):
        assert {k: v for k, v in entries_as_rows(out + "/" + n).items() if v} == read_dump(os.path.join(C1, g + ".txt.gz")), n
    assert "Combined score list too short" in r.stdout
    # the default (nucleotide re-alignment on, tests/test_nuclaln.py) needs the <db>_nucl_orf DBs: without them (only their
    # gzipped data files live in this directory) it fails loudly instead of silently skipping the branch
    r = subprocess.run([BIN, "predictmatch", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), os.path.join(C1, "targets_rev"),
                        out + "/m2", out], capture_output=True, text=True)
    assert r.returncode == 1 and "_nucl_orf" in r.stderr


def test_predictmatch_synthetic_arrays(tmp_path):
    """BASELINE config 3, scaled down: synthetic CRISPR arrays vs synthetic phage setDB + reversed-ORF control, search on the
    GPU, set-level stages and FDR filter in memory; alignment rows, combined scores and the match DB are byte-identical to
    the reference's (tests/golden/sets)."""
    from spacepharer_b200 import synth
    sets = os.path.join(GOLDEN, "sets")
    w, out = str(tmp_path / "db"), str(tmp_path / "out")
    os.makedirs(w)
    os.makedirs(out)
    assert synth.write_array_search_dbs(w) == open(os.path.join(sets, "inputs.sha256")).read().strip()
    run("predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev", out + "/match", out, "--perform-nucl-aln", "0", "--threads", "4",
        "--keep-intermediate", "1", "--fdr", "0.05")
    for n in ("result", "result_rev", "aggregate", "aggregate_rev", "aggregate_merged", "cScore", "cScore_rev", "match"):
        assert {k: v for k, v in entries_as_rows(out + "/" + n).items() if v} == read_dump(os.path.join(sets, n + ".txt.gz")), n
    assert len(read_dump(os.path.join(sets, "match.txt.gz"))) > 10


def test_search_exhaustive_contains_the_prefiltered_rows(tmp_path):
    """--exhaustive-search 1 (no k-mer stage, BASELINE config 5) on the reference's example DBs: every row the reference's
    prefilter+align search reports is reported, byte-identical and in the same relative order, among the all-vs-all rows"""
    ex = str(tmp_path / "ex")
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    r = run("search", os.path.join(C1, "qsetdb"), os.path.join(C1, "targets"), ex, *merged, "--exhaustive-search", "1")
    assert "Exhaustive search" in r.stdout
    got = entries_as_rows(ex)
    ref = read_dump(os.path.join(C1, "prot_result.txt.gz"))
    more = 0
    for k, rows in ref.items():
        tk = {x.split("\t")[0] for x in rows}
        assert [x for x in got[k] if x.split("\t")[0] in tk] == rows, k
        more += len(got[k]) - len(rows)
    assert more > 0


def test_createindex_cache_is_used_and_validated(tmp_path):
    """<targetDB>.spbidx written by `spb200 createindex` replaces the tantan pre-step of prefilter / search with identical
    results; a cache written for other masking parameters is ignored"""
    import shutil
    for f in os.listdir(C1):
        if f.startswith("targets") and "_" not in f:
            shutil.copyfile(os.path.join(C1, f), str(tmp_path / f))
    db, pref = str(tmp_path / "targets"), str(tmp_path / "pref")
    run("createindex", db, str(tmp_path / "tmp"), "--threads", "2")
    r = run("prefilter", os.path.join(C1, "qsetdb"), db, pref, *PREF_FLAGS)
    assert "masked residues from" in r.stdout
    assert dbio.read_entries(pref) == dbio.read_entries(os.path.join(C1, "pref_0"))
    run("createindex", db, str(tmp_path / "tmp"), "--mask", "0")
    r = run("prefilter", os.path.join(C1, "qsetdb"), db, pref + "2", *PREF_FLAGS)
    assert "ignored" in r.stdout and "masked residues from" not in r.stdout
    assert dbio.read_entries(pref + "2") == dbio.read_entries(os.path.join(C1, "pref_0"))


def _long_set_dbs(tmp_path, fixture="long"):
    """the `long` fixture set as on-disk DBs (keys = order of the FASTA, createdb --shuffle 0) + its all-pairs prefilter DB"""
    from conftest import read_fasta_gz
    w = str(tmp_path)
    os.makedirs(w, exist_ok=True)
    n = {}
    for name in ("t", "q"):
        seqs = [s.encode() for s in read_fasta_gz(os.path.join(GOLDEN, fixture, name + ".fasta.gz"))]
        dbio.write_sequence_db(os.path.join(w, name + "db"), range(len(seqs)), seqs)
        n[name] = len(seqs)
    body = b"".join(b"%d\t0\t0\n" % t for t in range(n["t"]))
    dbio.write_db(os.path.join(w, "allpairs"), ((q, body) for q in range(n["q"])), dbio.DBTYPE_PREFILTER_RES)
    return w


def _with(flags, **over):
    out = list(flags)
    for k, v in over.items():
        out[out.index(k) + 1] = str(v)
    return out


@pytest.mark.parametrize("tag", ["long_c0.5_m0", "long_c0.3_m1", "long_c0.7_m2", "long_c0.5_m4", "long_id0.3_mode1", "long_len40",
                                 "long_c0.6_m2_id0.5_len25", "long_c0.6_m2_id0.5_len25_noback", "long_mode1", "long_mode2",
                                 "long_mode2_c0.5_m0_id0.3", "long_mode0_c0.6_m2"])
def test_align_acceptance_criteria_flags(tmp_path, tag):
    """`spb200 align` with -c / --cov-mode / --min-seq-id / --seq-id-mode / --min-aln-len / -a 0 --alignment-mode 3 writes the
    reference binary's alignment DB (tests/golden/criteria)"""
    import json
    c = json.load(open(os.path.join(GOLDEN, "criteria", "cases.json")))[tag]
    w = _long_set_dbs(tmp_path)
    over = {"-e": c["evalue"], "-c": c["cov"], "--cov-mode": c["cov_mode"], "--min-seq-id": c["min_seq_id"],
            "--min-aln-len": c["min_aln_len"], "--seq-id-mode": c["seq_id_mode"]}
    if not c["backtrace"]:      # ten columns; --alignment-mode 1 score only, 2 score + coverage, 0 picks by the thresholds, 3 full
        over.update({"-a": 0, "--alignment-mode": c["alignment_mode"]})
    run("align", w + "/qdb", w + "/tdb", w + "/allpairs", w + "/aln", *_with(ALN_FLAGS, **over))
    got = {k: v for k, v in entries_as_rows(w + "/aln").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", tag + ".txt.gz"))


@pytest.mark.parametrize("mode", [0, 1, 5])
def test_prefilter_coverage_flag(tmp_path, mode):
    """`spb200 prefilter -c 0.5`: hits that cannot reach the coverage are left out of the written lists in the modes 0 / 2 / 5
    (Prefiltering.cpp:834-842); the fused search with the same flags equals prefilter + align"""
    w = _long_set_dbs(tmp_path)
    flags = _with(PREF_FLAGS, **{"-c": 0.5, "--cov-mode": mode})
    run("prefilter", w + "/qdb", w + "/tdb", w + "/pref", *flags)
    got = {k: v for k, v in entries_as_rows(w + "/pref").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", "pref_long_c0.5_m%d.txt.gz" % mode))
    aflags = _with(ALN_FLAGS, **{"-c": 0.5, "--cov-mode": mode})
    run("align", w + "/qdb", w + "/tdb", w + "/pref", w + "/aln", *aflags)
    seen, merged = set(), []
    for lst in (flags, aflags):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    run("search", w + "/qdb", w + "/tdb", w + "/fused", *merged)
    assert entries_as_rows(w + "/fused") == entries_as_rows(w + "/aln")


def test_compressed_dbs_in_and_out(tmp_path):
    """--compressed 1: the modules read the reference's compressed query / target / prefilter DBs (tests/golden/compressed) and
    write compressed result DBs whose entries are the reference's (DBReader.cpp:512-539, DBWriter.cpp:254-397)"""
    CZ = os.path.join(GOLDEN, "compressed")
    pref, aln, fused = str(tmp_path / "pref"), str(tmp_path / "aln"), str(tmp_path / "fused")
    pflags, aflags = _with(PREF_FLAGS, **{"--compressed": 1}), _with(ALN_FLAGS, **{"--compressed": 1})
    run("prefilter", os.path.join(CZ, "qsetdb"), os.path.join(CZ, "targets"), pref, *pflags)
    assert dbio.read_dbtype(pref) == dbio.DBTYPE_PREFILTER_RES - (1 << 31)
    assert dbio.read_entries(pref) == dbio.read_entries(os.path.join(C1, "pref_0"))
    run("align", os.path.join(CZ, "qsetdb"), os.path.join(CZ, "targets"), os.path.join(CZ, "pref_0"), aln, *aflags)
    assert dbio.read_dbtype(aln) == dbio.DBTYPE_ALIGNMENT_RES - (1 << 31)
    assert dbio.read_entries(aln) == dbio.read_entries(os.path.join(C1, "prot_result"))
    # compressed input, plain output
    seen, merged = set(), []
    for lst in (PREF_FLAGS, ALN_FLAGS):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    run("search", os.path.join(CZ, "qsetdb"), os.path.join(CZ, "targets"), fused, *merged)
    assert dbio.read_dbtype(fused) == dbio.DBTYPE_ALIGNMENT_RES
    assert dbio.read_entries(fused) == dbio.read_entries(os.path.join(C1, "prot_result"))


def test_exact_kmer_matching_flag(tmp_path):
    """`spb200 prefilter --exact-kmer-matching 1` writes the reference binary's lists (tests/golden/criteria/pref_long_exact)"""
    w = _long_set_dbs(tmp_path)
    run("prefilter", w + "/qdb", w + "/tdb", w + "/pref", *_with(PREF_FLAGS, **{"--exact-kmer-matching": 1}))
    got = {k: v for k, v in entries_as_rows(w + "/pref").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", "pref_long_exact.txt.gz"))
    # -s 1.0: the index of an exact-matching search is not pruned by the k-mer threshold (Prefiltering.cpp:509-512)
    w = _long_set_dbs(tmp_path / "small", "small")
    run("prefilter", w + "/qdb", w + "/tdb", w + "/pref", *_with(PREF_FLAGS, **{"--exact-kmer-matching": 1, "-s": "1.0"}))
    got = {k: v for k, v in entries_as_rows(w + "/pref").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", "pref_small_exact_s1.0.txt.gz"))


@pytest.mark.parametrize("mask", [0, 1])
def test_mask_lower_case_flag(tmp_path, mask):
    """`spb200 prefilter --mask-lower-case 1` on a target DB with lower-case stretches writes the reference binary's lists; the
    index cache written with the flag is picked up"""
    from conftest import read_fasta_gz
    w = _long_set_dbs(tmp_path, "small")
    seqs = [x.encode() for x in read_fasta_gz(os.path.join(GOLDEN, "criteria", "t_lower.fasta.gz"))]
    dbio.write_sequence_db(w + "/tlow", range(len(seqs)), seqs)
    flags = _with(PREF_FLAGS, **{"--mask-lower-case": 1, "--mask": mask})
    ref = read_dump(os.path.join(GOLDEN, "criteria", "pref_small_lower_mask%d.txt.gz" % mask))
    run("prefilter", w + "/qdb", w + "/tlow", w + "/pref", *flags)
    assert {k: v for k, v in entries_as_rows(w + "/pref").items() if v} == ref
    run("createindex", w + "/tlow", w + "/tmp", "--mask", str(mask), "--mask-lower-case", "1", "--threads", "2")
    r = run("prefilter", w + "/qdb", w + "/tlow", w + "/pref2", *flags)
    assert "masked residues from" in r.stdout
    assert {k: v for k, v in entries_as_rows(w + "/pref2").items() if v} == ref


@pytest.mark.parametrize("name,split,maxseqs", [("family", 3, 20), ("small", 2, 100)])
def test_reference_target_split_flag(tmp_path, name, split, maxseqs):
    """`spb200 prefilter --split N --split-mode 0` writes what the reference's own (lossy) target split writes: targets cut by
    decomposeDomainByAminoAcid, max-seqs / N + 4 sqrt(max-seqs / N) hits per split, lists concatenated and sorted
    (Prefiltering.cpp:343-346,364-481); --split-mode 1 stays the unsplit list"""
    w = _long_set_dbs(tmp_path, name)
    flags = _with(PREF_FLAGS, **{"--split": split, "--split-mode": 0, "--max-seqs": maxseqs})
    r = run("prefilter", w + "/qdb", w + "/tdb", w + "/pref", *flags)
    assert "Target split mode" in r.stdout
    got = {k: v for k, v in entries_as_rows(w + "/pref").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", "pref_%s_split%d_max%d.txt.gz" % (name, split, maxseqs)))
    if name == "family":
        run("prefilter", w + "/qdb", w + "/tdb", w + "/pref1", *_with(PREF_FLAGS, **{"--split": split, "--split-mode": 1, "--max-seqs": 10}))
        got = {k: v for k, v in entries_as_rows(w + "/pref1").items() if v}
        assert got == read_dump(os.path.join(GOLDEN, "family", "pref_max10.txt.gz"))


@pytest.mark.parametrize("tag", ["family_acc2", "family_rej3_id0.95", "family_acc3_rej2_id0.9"])
def test_max_accept_max_rejected_flags(tmp_path, tag):
    """`spb200 align --max-accept / --max-rejected` on the family set writes the reference binary's alignment DB; the fused search
    (lists straight from the resident prefilter) too"""
    import json
    c = json.load(open(os.path.join(GOLDEN, "criteria", "limits.json")))[tag]
    w = _long_set_dbs(tmp_path, "family")
    pref = read_dump(os.path.join(GOLDEN, "family", "pref_max300.txt.gz"))
    nq = len(dbio.read_entries(w + "/qdb"))
    dbio.write_db(w + "/pref300", ((q, "".join(r + "\n" for r in pref.get(q, [])).encode()) for q in range(nq)), dbio.DBTYPE_PREFILTER_RES)
    over = {"-e": c["evalue"], "--min-seq-id": c["min_seq_id"], "--max-accept": c["max_accept"], "--max-rejected": c["max_rejected"]}
    aflags = _with(ALN_FLAGS, **over)
    run("align", w + "/qdb", w + "/tdb", w + "/pref300", w + "/aln", *aflags)
    ref = read_dump(os.path.join(GOLDEN, "criteria", tag + ".txt.gz"))
    assert {k: v for k, v in entries_as_rows(w + "/aln").items() if v} == ref
    seen, merged = set(), []
    for lst in (PREF_FLAGS, aflags):
        for i in range(0, len(lst), 2):
            if lst[i] not in seen:
                seen.add(lst[i])
                merged += [lst[i], lst[i + 1]]
    run("search", w + "/qdb", w + "/tdb", w + "/fused", *merged)
    assert {k: v for k, v in entries_as_rows(w + "/fused").items() if v} == ref


@pytest.mark.parametrize("tag", ["small_alt2", "long_alt3", "long_alt2_c0.5_m2"])
def test_alternative_alignments_flag(tmp_path, tag):
    """`spb200 align --alt-ali N` writes the reference binary's alignment DB byte for byte -- including the order of rows that tie in
    every sort field, which is whatever std::sort makes of the reference's initial order (first-pass results in list order, then
    the alternatives of result 0, 1, ...)"""
    import json
    c = json.load(open(os.path.join(GOLDEN, "criteria", "alt_ali.json")))[tag]
    w = _long_set_dbs(tmp_path, c["set"])
    if c["set"] == "small":       # the all-pairs subset of tools/gen_golden.py: every third target for the first 40 queries
        nt, nq = len(dbio.read_entries(w + "/tdb")), len(dbio.read_entries(w + "/qdb"))
        body = b"".join(b"%d\t0\t0\n" % t for t in range(0, nt, 3))
        dbio.write_db(w + "/allpairs", ((q, body if q < 40 else b"") for q in range(nq)), dbio.DBTYPE_PREFILTER_RES)
    over = {"-e": c["evalue"], "-c": c["cov"], "--cov-mode": c["cov_mode"], "--alt-ali": c["alt_ali"]}
    r = run("align", w + "/qdb", w + "/tdb", w + "/allpairs", w + "/aln", *_with(ALN_FLAGS, **over), "--batch", "25")
    assert "alternative alignments" in r.stdout
    got = {k: v for k, v in entries_as_rows(w + "/aln").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", tag + ".txt.gz"))
    # the fused search does not take the flag
    r = subprocess.run([BIN, "search", w + "/qdb", w + "/tdb", w + "/x", "--alt-ali", "1", "-a", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "only served by the align module" in r.stderr


@pytest.mark.parametrize("tag", ["small_realign_b0.6", "long_realign_b1.0", "long_realign_b1.0_c0.5_m2_max3", "long_realign_b0.2"])
def test_realign_flag(tmp_path, tag):
    """`spb200 align --realign 1` (two passes over the same spb_align call, matrices swapped in between) writes the reference binary's
    alignment DB"""
    import json
    c = json.load(open(os.path.join(GOLDEN, "criteria", "realign.json")))[tag]
    w = _long_set_dbs(tmp_path, c["set"])
    if c["set"] == "small":
        nt, nq = len(dbio.read_entries(w + "/tdb")), len(dbio.read_entries(w + "/qdb"))
        body = b"".join(b"%d\t0\t0\n" % t for t in range(0, nt, 3))
        dbio.write_db(w + "/allpairs", ((q, body if q < 40 else b"") for q in range(nq)), dbio.DBTYPE_PREFILTER_RES)
    over = {"-e": c["evalue"], "-c": c["cov"], "--cov-mode": c["cov_mode"], "--realign": 1, "--realign-score-bias": c["realign_score_bias"],
            "--realign-max-seqs": c["realign_max_seqs"]}
    run("align", w + "/qdb", w + "/tdb", w + "/allpairs", w + "/aln", *_with(ALN_FLAGS, **over), "--batch", "30")
    got = {k: v for k, v in entries_as_rows(w + "/aln").items() if v}
    assert got == read_dump(os.path.join(GOLDEN, "criteria", tag + ".txt.gz"))
