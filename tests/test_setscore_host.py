"""Set-level stages of predictmatch (BASELINE config 3) on the host: besthitperset ->
mergeresultsbyset -> combinescore for the target and the control run, then filtermatchbyfdr.  The spb200 module commands
take the reference's argv (data/predictmatch.sh:57-116) and must write, per key, exactly the bytes of the reference's DBs
(tests/golden/sets and tests/golden/config1/sets_*, produced by tools/gen_golden_sets.py with the reference binary).
No GPU needed: the input is the reference's own alignment DB."""
import os

import pytest

from conftest import GOLDEN, read_dump
from spacepharer_b200 import dbio, synth
from test_host_cli import run

SETS = os.path.join(GOLDEN, "sets")
C1 = os.path.join(GOLDEN, "config1")


def db_from_dump(path, dump, keys, dbtype=dbio.DBTYPE_ALIGNMENT_RES):
    rows = read_dump(dump)
    dbio.write_db(path, ((k, "".join(l + "\n" for l in rows.get(k, [])).encode()) for k in keys), dbtype)


def rows_of(db):
    return {k: v.decode().split("\n")[:-1] for k, v in dbio.read_entries(db).items() if v}


@pytest.fixture(scope="module")
def sets_dbs(tmp_path_factory):
    w = str(tmp_path_factory.mktemp("sets"))
    digest = synth.write_array_search_dbs(w)
    assert digest == open(os.path.join(SETS, "inputs.sha256")).read().strip(), "synthetic inputs drifted from the pinned fixture"
    return w


def stages(q, t, result, out, suffix, threads="3"):
    run("besthitperset", q, t, result, out + "/aggregate" + suffix, "--simple-best-hit", "1", "--threads", threads, "--compressed", "0", "-v", "3")
    run("mergeresultsbyset", q + "_set_to_member", out + "/aggregate" + suffix, out + "/aggregate_merged" + suffix, "--threads", threads, "-v", "3")
    run("combinescore", q, t, out + "/aggregate_merged" + suffix, out + "/cScore" + suffix, out, "--threads", threads, "-v", "3")


def test_synthetic_arrays_all_stages(sets_dbs, tmp_path):
    w, out = sets_dbs, str(tmp_path)
    qkeys = sorted(dbio.read_entries(w + "/qsetdb"))
    for suffix in ("", "_rev"):
        db_from_dump(out + "/result" + suffix, os.path.join(SETS, "result%s.txt.gz" % suffix), qkeys)
        stages(w + "/qsetdb", w + "/targets" + suffix, out + "/result" + suffix, out, suffix)
        for n in ("aggregate", "aggregate_merged", "cScore"):
            assert rows_of(out + "/" + n + suffix) == read_dump(os.path.join(SETS, n + suffix + ".txt.gz")), n + suffix
        assert dbio.read_dbtype(out + "/cScore" + suffix) == dbio.DBTYPE_ALIGNMENT_RES
    for fdr, rep, name in (("0.05", "0", "match"), ("0.2", "1", "match_fdr0.2_report"), ("1", "0", "match_fdr1")):
        r = run("filtermatchbyfdr", out + "/cScore", out + "/cScore_rev", out + "/" + name, "--fdr", fdr, "--report-fdr", rep,
                "--compressed", "0", "--threads", "2", "-v", "3")
        assert rows_of(out + "/" + name) == read_dump(os.path.join(SETS, name + ".txt.gz")), name
        assert dbio.read_dbtype(out + "/" + name) == dbio.DBTYPE_GENERIC_DB
        if name == "match":
            assert "Combined score threshold is" in r.stdout
    # every query set keeps an entry (possibly empty), like the reference's writer
    assert sorted(dbio.read_entries(out + "/match")) == sorted(dbio.read_entries(w + "/qsetdb_set_to_member"))


def test_reference_example_all_stages(tmp_path):
    out = str(tmp_path)
    qkeys = sorted(dbio.read_entries(C1 + "/qsetdb"))
    for suffix in ("", "_rev"):
        db_from_dump(out + "/result" + suffix, os.path.join(C1, "prot_result%s.txt.gz" % suffix), qkeys)
        stages(C1 + "/qsetdb", C1 + "/targets" + suffix, out + "/result" + suffix, out, suffix, threads="1")
        assert rows_of(out + "/aggregate" + suffix) == read_dump(os.path.join(C1, "sets_aggregate%s.txt.gz" % suffix))
        assert rows_of(out + "/cScore" + suffix) == read_dump(os.path.join(C1, "sets_cScore%s.txt.gz" % suffix))
    r = run("filtermatchbyfdr", out + "/cScore", out + "/cScore_rev", out + "/match", "--fdr", "0.05", "--report-fdr", "0", "--threads", "1")
    assert rows_of(out + "/match") == read_dump(os.path.join(C1, "sets_match.txt.gz"))
    assert "Combined score list too short" in r.stdout


def test_empty_control_and_edge_cases(tmp_path):
    out = str(tmp_path)
    dbio.write_db(out + "/pos", [(0, b"3\t5.000E+00\t9.000E-01\n7\t1.000E+00\t5.000E-01\n"), (1, b""), (2, b"4\t2.500E+00\t1.000E+00\n")],
                  dbio.DBTYPE_ALIGNMENT_RES)
    dbio.write_db(out + "/neg", [], dbio.DBTYPE_ALIGNMENT_RES)
    r = run("filtermatchbyfdr", out + "/pos", out + "/neg", out + "/m", "--report-fdr", "1")
    assert "control set is empty" in r.stdout
    # threshold = smallest score: everything is printed, FDR 0 (FilterMatchbyFdr.cpp:203-215)
    assert rows_of(out + "/m") == {0: ["3\t5.000E+00\t9.000E-01\t0.000E+00", "7\t1.000E+00\t5.000E-01\t0.000E+00"],
                                  2: ["4\t2.500E+00\t1.000E+00\t0.000E+00"]}
    assert dbio.read_entries(out + "/m")[1] == b""
    # malformed record and unknown member key -> the reference's fatal errors
    dbio.write_db(out + "/bad", [(0, b"3\n")], dbio.DBTYPE_ALIGNMENT_RES)
    r = run("filtermatchbyfdr", out + "/bad", out + "/neg", out + "/m2", ok=False)
    assert r.returncode == 1 and "Invalid result record" in r.stderr
    dbio.write_db(out + "/sets", [(0, b"0\n5\n")], dbio.DBTYPE_ALIGNMENT_RES)
    dbio.write_db(out + "/res", [(0, b"x\n")], dbio.DBTYPE_ALIGNMENT_RES)
    r = run("mergeresultsbyset", out + "/sets", out + "/res", out + "/mg", ok=False)
    assert r.returncode == 1 and "Invalid key 5" in r.stderr


def test_non_simple_best_hit(tmp_path):
    """--simple-best-hit 0: best and second best hit of the set (besthitperset.cpp:62-71,92-94)"""
    import math
    out = str(tmp_path)
    dbio.write_db(out + "/t_member_to_set", [(0, b"0\n"), (1, b"0\n"), (2, b"1\n")], dbio.DBTYPE_ALIGNMENT_RES)
    dbio.write_db(out + "/t_set_size", [(0, b"2\n"), (1, b"1\n")], dbio.DBTYPE_GENERIC_DB)
    dbio.write_db(out + "/res", [(5, b"0\t30\t0.900\t1.000E-03\t0\t9\t10\t0\t9\t50\t10M\n1\t40\t1.00\t1.000E-05\t0\t9\t10\t0\t9\t60\t10M\n"
                                     b"2\t20\t0.800\t2.000E+00\t0\t9\t10\t0\t9\t70\t10M\n")], dbio.DBTYPE_ALIGNMENT_RES)
    run("besthitperset", out + "/q", out + "/t", out + "/res", out + "/agg", "--simple-best-hit", "0")
    got = rows_of(out + "/agg")[5]
    second_minus_best = -math.log(1e-3 / 2) + math.log(1e-5 / 2)
    assert got[0] == "1\t%.3E\t1.00\t1.000E-05\t0\t9\t10\t0\t9\t60\t10M" % second_minus_best
    assert got[1] == "2\t%.3E\t0.800\t2.000E+00\t0\t9\t10\t0\t9\t70\t10M" % math.log(1 - math.exp(-2.0))
