"""Host side of the drop-in boundary (no GPU needed): the C-ABI library loads and exports every symbol of
include/spb200.h, the spb200 binary parses the workflow-generated argv, builds the reference's matrices, reads and
writes the MMseqs2 DB format, and fails loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
import common as Cm
from conftest import GOLDEN, ROOT
from spacepharer_b200 import capi, dbio

BIN = os.path.join(ROOT, "spacepharer_b200", "lib", "spb200")

# the exact flag strings the reference workflow generates for the two modules
PREF_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out --seed-sub-mat nucl:nucleotide.out,aa:VTML80.out -k 6 "
              "--k-score 2147483647 --alph-size nucl:5,aa:21 --max-seq-len 65535 --max-seqs 300 --split 0 "
              "--split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 --comp-bias-corr 1 --diag-score 1 "
              "--exact-kmer-matching 0 --mask 1 --mask-lower-case 0 --min-ungapped-score 15 --add-self-matches 0 "
              "--spaced-kmer-mode 1 --db-load-mode 0 --pca 1 --pcb 1.5 --spaced-kmer-pattern 11011101 --threads 4 "
              "--compressed 0 -v 3 -s 5.7").split()
ALN_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out -a 1 --alignment-mode 2 --alignment-output-mode 0 "
             "--wrapped-scoring 0 -e 200 --min-seq-id 0 --min-aln-len 0 --seq-id-mode 0 --alt-ali 0 -c 0 --cov-mode 0 "
             "--max-seq-len 65535 --comp-bias-corr 1 --max-rejected 2147483647 --max-accept 2147483647 "
             "--add-self-matches 0 --db-load-mode 0 --pca 1 --pcb 1.5 --score-bias 0 --realign 0 "
             "--realign-score-bias -0.2 --realign-max-seqs 2147483647 --gap-open nucl:10,aa:16 --gap-extend 2 "
             "--zdrop 40 --threads 4 --compressed 0 -v 3").split()


def run(*args, ok=True):
    r = subprocess.run([BIN] + list(args), capture_output=True, text=True)
    if ok:
        assert r.returncode == 0, r.stderr
    return r


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "spb200.h")).read()
    declared = set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", hdr))
    lib = C.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert lib.spb_version  # loads without a GPU; no compute call is made here


@pytest.mark.parametrize("name,bf,bias,key", [("VTML40.out", 2.0, 0.0, "m40"), ("VTML40.out", 2.0, -0.2, "m40u"),
                                              ("VTML80.out", 8.0, -0.2, "m80")])
def test_host_matrices_equal_oracle(name, bf, bias, key):
    m80, m40u, m40 = Cm.mats()
    ref = {"m80": m80, "m40u": m40u, "m40": m40}[key]
    out = run("dumpmatrix", name, str(bf), str(bias)).stdout.split("\n")
    assert out[0] == "ACDEFGHIKLMNPQRSTVWYX"
    sub = np.array([[int(x) for x in out[1 + i].split()] for i in range(21)])
    assert (sub == O.matrix_sub(ref)).all()
    pb = np.array([float(x) for x in out[22].split()])
    assert (pb == np.array(list(ref.pBack))).all()      # bit-identical doubles


@pytest.mark.parametrize("db", ["targets", "qsetdb", "pref_0", "prot_result"])
@pytest.mark.parametrize("threads", [1, 3])
def test_db_round_trip(tmp_path, db, threads):
    """reader -> writer: same entries per key (the reference's multi-threaded writers leave them in any order),
    index sorted by key, dbtype preserved"""
    src = os.path.join(GOLDEN, "config1", db)
    dst = str(tmp_path / "copy")
    run("dbcopy", src, dst, str(threads))
    assert dbio.read_entries(src) == dbio.read_entries(dst)
    assert dbio.read_dbtype(src) == dbio.read_dbtype(dst)
    keys = [int(l.split()[0]) for l in open(dst + ".index")]
    assert keys == sorted(keys)


@pytest.mark.parametrize("threads,every", [(1, 7), (4, 3), (3, 100)])
def test_db_streaming_writer(tmp_path, threads, every):
    """DBWriter::flush between batches (what `spb200 search` does while the next batch is on the device): the entries written
    so far go to the file, the index still carries their final offsets"""
    src = os.path.join(GOLDEN, "config1", "prot_result")
    dst = str(tmp_path / "copy")
    run("dbcopy", src, dst, str(threads), str(every))
    assert dbio.read_entries(src) == dbio.read_entries(dst)
    assert dbio.read_dbtype(src) == dbio.read_dbtype(dst)
    keys = [int(l.split()[0]) for l in open(dst + ".index")]
    assert keys == sorted(keys)


@pytest.mark.parametrize("db", ["targets", "qsetdb", "pref_0", "prot_result"])
def test_compressed_db_is_read(tmp_path, db):
    """--compressed 1 DBs written by the reference (tests/golden/compressed: zstd frames, short entries stored; bit 31 of the
    dbtype) are inflated by the host reader and by the Python mirror into exactly the uncompressed DB"""
    src, plain = os.path.join(GOLDEN, "compressed", db), os.path.join(GOLDEN, "config1", db)
    assert dbio.read_dbtype(src) & 0x7FFFFFFF == dbio.read_dbtype(plain) and dbio.read_dbtype(src) < 0
    assert dbio.read_entries(src) == dbio.read_entries(plain)
    dst = str(tmp_path / "copy")
    out = run("dbcopy", src, dst, "2").stdout
    assert dbio.read_dbtype(dst) == dbio.read_dbtype(plain)
    assert dbio.read_entries(dst) == dbio.read_entries(plain)
    if db in ("targets", "qsetdb"):       # getAminoAcidDBSize counts uncompressed residues (sum of the index lengths)
        assert "%d residues" % dbio.amino_acid_db_size(plain) in out


@pytest.mark.parametrize("db,threads,every", [("targets", 3, 0), ("prot_result", 2, 40), ("qsetdb", 1, 0)])
def test_compressed_db_is_written(tmp_path, db, threads, every):
    """the writer's compressed mode (DBWriter.cpp:254-397): entry = u32 size | bytes | terminator (0 zstd frame, 0xFF stored
    below 60 bytes), index lengths stay uncompressed, dbtype bit 31; read back by both readers, also after streaming flushes"""
    import struct
    plain = os.path.join(GOLDEN, "config1", db)
    dst = str(tmp_path / "c")
    run("dbcopy", plain, dst, str(threads), str(every), "1")
    assert dbio.read_dbtype(dst) == dbio.read_dbtype(plain) - (1 << 31)
    want = dbio.read_entries(plain)
    assert dbio.read_entries(dst) == want
    raw = open(dst, "rb").read()
    n_stored = n_frames = 0
    for line in open(dst + ".index"):
        k, o, l = (int(x) for x in line.split())
        assert l == len(want[k]) + 1
        size = struct.unpack_from("<I", raw, o)[0]
        term = raw[o + 4 + size]
        assert term in (0, 0xFF) and (term == 0xFF) == (len(want[k]) < 60)
        if term == 0:
            assert raw[o + 4:o + 8] == b"\x28\xb5\x2f\xfd"       # zstd frame magic
            n_frames += 1
        else:
            assert size == len(want[k])
            n_stored += 1
    assert n_stored + n_frames == len(want) and (n_frames > 0 or db != "targets")
    run("dbcopy", dst, str(tmp_path / "back"), "2")
    assert dbio.read_entries(str(tmp_path / "back")) == want


def test_large_index_is_parsed_in_parallel(tmp_path):
    """an index text above 1 MB is cut into slices at line ends and parsed by all threads; the entries and their order do not change"""
    n = 90000
    name = str(tmp_path / "big")
    dbio.write_db(name, ((k, b"%d" % (k * 7)) for k in range(n)), dbio.DBTYPE_GENERIC_DB)
    assert os.path.getsize(name + ".index") > (1 << 20)
    run("dbcopy", name, name + "_copy", "5", "1000")
    src, dst = dbio.read_entries(name), dbio.read_entries(name + "_copy")
    assert len(dst) == n and src == dst


def test_vectorised_sequence_db_writer(tmp_path):
    """dbio.write_sequence_db_arrays (the million-entry DBs of bench.py's disk_to_disk leg) writes the bytes of the entry-by-entry
    writer, and the host binary reads them back"""
    import numpy as np
    from spacepharer_b200 import synth
    tres, toff, _ = synth.make_targets(n_genomes=20, seed=11)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    dbio.write_sequence_db_arrays(a, tres, toff)
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    txt = lut[tres]
    n = len(toff) - 1
    dbio.write_sequence_db(b, range(n), (txt[int(toff[i]):int(toff[i + 1])].tobytes() for i in range(n)))
    for ext in ("", ".index", ".dbtype"):
        assert open(a + ext, "rb").read() == open(b + ext, "rb").read()
    out = run("dbcopy", a, str(tmp_path / "c")).stdout
    assert "%d entries, %d residues" % (n, int(toff[-1])) in out


def test_row_formatter_matches_printf():
    """fastfmt.h: `%.3E` and integers of the result rows without stdio, byte for byte what printf prints (random doubles, the
    E-value range, values on and next to the four-digit rounding boundaries, denormals)"""
    out = run("fmtcheck", "400000", "11").stdout
    assert "400000 values, 0 mismatches" in out


def test_workflow_argv_is_accepted():
    out = run("showparams", "prefilter", "q", "t", "p", *PREF_FLAGS).stdout
    assert "kmer_thr=112" in out and "max_seqs=300" in out and "pattern=11011101" in out and "seed_sub_mat=VTML80.out" in out
    out = run("showparams", "align", "q", "t", "p", "a", *ALN_FLAGS).stdout
    assert "evalue=200" in out and "gap_open=16" in out and "gap_extend=2" in out and "sub_mat=VTML40.out" in out


def test_acceptance_criteria_flags_are_parsed():
    """-c / --cov-mode / --min-seq-id / --seq-id-mode / --min-aln-len and --alignment-mode 3 without -a reach the align
    parameters (Parameters.h:394-402 holds -c and --min-seq-id as float)"""
    out = run("showparams", "align", "q", "t", "p", "a", *ALN_FLAGS, "-c", "0.8", "--cov-mode", "2", "--min-seq-id", "0.3",
              "--seq-id-mode", "1", "--min-aln-len", "12", "-a", "0", "--alignment-mode", "3").stdout
    assert "cov=0.8 cov_mode=2 min_seq_id=0.3 min_aln_len=12 seq_id_mode=1 backtrace=0" in out
    # Alignment::initSWMode (Alignment.cpp:174-196): -a forces the full mode (0); without it 1 = score only (2), 2 = score + coverage (1),
    # 0 picks by the thresholds
    for extra, want in ((["-a", "1", "--alignment-mode", "1"], 0), (["-a", "0", "--alignment-mode", "3"], 0),
                        (["-a", "0", "--alignment-mode", "2"], 1), (["-a", "0", "--alignment-mode", "1"], 2),
                        (["-a", "0", "--alignment-mode", "0"], 2), (["-a", "0", "--alignment-mode", "0", "-c", "0.5"], 1),
                        (["-a", "0", "--alignment-mode", "0", "-c", "0.5", "--min-seq-id", "0.3"], 0)):
        assert ("sw_mode=%d" % want) in run("showparams", "align", "q", "t", "p", "a", *ALN_FLAGS, *extra).stdout, extra
    out = run("showparams", "prefilter", "q", "t", "p", *PREF_FLAGS, "-c", "0.5", "--cov-mode", "5", "--exact-kmer-matching", "1").stdout
    assert "cov=0.5 cov_mode=5" in out


def test_split_flags():
    """--split N: query split / auto (modes 1, 2) give the unsplit result and are accepted; mode 0, the reference's own lossy target
    split, is served by the prefilter module"""
    for mode in ("0", "1", "2"):
        run("showparams", "prefilter", "q", "t", "p", *PREF_FLAGS, "--split", "3", "--split-mode", mode)
    r = run("showparams", "prefilter", "q", "t", "p", *PREF_FLAGS, "--split-mode", "3", ok=False)
    assert r.returncode == 1 and "Error in argument" in r.stderr
    # the fused search refuses it before touching a device or a DB
    r = run("search", "q", "t", "out", "--split", "2", "--split-mode", "0", ok=False)
    assert r.returncode == 1 and "only served by the prefilter module" in r.stderr


def test_errors_like_the_reference():
    r = run("prefilter", "q", "t", "p", "--no-such-flag", "1", ok=False)
    assert r.returncode == 1 and "Unrecognized parameter" in r.stderr
    run("showparams", "align", "q", "t", "p", "a", "--realign", "1", "-a", "1", "--realign-score-bias", "-0.5", "--realign-max-seqs", "4")
    r = run("showparams", "align", "q", "t", "p", "a", "--realign", "1", "--alt-ali", "2", "-a", "1", ok=False)
    assert r.returncode == 1 and "not supported" in r.stderr
    r = run("search", "q", "t", "o", "--realign", "1", "-a", "1", ok=False)
    assert r.returncode == 1 and "only served by the align module" in r.stderr
    run("showparams", "align", "q", "t", "p", "a", "--alt-ali", "3", "-a", "1")
    r = run("showparams", "align", "q", "t", "p", "a", "--alt-ali", "3", "--alignment-mode", "1", ok=False)
    assert r.returncode == 1 and "start positions" in r.stderr
    r = run("dbcopy", "/nonexistent/db", "/tmp/x", ok=False)
    assert r.returncode == 1 and "Could not open" in r.stderr


@pytest.mark.parametrize("module,flags", [
    ("prefilter", ["--mask-lower-case", "2"]), ("prefilter", ["--alph-size", "nucl:5,aa:13"]), ("prefilter", ["--alph-size", "13"]),
    ("prefilter", ["--taxon-list", "10239"]), ("align", ["--alignment-mode", "4", "-a", "0"]), ("align", ["--mask-lower-case", "2"]),
    ("align", ["--max-accept", "0"]), ("align", ["--cov-mode", "6"]), ("prefilter", ["--add-self-matches", "1"])])
def test_flags_that_would_change_the_result_are_refused(module, flags):
    """A drop-in must not exit 0 with a DB that differs from the reference's: flags of the two modules that this path does not
    implement keep the reference default and any other value fails loudly."""
    base = PREF_FLAGS if module == "prefilter" else ALN_FLAGS
    dbs = ["q", "t", "p"] if module == "prefilter" else ["q", "t", "p", "a"]
    r = run("showparams", module, *dbs, *base, *flags, ok=False)
    assert r.returncode == 1 and ("not supported" in r.stderr or "only the" in r.stderr or "Error in argument" in r.stderr), r.stderr
    # the workflow's own values of the same flags pass
    run("showparams", module, *dbs, *base)
    run("showparams", module, *dbs, *base, "--compressed", "1")


def test_same_query_and_target_db_is_refused():
    """the reference treats a DB searched against itself specially (identity hit first, scoreIdentical); not on this path"""
    for module, dbs in (("prefilter", ["db", "db", "p"]), ("align", ["db", "db", "p", "a"])):
        r = run("showparams", module, *dbs, *(PREF_FLAGS if module == "prefilter" else ALN_FLAGS), ok=False)
        assert r.returncode == 1 and "not supported" in r.stderr


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run("prefilter", os.path.join(GOLDEN, "config1", "qsetdb"), os.path.join(GOLDEN, "config1", "targets"),
            str(tmp_path / "pref"), *PREF_FLAGS, ok=False)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not os.path.exists(str(tmp_path / "pref.dbtype"))


def test_createindex_writes_the_masked_residue_cache(tmp_path):
    """spb200 createindex is host-only: <targetDB>.spbidx holds the tantan-masked residues the index build
    consumes -- the same masking as the oracle's restatement of tantan (IndexBuilder.cpp:137-149)"""
    import shutil
    import struct
    c1 = os.path.join(GOLDEN, "config1")
    for f in os.listdir(c1):
        if f.startswith("targets") and "_" not in f:
            shutil.copyfile(os.path.join(c1, f), str(tmp_path / f))
    db = str(tmp_path / "targets")
    r = run("createindex", db, str(tmp_path / "tmp"), "--threads", "2")
    assert "866 masked" in r.stdout
    raw = open(db + ".spbidx", "rb").read()
    assert raw[:7] == b"SPBIDX1"
    version, mask, mask_lower, _ = struct.unpack_from("<4I", raw, 8)
    T, N, src, _ = struct.unpack_from("<4Q", raw, 24)
    assert (version, mask, mask_lower) == (1, 1, 0) and src == os.path.getsize(db)
    assert raw[56:66] == b"VTML80.out"
    keys, seqs = dbio.read_sequence_db(db)
    assert T == len(keys) and N == sum(len(s) for s in seqs)
    m80, _, _ = Cm.mats()
    masked = np.frombuffer(raw, dtype=np.uint8, count=N, offset=56 + 64 + 4 * T + 8 * (T + 1))
    # entries are stored in data-offset order (the prefilter's target ids)
    order = np.argsort(dbio.read_db(db)[1], kind="stable")
    want = np.concatenate([O.tantan_mask(m80, O.encode(seqs[i], m80))[0] for i in order])
    assert (masked == want).all() and int((want == 20).sum()) >= 866


@pytest.mark.parametrize("mask", [0, 1])
def test_createindex_masks_lower_case_letters(tmp_path, mask):
    """--mask-lower-case 1 (IndexBuilder.cpp:151-160): the residues the index is built from carry X wherever the DB holds a lower-case
    letter, on top of tantan; the cache of the same DB without the flag is not accepted for it"""
    import struct
    from conftest import read_fasta_gz
    seqs = [x.encode() for x in read_fasta_gz(os.path.join(GOLDEN, "criteria", "t_lower.fasta.gz"))]
    db = str(tmp_path / "tdb")
    dbio.write_sequence_db(db, range(len(seqs)), seqs)
    r = run("createindex", db, str(tmp_path / "tmp"), "--threads", "2", "--mask", str(mask), "--mask-lower-case", "1")
    raw = open(db + ".spbidx", "rb").read()
    version, m, mask_lower, _ = struct.unpack_from("<4I", raw, 8)
    T, N, _, checksum = struct.unpack_from("<4Q", raw, 24)
    assert (m, mask_lower) == (mask, 1) and T == len(seqs)
    masked = np.frombuffer(raw, dtype=np.uint8, count=N, offset=56 + 64 + 4 * T + 8 * (T + 1))
    Tn, low = Cm.lower_case_targets()
    want = np.concatenate(Cm.mask_lower_case(Cm.mask_targets(Tn) if mask else Tn, low))
    assert (masked == want).all()
    assert "%d masked" % (int(sum(l.sum() for l in low)) + (int(sum((a != b).sum() for a, b in zip(Cm.mask_targets(Tn), Tn))) if mask else 0)) in r.stdout
    # without the flag: other header, other checksum input
    run("createindex", db, str(tmp_path / "tmp"), "--threads", "2", "--mask", str(mask))
    raw2 = open(db + ".spbidx", "rb").read()
    assert struct.unpack_from("<4I", raw2, 8)[2] == 0 and struct.unpack_from("<4Q", raw2, 24)[3] != checksum


def test_tantan_four_sequences_per_vector_equals_the_scalar_path():
    """spb_tantan_mask on AVX2 hosts runs four targets in the lanes of a vector (same operation order per lane, no FMA): the masks
    are bit-identical to the one-sequence path -- ragged lengths (0 .. 400, incl. groups where one lane ends far earlier), repeat-rich
    sequences with thousands of masked residues, X residues, and a tail that does not fill a group"""
    from spacepharer_b200 import api, submat, matrices
    seed = submat.SubMat(matrices.matrix_text("VTML80.out"), 8.0, -0.2)
    rng = np.random.default_rng(5)
    seqs = []
    for i in range(4003):
        L = int(rng.integers(0, 400)) if i % 7 else int(rng.integers(0, 60))
        if i % 3 == 0:
            s = np.resize(rng.integers(0, 20, size=int(rng.integers(1, 8))), L).astype(np.uint8)
            if L:
                s[rng.integers(0, L, size=L // 10)] = rng.integers(0, 21, size=L // 10)
        else:
            s = rng.integers(0, 21, size=L).astype(np.uint8)
        seqs.append(s)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    res = np.concatenate(seqs)
    got, n = api.tantan_mask(seed, res, off, threads=3)
    os.environ["SPB_TANTAN_SCALAR"] = "1"
    try:
        want, n2 = api.tantan_mask(seed, res, off, threads=3)
    finally:
        del os.environ["SPB_TANTAN_SCALAR"]
    assert n == n2 > 10000 and (got == want).all()
    # ... and to the oracle's restatement on a sample
    m80, _, _ = Cm.mats()
    for i in range(0, 4003, 97):
        a, b = int(off[i]), int(off[i + 1])
        assert (O.tantan_mask(m80, res[a:b])[0] == got[a:b]).all()
