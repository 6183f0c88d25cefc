"""Query preparation: nucleotide spacers -> six-frame peptide fragments.
CPU: the oracle restatement of extractorfs --orf-start-mode 1 + translatenucs against the reference's own output
(tests/golden/frames, written by tools/gen_golden_frames.py with the reference binary).
GPU: spb_spacer_frames against the same fixtures and against the oracle on a seeded set; fragments made resident give
the same search result as uploading them."""
import gzip
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from conftest import GOLDEN

FR = os.path.join(GOLDEN, "frames")
NT_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
AA = "ACDEFGHIKLMNPQRSTVWYX"


def _golden():
    inp = [l.rstrip("\n").split("\t") for l in gzip.open(os.path.join(FR, "input_db.txt.gz"), "rt")]
    aa = [l.rstrip("\n").split("\t")[1] for l in gzip.open(os.path.join(FR, "aa.txt.gz"), "rt")]
    hdr = [l.rstrip("\n").split("\t", 1)[1] for l in gzip.open(os.path.join(FR, "hdr.txt.gz"), "rt")]
    return [(int(k), s) for k, s in inp], aa, hdr


def test_oracle_frames_match_reference():
    inp, aa, hdr = _golden()
    got_aa, got_h = [], []
    for k, s in inp:
        for o in O.extract_frames(s):
            got_aa.append(o[5])
            got_h.append(O.orf_header(k, len(s), o))
    assert len(aa) == 2685
    assert got_aa == aa
    assert got_h == hdr


def test_oracle_frames_edge_cases():
    assert O.extract_frames("AC") == [] and O.extract_frames("") == []
    assert O.extract_frames("GCT" * 8) == []                      # 8 codons < min 9
    r = O.extract_frames("GCT" * 9)
    assert [(o[0], o[1], o[2], o[5]) for o in r if o[2] == 0] == [(0, 26, 0, "A" * 9)]
    assert O.extract_frames("GCT" * 9, min_codons=10) == []
    f0 = [o for o in O.extract_frames("GCTNNN" + "GCT" * 8) if o[2] == 0 and o[0] == 0]
    assert f0[0][5] == "AX" + "A" * 8                             # NNN is X ...
    assert O.extract_frames("GCN" * 9)[0][5] == "A" * 9           # ... GCN is A whatever N stands for
    # fragments come in the order in which they END: the last complete codons of frames 1, 2, 0 sit at 25, 26, 27
    assert [o[0] for o in O.extract_frames("GCT" * 10) if o[2] == 0] == [1, 2, 0]


def _encode(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    nt = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    code = np.full(256, 4, dtype=np.uint8)
    for c, v in NT_CODE.items():
        code[ord(c)] = v
        code[ord(c.lower())] = v
    return code[nt], off


@pytest.fixture(scope="module")
def searcher():
    from spacepharer_b200 import api
    s = api.Searcher(device=0)
    yield s
    s.close()


@pytest.mark.gpu
def test_gpu_frames_match_reference(searcher):
    from spacepharer_b200 import api
    inp, aa, hdr = _golden()
    nt, off = _encode([s for _, s in inp])
    qres, qoff, orfs, st = searcher.spacer_frames(nt, off)
    got = ["".join(AA[x] for x in qres[int(qoff[i]):int(qoff[i + 1])]) for i in range(len(qoff) - 1)]
    assert got == aa
    assert api.format_orf_headers(orfs, np.diff(off).astype(np.int64), [k for k, _ in inp]) == hdr
    assert st["n_frames"] == len(aa)


@pytest.mark.gpu
def test_gpu_frames_vs_oracle_random(searcher):
    rng = np.random.default_rng(99)
    seqs = []
    for i in range(3000):
        L = int(rng.integers(0, 80)) if i % 50 else int(rng.integers(200, 3000))
        p = [0.3, 0.2, 0.2, 0.3] if i % 3 else [0.25, 0.25, 0.25, 0.25]
        s = "".join("ACGT"[c] for c in rng.choice(4, size=L, p=p))
        if i % 11 == 0 and L:
            s = list(s)
            s[int(rng.integers(0, L))] = "N"
            s = "".join(s)
        seqs.append(s)
    nt, off = _encode(seqs)
    for mn, mx in ((9, 32734), (1, 12), (3, 5)):
        qres, qoff, orfs, _ = searcher.spacer_frames(nt, off, min_codons=mn, max_codons=mx)
        k = 0
        for c, s in enumerate(seqs):
            for o in O.extract_frames(s, mn, mx):
                r = orfs[k]
                assert (int(r["contig"]), int(r["from_"]), int(r["to"]), int(r["strand"]), int(r["incomplete_start"]),
                        int(r["incomplete_end"])) == (c, o[0], o[1], o[2], o[3], o[4]), (c, k)
                assert "".join(AA[x] for x in qres[int(qoff[k]):int(qoff[k + 1])]) == o[5], (c, k)
                k += 1
        assert k == len(orfs)
    qres, qoff, orfs, _ = searcher.spacer_frames(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(orfs) == 0 and len(qoff) == 1


@pytest.mark.gpu
def test_gpu_frames_resident_search(searcher):
    """nucleotides -> frames -> prefilter -> align without leaving HBM == the same search on uploaded frames"""
    from spacepharer_b200 import api, synth
    tres, toff, _ = synth.make_targets(n_genomes=60, seed=3)
    ntm, lens = synth.make_spacers(800, tres, toff, seed=8)
    seqs = ["".join("ACGT"[c] for c in ntm[i, :lens[i]]) for i in range(len(lens))]
    nt, off = _encode(seqs)
    masked, _ = api.tantan_mask(searcher.seed, tres, toff)
    searcher.build_index(masked, toff)
    searcher.set_align_targets(tres, toff)
    db_res = int(toff[-1])
    qres, qoff, orfs, _ = searcher.spacer_frames(nt, off, make_resident=True)
    searcher.prefilter_resident()
    searcher.pairs_from_prefilter()
    searcher.align_resident(db_res)
    a = searcher.align_fetch()
    b = searcher.search(qres, qoff, db_res)
    assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes() and a[2] == b[2]
    assert len(a[1]) > 100
    # the synthetic generator's own six-frame translation yields the same set of fragments (its order is frame-major)
    sres, soff, _ = synth.six_frame_fragments(ntm, lens)
    x = sorted(qres[int(qoff[i]):int(qoff[i + 1])].tobytes() for i in range(len(qoff) - 1))
    y = sorted(sres[int(soff[i]):int(soff[i + 1])].tobytes() for i in range(len(soff) - 1))
    assert x == y


@pytest.mark.gpu
def test_gpu_extractframes_host_command(tmp_path):
    """spb200 extractframes = extractorfs --min-length 9 --orf-start-mode 1 + translatenucs (data/createsetdb.sh:88-113) on the
    reference's own input: amino-acid DB and ORF headers byte-identical to the reference's, nucleotide ORF DB consistent"""
    import subprocess
    from spacepharer_b200 import dbio, synth
    from test_host_cli import BIN
    inp, aa, hdr = _golden()
    db = str(tmp_path / "q_nucl")
    dbio.write_db(db, ((k, (s + "\n").encode()) for k, s in inp), dbio.DBTYPE_NUCLEOTIDES)
    out = str(tmp_path / "q")
    r = subprocess.run([BIN, "extractframes", db, out, "--min-length", "9", "--orf-start-mode", "1", "--create-lookup", "1", "--threads", "3"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got_aa = dbio.read_entries(out)
    assert sorted(got_aa) == list(range(len(aa)))
    assert [got_aa[i].decode().rstrip("\n") for i in range(len(aa))] == aa
    assert dbio.read_dbtype(out) == dbio.DBTYPE_AMINO_ACIDS and dbio.read_dbtype(out + "_nucl_orf") == dbio.DBTYPE_NUCLEOTIDES
    for h in (out + "_nucl_orf_h", out + "_h"):
        got_h = dbio.read_entries(h)
        assert [got_h[i].decode().rstrip("\n") for i in range(len(hdr))] == hdr
    # the nucleotide fragments translate to the peptides (canonical table; ambiguous codons -> X)
    nuc = dbio.read_entries(out + "_nucl_orf")
    import ctypes as C
    L = O.lib()
    L.spo_translate_codon.argtypes = [C.c_char_p]
    L.spo_translate_codon.restype = C.c_char
    for i in range(len(aa)):
        s = nuc[i].decode().rstrip("\n")
        assert len(s) == 3 * len(aa[i]), i
        tr = "".join(L.spo_translate_codon(s[j:j + 3].encode()).decode() for j in range(0, len(s), 3))
        assert tr == aa[i], i


@pytest.mark.gpu
def test_gpu_extractframes_at_scale_matches_reference_digests(tmp_path):
    """200 000 nucleotide spacers (some with ambiguous bases) -> 759 138 fragments: the nucleotide ORF DB, its header DB and the
    amino-acid query DB written by `spb200 extractframes` have the SHA-256 of the reference's extractorfs + translatenucs output
    (tests/golden/frames_digest.json, tools/gen_golden_frames_scale.py)"""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from gen_golden_config2 import db_digest
    from gen_golden_frames_scale import write_spacer_db
    from test_host_cli import BIN
    want = json.load(open(os.path.join(GOLDEN, "frames_digest.json")))
    db, out = str(tmp_path / "sp_nucl"), str(tmp_path / "q")
    write_spacer_db(db, **want["shape"])
    r = subprocess.run([BIN, "extractframes", db, out, "--min-length", "9", "--orf-start-mode", "1", "--threads", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for name, path in (("aa", out), ("nucl_orf_h", out + "_nucl_orf_h"), ("nucl_orf", out + "_nucl_orf")):
        assert db_digest(path) == (want[name]["sha256"], want[name]["rows"]), name
