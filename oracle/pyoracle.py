"""ctypes binding of the parity oracle (oracle/libsp_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALPH = 21
AA = "ACDEFGHIKLMNPQRSTVWYX"


class Matrix(C.Structure):
    _fields_ = [
        ("alph", C.c_int),
        ("aa2num", C.c_ubyte * 256),
        ("num2aa", C.c_char * 32),
        ("sub", (C.c_short * ALPH) * ALPH),
        ("pBack", C.c_double * ALPH),
        ("prob", (C.c_double * ALPH) * ALPH),
        ("lambda_", C.c_double),
    ]


class Hit(C.Structure):
    _fields_ = [("id", C.c_uint32), ("score", C.c_int32), ("diag", C.c_int32)]


class PrefilterParams(C.Structure):
    _fields_ = [
        ("kmerThr", C.c_int),
        ("maxSeqs", C.c_int),
        ("minDiagScore", C.c_int),
        ("binSize", C.c_int),
        ("compBias", C.c_int),
        ("pattern", C.c_char * 33),
        ("exactKmer", C.c_int),
    ]


class QueryStats(C.Structure):
    _fields_ = [("kmers", C.c_uint64), ("dbMatches", C.c_uint64), ("events", C.c_uint64)]


class Gumbel(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "lambda_", "K", "a_I", "b_I", "alpha_I", "beta_I", "a_J", "b_J", "alpha_J", "beta_J", "sigma", "tau")]


class AlignParams(C.Structure):
    _fields_ = [
        ("gapOpen", C.c_int), ("gapExtend", C.c_int),
        ("evalThr", C.c_double),
        ("dbResCount", C.c_uint64),
        ("compBias", C.c_int),
        ("byteLanes", C.c_int), ("wordLanes", C.c_int),
        ("g", Gumbel),
        ("covThr", C.c_float), ("covMode", C.c_int), ("seqIdThr", C.c_float), ("alnLenThr", C.c_int), ("seqIdMode", C.c_int),
        ("swMode", C.c_int),
    ]


class Aln(C.Structure):
    _fields_ = [
        ("ok", C.c_int), ("score", C.c_int), ("bits", C.c_int), ("eval", C.c_double),
        ("qStart", C.c_int), ("qEnd", C.c_int), ("tStart", C.c_int), ("tEnd", C.c_int),
        ("nIdent", C.c_int), ("alnLen", C.c_int), ("seqId", C.c_float),
        ("cigar", C.c_char * 1024),
    ]


# ALP parameters of VTML40, gap open 16 / extend 2 (probed from the reference)
GUMBEL_VTML40_16_2 = dict(
    lambda_=0.34420393530174209, K=0.27487374979561635,
    a_I=0.22018895496520485, b_I=0.026237512825550646,
    alpha_I=0.18481499377626534, beta_I=1.7126891360391483,
    a_J=0.22018895496520485, b_J=0.026237512825550646,
    alpha_J=0.18481499377626534, beta_J=1.7126891360391483,
    sigma=0.18283871043913499, tau=1.7838353361758408)


# ALP parameters of nucleotide.out, gap open 10 / extend 2 (proteinaln2nucl; probed from the reference build, seed 42)
GUMBEL_NUCL_10_2 = dict(
    lambda_=0.63361984397325744, K=0.42150719938226583,
    a_I=0.68666204026925159, b_I=0.18923575026627315,
    alpha_I=0.71225621255786731, beta_I=2.9058945364387716,
    a_J=0.68666204026925159, b_J=0.18923575026627315,
    alpha_J=0.71225621255786731, beta_J=2.9058945364387716,
    sigma=0.7125673897037631, tau=2.8984262849372726)


class NuclParams(C.Structure):
    _fields_ = [("gapOpen", C.c_int), ("gapExtend", C.c_int), ("dbResCount", C.c_uint64), ("g", Gumbel),
                ("letter", C.c_ubyte * 256), ("sub", (C.c_short * 5) * 5)]


class NuclRow(C.Structure):
    _fields_ = [("err", C.c_int), ("rawScore", C.c_int), ("bits", C.c_int), ("eval", C.c_double),
                ("qStart", C.c_int), ("qEnd", C.c_int), ("tStart", C.c_int), ("tEnd", C.c_int), ("qLen", C.c_int), ("tLen", C.c_int),
                ("idCnt", C.c_uint), ("alnLen", C.c_uint), ("seqId", C.c_float), ("cigar", C.c_char * 1024)]


class Orf(C.Structure):
    _fields_ = [("from_", C.c_uint32), ("to", C.c_uint32), ("strand", C.c_int), ("incomplete_start", C.c_int),
                ("incomplete_end", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libsp_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("sp_oracle.c", "sp_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libsp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    u8p = C.POINTER(C.c_ubyte)
    L.spo_matrix_parse.argtypes = [C.c_char_p, C.c_double, C.c_double, C.POINTER(Matrix)]
    L.spo_matrix_parse.restype = C.c_int
    L.spo_comp_bias.argtypes = [C.POINTER(Matrix), u8p, C.c_int, C.POINTER(C.c_float)]
    L.spo_tantan_mask.argtypes = [C.POINTER(Matrix), u8p, C.c_int]
    L.spo_tantan_mask.restype = C.c_int
    L.spo_kmer_threshold.argtypes = [C.c_float]
    L.spo_kmer_threshold.restype = C.c_int
    L.spo_prefilter_create.argtypes = [u8p, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(Matrix),
                                       C.POINTER(Matrix), C.POINTER(PrefilterParams)]
    L.spo_prefilter_create.restype = C.c_void_p
    L.spo_prefilter_destroy.argtypes = [C.c_void_p]
    L.spo_prefilter_num_postings.argtypes = [C.c_void_p]
    L.spo_prefilter_num_postings.restype = C.c_uint64
    L.spo_prefilter_postings.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.spo_prefilter_query.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(Hit), C.POINTER(QueryStats)]
    L.spo_prefilter_query.restype = C.c_int
    L.spo_prefilter_batch.argtypes = [C.c_void_p, u8p, C.POINTER(C.c_uint64), C.c_uint32,
                                      C.POINTER(C.c_uint64), C.POINTER(Hit), C.c_int]
    L.spo_similar_kmers.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_uint32), C.c_int]
    L.spo_similar_kmers.restype = C.c_int
    L.spo_evalue.argtypes = [C.POINTER(Gumbel), C.c_double, C.c_double, C.c_double]
    L.spo_evalue.restype = C.c_double
    L.spo_bitscore.argtypes = [C.POINTER(Gumbel), C.c_double]
    L.spo_bitscore.restype = C.c_double
    L.spo_align_pair.argtypes = [C.POINTER(Matrix), C.POINTER(AlignParams), u8p, C.c_int, u8p, C.c_int,
                                 C.POINTER(Aln)]
    L.spo_sw_forward.argtypes = [C.POINTER(Matrix), C.POINTER(AlignParams), u8p, C.c_int, u8p, C.c_int,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.spo_align_batch.argtypes = [C.POINTER(Matrix), C.POINTER(AlignParams), u8p, C.POINTER(C.c_uint64), C.c_uint32, u8p,
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p, C.c_int]
    L.spo_align_batch.restype = None
    L.spo_extract_frames.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(Orf), C.c_int, C.c_char_p, C.c_int,
                                     C.POINTER(C.c_int)]
    L.spo_extract_frames.restype = C.c_int
    L.spo_format_row.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(Aln), C.c_int, C.c_int]
    L.spo_format_row.restype = C.c_int
    L.spo_can_be_covered.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float]
    L.spo_can_be_covered.restype = C.c_int
    L.spo_format_row_nobt.argtypes = L.spo_format_row.argtypes
    L.spo_format_row_nobt.restype = C.c_int
    L.spo_nucl_default_params.argtypes = [C.POINTER(NuclParams), C.c_uint64]
    L.spo_nucl_default_params.restype = None
    L.spo_nucl_rescore.argtypes = [C.POINTER(NuclParams), C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(NuclRow)]
    L.spo_nucl_rescore.restype = None
    L.spo_format_nucl_row.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(NuclRow)]
    L.spo_format_nucl_row.restype = C.c_int
    L.spo_combine_evalue.argtypes = [C.c_double, C.c_double]
    L.spo_combine_evalue.restype = C.c_double
    _LIB = L
    return L


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def parse_matrix(text, bit_factor, score_bias):
    m = Matrix()
    if isinstance(text, str):
        text = text.encode()
    rc = lib().spo_matrix_parse(text, bit_factor, score_bias, C.byref(m))
    if rc != 0:
        raise ValueError("spo_matrix_parse failed: %d" % rc)
    return m


def matrix_sub(m):
    return np.array([[m.sub[i][j] for j in range(ALPH)] for i in range(ALPH)], dtype=np.int16)


def matrix_pback(m):
    return np.array([m.pBack[i] for i in range(ALPH)], dtype=np.float64)


def encode(seq, m=None):
    """ASCII residues -> numeric (Sequence::mapSequence)."""
    if m is None:
        tab = np.full(256, 20, dtype=np.uint8)
        for i, c in enumerate(AA):
            tab[ord(c)] = i
            tab[ord(c.lower())] = i
        for c, d in (("J", "L"), ("Z", "E"), ("B", "D")):
            tab[ord(c)] = tab[ord(c.lower())] = AA.index(d)
    else:
        tab = np.frombuffer(bytes(m.aa2num), dtype=np.uint8)
    return tab[np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)].copy()


def comp_bias(m, seq):
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    out = np.zeros(len(seq), dtype=np.float32)
    lib().spo_comp_bias(C.byref(m), _u8(seq), len(seq), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def tantan_mask(m80, seq):
    seq = np.array(seq, dtype=np.uint8, copy=True)
    n = lib().spo_tantan_mask(C.byref(m80), _u8(seq), len(seq))
    return seq, n


def concat(seqs):
    """list of uint8 arrays -> (residues, offsets[T+1] uint64)."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    res = np.concatenate(seqs).astype(np.uint8) if len(seqs) else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(res), off


class Prefilter:
    def __init__(self, tres, toff, m80, m40u, kmer_thr=112, max_seqs=300, min_diag=15, bin_size=2,
                 comp_bias=1, pattern="11011101", exact_kmer=0):
        self.tres = np.ascontiguousarray(tres, dtype=np.uint8)
        self.toff = np.ascontiguousarray(toff, dtype=np.uint64)
        self.m80, self.m40u = m80, m40u
        self.p = PrefilterParams(kmer_thr, max_seqs, min_diag, bin_size, comp_bias, pattern.encode(), exact_kmer)
        self.T = len(self.toff) - 1
        self.h = lib().spo_prefilter_create(_u8(self.tres), self.toff.ctypes.data_as(C.POINTER(C.c_uint64)),
                                            self.T, C.byref(m80), C.byref(m40u), C.byref(self.p))
        if not self.h:
            raise ValueError("spo_prefilter_create failed")

    def close(self):
        if self.h:
            lib().spo_prefilter_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def postings(self):
        n = lib().spo_prefilter_num_postings(self.h)
        k = np.zeros(n, dtype=np.uint32)
        i = np.zeros(n, dtype=np.uint32)
        p = np.zeros(n, dtype=np.uint16)
        lib().spo_prefilter_postings(self.h, k.ctypes.data, i.ctypes.data, p.ctypes.data)
        return k, i, p

    def similar(self, kmer6, thr, cap=1 << 20):
        kmer6 = np.ascontiguousarray(kmer6, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint32)
        n = lib().spo_similar_kmers(self.h, _u8(kmer6), thr, out.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
        return out[:n].copy()

    def query(self, q):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        out = (Hit * max(1, self.p.maxSeqs))()
        st = QueryStats()
        n = lib().spo_prefilter_query(self.h, _u8(q), len(q), out, C.byref(st))
        return [(out[i].id, out[i].score, out[i].diag) for i in range(n)], st

    def batch(self, qres, qoff, threads=0):
        qres = np.ascontiguousarray(qres, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        Q = len(qoff) - 1
        off = np.zeros(Q + 1, dtype=np.uint64)
        hits = np.zeros(max(1, Q * self.p.maxSeqs), dtype=[("id", "<u4"), ("score", "<i4"), ("diag", "<i4")])
        lib().spo_prefilter_batch(self.h, _u8(qres), qoff.ctypes.data_as(C.POINTER(C.c_uint64)), Q,
                                  off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  hits.ctypes.data_as(C.POINTER(Hit)), threads)
        return off, hits[: int(off[-1])]


def align_params(db_res_count, gap_open=16, gap_extend=2, eval_thr=200.0, comp_bias=1, gumbel=None,
                 byte_lanes=32, word_lanes=16, cov=0.0, cov_mode=0, min_seq_id=0.0, min_aln_len=0, seq_id_mode=0, sw_mode=0):
    g = Gumbel(**(gumbel or GUMBEL_VTML40_16_2))
    return AlignParams(gap_open, gap_extend, eval_thr, db_res_count, comp_bias, byte_lanes, word_lanes, g,
                       cov, cov_mode, min_seq_id, min_aln_len, seq_id_mode, sw_mode)


def sw_mode_for(alignment_mode, backtrace, cov=0.0, min_seq_id=0.0):
    """Alignment::Alignment + initSWMode (Alignment.cpp:29-37,174-196): -a forces the full mode; --alignment-mode 0 picks by the
    thresholds.  -> sw_mode of align_params (0 full, 1 score + coverage, 2 score only)"""
    if backtrace or alignment_mode == 3:
        return 0
    if alignment_mode == 2:
        return 1
    if alignment_mode == 0:
        if cov > 0.0 and min_seq_id == 0.0:
            return 1
        if cov > 0.0 and min_seq_id > 0.0:
            return 0
    return 2


def align_pair(m40, p, q, t):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    a = Aln()
    lib().spo_align_pair(C.byref(m40), C.byref(p), _u8(q), len(q), _u8(t), len(t), C.byref(a))
    return a


def sw_forward(m40, p, q, t):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    s, qe, te, w = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().spo_sw_forward(C.byref(m40), C.byref(p), _u8(q), len(q), _u8(t), len(t),
                         C.byref(s), C.byref(qe), C.byref(te), C.byref(w))
    return s.value, qe.value, te.value, w.value


def can_be_covered(cov, cov_mode, qlen, tlen):
    return bool(lib().spo_can_be_covered(cov, cov_mode, float(qlen), float(tlen)))


def prefilter_coverage_filter(rows, cov, cov_mode, qlen, tlens):
    """Prefiltering.cpp:834-842: with -c > 0 and --cov-mode 0 / 2 / 5 the hits whose lengths cannot reach the coverage are not
    written (after the max-seqs cut); rows = [(target id, score, diagonal)], tlens = target lengths by id"""
    if not (cov > 0.0 and cov_mode in (0, 2, 5)):
        return list(rows)
    return [r for r in rows if can_be_covered(cov, cov_mode, qlen, tlens[r[0]])]


def decompose_by_amino_acid(entry_lengths, world):
    """DBReader::decomposeDomainByAminoAcid (M/commons/DBReader.cpp:1171-1213) for all ranks: entry_lengths = index lengths
    (residues + 2) in id order -> [(start, count)] per split; greedy cut once a split holds >= ceil(dataSize / world) bytes"""
    n = len(entry_lengths)
    if world == 1:
        return [(0, n)]
    if n <= world:
        return [(r if r < n else 0, 1 if r < n else 0) for r in range(world)]
    data = int(sum(entry_lengths))
    chunk = -(-data // world)
    per = [0] * world
    rank = acc = 0
    for l in entry_lengths:
        if acc >= chunk:
            acc = 0
            rank += 1
        acc += int(l)
        per[rank] += 1
    out, start = [], 0
    for c in per:
        out.append((start, c))
        start += c
    return out


def target_split_prefilter(T, Q, m80, m40u, split, max_seqs=300, threads=4, **kw):
    """MMseqs2's own target split (--split N --split-mode 0): every split is searched with max-seqs / N + 4 sqrt(max-seqs / N)
    (Prefiltering.cpp:343-346), the lists of a query are concatenated and sorted by (|score| desc, key asc) without a cut
    (mergeTargetSplits, Prefiltering.cpp:364-481; hit_t::compareHitsByScoreAndId).  T = masked targets, keys = positions.
    -> per query [(id, score, diag)]"""
    import math
    max_res = min(len(T), max_seqs)
    if split > 1:
        max_res = max(1, max_res // split + int(4 * math.sqrt(max_res / split)))
    qres, qoff = concat(Q)
    rows = [[] for _ in Q]
    for start, count in decompose_by_amino_acid([len(t) + 2 for t in T], split):
        if count == 0:
            continue
        tres, toff = concat(T[start:start + count])
        pf = Prefilter(tres, toff, m80, m40u, max_seqs=max_res, **kw)
        off, hits = pf.batch(qres, qoff, threads=threads)
        pf.close()
        for qi in range(len(Q)):
            rows[qi] += [(int(h["id"]) + start, int(h["score"]), int(h["diag"])) for h in hits[int(off[qi]):int(off[qi + 1])]]
    return [sorted(r, key=lambda h: (-abs(h[1]), h[0])) for r in rows]


def format_row(tkey, a, qlen, tlen, backtrace=True):
    buf = C.create_string_buffer(2048)
    n = (lib().spo_format_row if backtrace else lib().spo_format_row_nobt)(buf, tkey, C.byref(a), qlen, tlen)
    return buf.raw[:n].decode()


def evalue(score, qlen, db_res, gumbel=None):
    g = Gumbel(**(gumbel or GUMBEL_VTML40_16_2))
    return lib().spo_evalue(C.byref(g), float(score), float(qlen), float(db_res))


def bitscore(score, gumbel=None):
    g = Gumbel(**(gumbel or GUMBEL_VTML40_16_2))
    return lib().spo_bitscore(C.byref(g), float(score))


def align_query(m40, p, q, targets, keys, backtrace=True, max_accept=2**31 - 1, max_rejected=2**31 - 1, alt_ali=0):
    """Align q against each (target residues, key) in list order; returns rows sorted like Matcher::compareHits.
    --max-accept / --max-rejected (Alignment.cpp:353-404): the list is walked until max_accept pairs passed or max_rejected
    pairs IN A ROW failed (E-value, coverage, identity, length, or lengths that cannot reach the coverage)."""
    res = []
    passed = rejected = 0
    for t, key in zip(targets, keys):
        if not (passed < max_accept and rejected < max_rejected):
            break
        a = align_pair(m40, p, q, t)
        if a.ok == 1:
            passed += 1
            rejected = 0
        else:
            rejected += 1
        if a.ok == 1:
            res.append((a.eval, -a.bits, len(t), key, format_row(key, a, len(q), len(t), backtrace), a.tStart, a.tEnd, t))
    # --alt-ali N (Alignment::computeAlternativeAlignment, Alignment.cpp:576-611): for every result of the first pass the aligned
    # target stretch [dbStart, dbEnd) is masked with X and the pair aligned again, up to N times while the criteria hold; every
    # further alignment masks its own stretch too
    for r in list(res) if alt_ali > 0 else []:
        key, t = r[3], np.array(r[7], dtype=np.uint8, copy=True)
        t[r[5]:r[6]] = 20
        for _ in range(alt_ali):
            a = align_pair(m40, p, q, t)
            if a.ok != 1:
                break
            res.append((a.eval, -a.bits, len(t), key, format_row(key, a, len(q), len(t), backtrace), a.tStart, a.tEnd, t))
            t[a.tStart:a.tEnd] = 20
    res.sort(key=lambda r: r[:4])       # (stable; the reference's std::sort leaves the order of fully tied rows to its implementation)
    return [r[4] for r in res]


def has_coverage(cov, cov_mode, q_cov, t_cov):
    """Util::hasCoverage (M/commons/Util.cpp:552-567) on float32 coverages"""
    cov, q_cov, t_cov = np.float32(cov), np.float32(q_cov), np.float32(t_cov)
    if cov_mode == 0:
        return bool(q_cov >= cov and t_cov >= cov)
    if cov_mode == 1:
        return bool(t_cov >= cov)
    if cov_mode == 2:
        return bool(q_cov >= cov)
    return True


def align_query_realign(m40, m40_realign, db_res_count, q, targets, keys, eval_thr=200.0, cov=0.0, cov_mode=0, min_seq_id=0.0,
                        min_aln_len=0, realign_max_seqs=2**31 - 1, gap_open=16, gap_extend=2):
    """--realign 1 (Alignment.cpp:47-56,415-447): first pass SCORE ONLY with the alignment matrix (coverage threshold off, the length
    pre-check of canBeCovered on), results sorted like Matcher::compareHits; second pass, in that order and for at most
    realign_max_seqs results: the full alignment with the REALIGN matrix (score bias + --realign-score-bias) without E-value limit,
    kept when its coverage reaches -c; the row carries the boundaries, identity and backtrace of the second pass and the score and
    E-value of the first."""
    p1 = align_params(db_res_count, gap_open, gap_extend, eval_thr, min_seq_id=min_seq_id, min_aln_len=min_aln_len, sw_mode=2)
    p2 = align_params(db_res_count, gap_open, gap_extend, 3.4028234663852886e38)
    first = []
    for t, key in zip(targets, keys):
        if not can_be_covered(cov, cov_mode, len(q), len(t)):
            continue
        a = align_pair(m40, p1, q, t)
        if a.ok == 1:
            first.append((a.eval, -a.bits, len(t), key, t, a.bits))
    first.sort(key=lambda r: r[:4])
    res = []
    for ev, _, tlen, key, t, bits in first:
        if len(res) >= realign_max_seqs:
            break
        a = align_pair(m40_realign, p2, q, t)
        q_cov = np.float32(a.qEnd - a.qStart + 1) / np.float32(len(q))
        t_cov = np.float32(a.tEnd - a.tStart + 1) / np.float32(tlen)
        if has_coverage(cov, cov_mode, q_cov, t_cov):
            a.bits, a.eval = bits, ev
            res.append((ev, -bits, tlen, key, format_row(key, a, len(q), tlen)))
    res.sort(key=lambda r: r[:4])
    return [r[4] for r in res]


ALN_ROW = np.dtype([("ok", "<i4"), ("score", "<i4"), ("bits", "<i4"), ("_pad", "<i4"), ("eval", "<f8"),
                    ("q_start", "<i4"), ("q_end", "<i4"), ("t_start", "<i4"), ("t_end", "<i4"),
                    ("n_ident", "<i4"), ("aln_len", "<i4")])


def align_batch(m40, p, qres, qoff, tres, toff, pair_off, pair_t, threads=0):
    """OpenMP batch alignment of a CSR pair list -> structured rows (one per pair, ok==1 means accepted)."""
    qres = np.ascontiguousarray(qres, dtype=np.uint8)
    tres = np.ascontiguousarray(tres, dtype=np.uint8)
    qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
    toff = np.ascontiguousarray(toff, dtype=np.uint64)
    pair_off = np.ascontiguousarray(pair_off, dtype=np.uint64)
    pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    rows = np.zeros(max(1, int(pair_off[-1])), dtype=ALN_ROW)
    u64p = C.POINTER(C.c_uint64)
    lib().spo_align_batch(C.byref(m40), C.byref(p), _u8(qres), qoff.ctypes.data_as(u64p), len(qoff) - 1, _u8(tres),
                          toff.ctypes.data_as(u64p), pair_off.ctypes.data_as(u64p),
                          pair_t.ctypes.data_as(C.POINTER(C.c_uint32)), rows.ctypes.data, threads)
    return rows[: int(pair_off[-1])]


def extract_frames(seq, min_codons=9, max_codons=32734):
    """one ASCII nucleotide sequence -> [(from, to, strand, incomplete_start, incomplete_end, aa string)] in the order
    of the reference's extractorfs output (M/util/extractorfs.cpp, Orf::findForward)"""
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = max(16, 2 * len(b))
    orfs = (Orf * cap)()
    aa = C.create_string_buffer(2 * len(b) + 16)
    off = (C.c_int * (cap + 1))()
    n = lib().spo_extract_frames(b, len(b), min_codons, max_codons, orfs, cap, aa, len(aa), off)
    assert n <= cap
    return [(orfs[k].from_, orfs[k].to, orfs[k].strand, orfs[k].incomplete_start, orfs[k].incomplete_end,
             aa.raw[off[k]:off[k + 1]].decode()) for k in range(n)]


def orf_header(key, length, o):
    """Orf::writeOrfHeader (M/commons/Orf.cpp:440-457) with the minus-strand flip of M/util/extractorfs.cpp:93-98"""
    f, t = o[0], o[1]
    if o[2] == 1:
        f, t = (length - 1) - f, (length - 1) - t
    flags = int(o[3]) | (int(o[4]) << 1)
    return "%d\t%d%s%d%s" % (key, f, "+" if f < t else "-", abs(f - t), ("\t%d" % flags) if flags else "")


# ---- nucleotide re-alignment branch (--perform-nucl-aln 1) -------------------------------------------------------
def nucl_params(db_res_count):
    p = NuclParams()
    lib().spo_nucl_default_params(C.byref(p), int(db_res_count))
    return p


def nucl_rescore(p, q_nt, t_nt, q_start, q_end, t_start, t_end, cigar, q_start_codon=False, t_start_codon=False):
    """proteinaln2nucl for one protein alignment row; q_nt / t_nt: ASCII nucleotides (bytes) of the two ORFs."""
    r = NuclRow()
    lib().spo_nucl_rescore(C.byref(p), q_nt, len(q_nt), int(q_start_codon), t_nt, len(t_nt), int(t_start_codon),
                           q_start, q_end, t_start, t_end, cigar.encode() if isinstance(cigar, str) else cigar, C.byref(r))
    return r


def format_nucl_row(tkey, r):
    buf = C.create_string_buffer(2048)
    n = lib().spo_format_nucl_row(buf, tkey, C.byref(r))
    return buf.raw[:n].decode()


def combine_prot_nucl(prot_rows, nucl_rows):
    """combineprotnuclaln for one query entry (src/util/CombineProtNuclAln.cpp:55-95) on text rows: per protein row the LAST
    nucleotide row of the same target key gives the sequence identity and min(geometric mean, nucleotide) E-value, both
    re-read from their printed form and printed with %.3E."""
    out = []
    for pr in prot_rows:
        f = pr.split("\t")
        upd, sid = 3.40282346638528859811704183484516925e+38, 0.0
        for nr in nucl_rows:
            g = nr.split("\t")
            if g[0] != f[0]:
                continue
            sid = float(g[2])
            upd = lib().spo_combine_evalue(float(f[3]), float(g[3]))
        out.append("\t".join(f[:2] + ["%.3E" % sid, "%.3E" % upd] + f[4:]))
    return out
