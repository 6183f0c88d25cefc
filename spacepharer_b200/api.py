"""Host-side mirror of the reference's prefilter / align operators over the C ABI.

`Searcher` plays the role of Prefiltering + Alignment (M/prefiltering/Prefiltering.cpp,
M/alignment/Alignment.cpp) for one target DB on one GPU: same parameter names and meaning as the
module flags (-s, -k, --max-seqs, --min-ungapped-score, --comp-bias-corr, -e, --gap-open,
--gap-extend), same error behaviour (exceptions instead of EXIT), results in the reference's row
layouts.  All arithmetic of the hot path runs in libspb200's CUDA kernels; there is no CPU path.
"""
import ctypes as C
import math
import re

import numpy as np

from . import capi, matrices
from .submat import SubMat

# ALP Gumbel parameter sets (M/alignment/EvalueComputation.h:56-75 table + the VTML40 16/2 set the
# reference estimates at start-up with seed 42).  Order: lambda K a_I b_I alpha_I beta_I
# a_J b_J alpha_J beta_J sigma tau.
GUMBEL = {
    ("VTML40.out", 16, 2): (0.34420393530174209, 0.27487374979561635, 0.22018895496520485, 0.026237512825550646,
                            0.18481499377626534, 1.7126891360391483, 0.22018895496520485, 0.026237512825550646,
                            0.18481499377626534, 1.7126891360391483, 0.18283871043913499, 1.7838353361758408),
    ("blosum62.out", 11, 1): (0.27359865037097330642, 0.044620920658722244834, 1.5938724404943873658,
                              -19.959867650284412122, 30.455610143099914211, -622.28684628915891608,
                              1.5938724404943873658, -19.959867650284412122, 30.455610143099914211,
                              -622.28684628915891608, 29.602444874818868215, -601.81087985041381216),
}


def sw_mode_for(alignment_mode, backtrace, cov=0.0, min_seq_id=0.0):
    """Alignment::Alignment + initSWMode (M/alignment/Alignment.cpp:29-37,174-196): `-a 1` forces the full mode, `--alignment-mode 0`
    picks by the thresholds -> sw_mode of the align calls (0 full, 1 score + coverage, 2 score only)"""
    if backtrace or alignment_mode == 3:
        return 0
    if alignment_mode == 2:
        return 1
    if alignment_mode == 0 and cov > 0.0:
        return 1 if min_seq_id == 0.0 else 0
    return 2


def kmer_threshold(sensitivity, kmer_size=6):
    """Prefiltering::getKmerThreshold (M/prefiltering/Prefiltering.cpp:987-1021), sequence search."""
    if kmer_size != 6:
        raise ValueError("only k = 6 is supported")
    base = float(np.float32(163.2))
    return int(base - float(np.float32(sensitivity)) * 8.917)


def gumbel_for(matrix_name, gap_open, gap_extend):
    key = (matrix_name, int(gap_open), int(gap_extend))
    if key not in GUMBEL:
        raise capi.SpbError("no Gumbel parameters for %s gap open %d extend %d: obtain them from the ALP estimator "
                            "of the reference run and pass gumbel=..." % key)
    return GUMBEL[key]


def _copy_struct_array(ptr, n, ctype):
    """one copy of n C structs at `ptr` into a numpy structured array"""
    dt = np.dtype(ctype)
    if not n:
        return np.zeros(0, dtype=dt)
    buf = (C.c_char * (n * dt.itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dt).copy()


def _view(ptr, n, dtype):
    """zero-copy numpy view of n items at a ctypes pointer (valid while the owning result is alive)"""
    dt = np.dtype(dtype)
    if not n:
        return np.zeros(0, dtype=dt)
    addr = C.cast(ptr, C.c_void_p).value
    return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(addr), dtype=dt)


class AlignResultView:
    """Host result of spb_align_fetch without extra copies: `offsets`, `rows`, `cigar` are views of the library's pinned
    result buffers; they are valid until close() (or garbage collection) hands the buffers back."""

    def __init__(self, lib, r, pending=False):
        self._lib, self._r = lib, r
        self._pending = pending
        if not pending:
            self._bind()

    def wait(self):
        """pipelined fetch (Searcher.align_fetch_begin): block until the copy has landed, then expose the views"""
        if self._pending:
            capi.check(self._lib.spb_align_fetch_wait(self._r))
            self._pending = False
            self._bind()
        return self

    def _bind(self):
        r = self._r
        R = r.contents
        self.n_queries = int(R.n_queries)
        self.offsets = _view(R.offsets, self.n_queries + 1, np.uint64)
        self.rows = _view(R.rows, int(R.n_rows), capi.Aln)
        self.cigar = _view(R.cigar_ops, int(R.n_cigar), np.uint8)
        self.stats = Searcher._aln_stats(R)
        self.nbytes = self.offsets.nbytes + self.rows.nbytes + self.cigar.nbytes

    def close(self):
        if self._r is not None:
            self.offsets = self.rows = self.cigar = None
            self._lib.spb_free_align_result(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def tantan_mask(seed, tres, toff, threads=0):
    """Host pre-step of the index build (--mask 1): returns (masked copy of tres, number of masked residues)."""
    import os
    lib = capi.load()
    tres = np.array(tres, dtype=np.uint8, copy=True)
    toff = np.ascontiguousarray(toff, dtype=np.uint64)
    prob = np.ascontiguousarray(seed.prob, dtype=np.float64)
    pb = np.ascontiguousarray(seed.pback, dtype=np.float64)
    n = C.c_uint64()
    f64 = C.POINTER(C.c_double)
    capi.check(lib.spb_tantan_mask(prob.ctypes.data_as(f64), pb.ctypes.data_as(f64), _u8(tres), _u64(toff),
                                   len(toff) - 1, threads or (os.cpu_count() or 1), C.byref(n)))
    return tres, int(n.value)


class Searcher:
    def __init__(self, device=0, sub_mat="VTML40.out", seed_sub_mat="VTML80.out", score_bias=0.0):
        self.lib = capi.load()
        self.seed = SubMat(matrices.matrix_text(seed_sub_mat), 8.0, -0.2)          # Prefiltering.cpp:64
        self.ungapped = SubMat(matrices.matrix_text(sub_mat), 2.0, -0.2)           # Prefiltering.cpp:65
        self.align_mat = SubMat(matrices.matrix_text(sub_mat), 2.0, score_bias)    # Alignment.cpp:159
        self.sub_mat_name = sub_mat.split("/")[-1]
        h = C.c_void_p()
        capi.check(self.lib.spb_ctx_create(device, C.byref(h)))
        self.h = h
        s80 = np.ascontiguousarray(self.seed.sub, dtype=np.int16)
        s40u = np.ascontiguousarray(self.ungapped.sub, dtype=np.int16)
        s40 = np.ascontiguousarray(self.align_mat.sub, dtype=np.int16)
        p80 = np.ascontiguousarray(self.seed.pback, dtype=np.float64)
        p40 = np.ascontiguousarray(self.align_mat.pback, dtype=np.float64)
        i16 = C.POINTER(C.c_int16)
        f64 = C.POINTER(C.c_double)
        capi.check(self.lib.spb_set_matrices(h, s80.ctypes.data_as(i16), p80.ctypes.data_as(f64),
                                             s40u.ctypes.data_as(i16), s40.ctypes.data_as(i16),
                                             p40.ctypes.data_as(f64)))
        self.T = 0

    def lane(self):
        """A second context on the same GPU that reads this searcher's index / alignment targets / nucleotide targets in place
        (spb_ctx_create_lane) and has its own stream, query batch and results.  Drive every lane from its own host thread:
        batches of different lanes overlap on the device.  Close lanes before the parent."""
        ln = Searcher.__new__(Searcher)
        ln.__dict__.update({k: v for k, v in self.__dict__.items() if k != "h"})
        h = C.c_void_p()
        capi.check(self.lib.spb_ctx_create_lane(self.h, C.byref(h)))
        ln.h = h
        return ln

    def stream_ptr(self):
        """cudaStream_t (as int) all work of this searcher is enqueued on"""
        return int(self.lib.spb_ctx_stream(self.h) or 0)

    def launches(self):
        """number of libspb200 kernels launched on this context so far"""
        return int(self.lib.spb_ctx_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.spb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- index -------------------------------------------------------------------------------
    def build_index(self, tres, toff, kmer_thr=112, pattern="11011101"):
        """tres: masked numeric residues (uint8), toff: uint64[T+1]"""
        tres = np.ascontiguousarray(tres, dtype=np.uint8)
        toff = np.ascontiguousarray(toff, dtype=np.uint64)
        self.T = len(toff) - 1
        capi.check(self.lib.spb_index_build(self.h, _u8(tres), _u64(toff), self.T, pattern.encode(), kmer_thr))
        return self.index_stats()

    def index_stats(self):
        st = capi.IndexStats()
        capi.check(self.lib.spb_index_stats_get(self.h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def export_index(self):
        n = self.index_stats()["n_postings"]
        k = np.zeros(n, dtype=np.uint32)
        i = np.zeros(n, dtype=np.uint32)
        p = np.zeros(n, dtype=np.uint16)
        capi.check(self.lib.spb_index_export(self.h, k.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             i.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             p.ctypes.data_as(C.POINTER(C.c_uint16))))
        return k, i, p

    def similar_kmers(self, kmer6, thr, cap=1 << 20):
        kmer6 = np.ascontiguousarray(kmer6, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32()
        capi.check(self.lib.spb_similar_kmers(self.h, _u8(kmer6), thr, out.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              cap, C.byref(n)))
        return out[: min(n.value, cap)].copy()

    # ---- resident stages ---------------------------------------------------------------------
    def set_queries(self, qres, qoff):
        qres = np.ascontiguousarray(qres, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        self.Q = len(qoff) - 1
        capi.check(self.lib.spb_queries_set(self.h, _u8(qres), _u64(qoff), self.Q))

    def spacer_frames(self, nt, nt_off, min_codons=9, max_codons=32734, make_resident=False):
        """Nucleotide spacers -> six-frame peptide fragments on the device (extractorfs --min-length 9 --orf-start-mode 1
        + translatenucs of createsetdb --extractorf-spacer 1).  nt: numeric nucleotides (0..3 = ACGT, else N), nt_off[N+1].
        -> (qres uint8, qoff uint64[F+1], orfs structured array (contig, from_, to, strand, incomplete_start/_end), stats);
        make_resident: the fragments also become the resident query batch (no host round trip before the search)."""
        nt = np.ascontiguousarray(nt, dtype=np.uint8)
        nt_off = np.ascontiguousarray(nt_off, dtype=np.uint64)
        N = len(nt_off) - 1
        r = C.POINTER(capi.FramesResult)()
        capi.check(self.lib.spb_spacer_frames(self.h, _u8(nt), _u64(nt_off), N, min_codons, max_codons, 1 if make_resident else 0,
                                              C.byref(r)))
        try:
            R = r.contents
            F = int(R.n_frames)
            qoff = _view(R.offsets, F + 1, np.uint64).copy()
            qres = _view(R.residues, int(R.n_residues), np.uint8).copy()
            orfs = _copy_struct_array(R.orfs, F, capi.Orf)
            stats = dict(n_frames=F, n_residues=int(R.n_residues), kernel_ms=float(R.kernel_ms))
        finally:
            self.lib.spb_free_frames_result(r)
        if make_resident:
            self.Q = F
        return qres, qoff, orfs, stats

    @staticmethod
    def _pref_stats(R):
        return dict(n_hits=int(R.n_hits), n_kmers=int(R.n_kmers), n_db_matches=int(R.n_db_matches),
                    n_events=int(R.n_events), n_slow_path=int(R.n_slow_path), kernel_ms=float(R.kernel_ms),
                    total_ms=float(R.total_ms), algo_bytes=int(R.algo_bytes), expand_ms=float(R.expand_ms),
                    lookup_ms=float(R.lookup_ms), bookkeep_ms=float(R.bookkeep_ms))

    @staticmethod
    def _aln_stats(R):
        return dict(n_rows=int(R.n_rows), n_pairs=int(R.n_pairs), n_cells=int(R.n_cells), kernel_ms=float(R.kernel_ms),
                    total_ms=float(R.total_ms))

    def prefilter_resident(self, kmer_thr=112, max_seqs=300, min_ungapped_score=15, bin_size=2, comp_bias=1, id_base=0,
                           emit_trunc_keys=0, exact_kmer=0):
        p = capi.PrefilterParams(kmer_thr, max_seqs, min_ungapped_score, bin_size, comp_bias, id_base, emit_trunc_keys, exact_kmer)
        capi.check(self.lib.spb_prefilter_resident(self.h, C.byref(p)))
        R = capi.PrefilterResult()
        capi.check(self.lib.spb_prefilter_stats(self.h, C.byref(R)))
        return self._pref_stats(R)

    def prefilter_fetch(self):
        r = C.POINTER(capi.PrefilterResult)()
        capi.check(self.lib.spb_prefilter_fetch(self.h, C.byref(r)))
        try:
            R = r.contents
            Q = int(R.n_queries)
            off = _view(R.offsets, Q + 1, np.uint64).copy()
            n = int(R.n_hits)
            hits = _copy_struct_array(R.hits, n, capi.Hit)
            stats = self._pref_stats(R)
        finally:
            self.lib.spb_free_prefilter_result(r)
        return off, hits, stats

    def prefilter_device_result(self):
        """(device pointer of the resident spb_hit[n] array, device pointer of offsets uint64[Q+1], n) -- for the
        device-to-device exchange of the multi-GPU target split"""
        hp, op, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        capi.check(self.lib.spb_prefilter_device_result(self.h, C.byref(hp), C.byref(op), C.byref(n)))
        return int(hp.value or 0), int(op.value or 0), int(n.value)

    def prefilter_device_trunc_keys(self):
        """device pointer of the resident hits' truncation keys (uint64 per hit; prefilter_resident(emit_trunc_keys=1))"""
        kp = C.c_void_p()
        capi.check(self.lib.spb_prefilter_device_trunc_keys(self.h, C.byref(kp)))
        return int(kp.value or 0)

    def merge_shards(self, q_lo, q_hi, recv_ptr, n_recv, counts_ptr, n_shards, max_seqs=300, recv_keys_ptr=0):
        """merge the per-shard hit lists of queries q_lo..q_hi (device pointers) into the resident result; with the
        shards' truncation keys (recv_keys_ptr) lists reaching max_seqs are cut in the reference's order"""
        capi.check(self.lib.spb_prefilter_merge_shards(self.h, q_lo, q_hi, C.c_void_p(recv_ptr),
                                                       C.c_void_p(recv_keys_ptr) if recv_keys_ptr else None, n_recv,
                                                       C.c_void_p(counts_ptr), n_shards, max_seqs))

    # ---- multi-GPU index split: k-mer ranges of the table on different GPUs, gathered through peer memory ----
    def shard_index(self, rank, world):
        """keep k-mer range `rank` of `world` of the built table (spb_index_shard); attach the other ranges before searching"""
        capi.check(self.lib.spb_index_shard(self.h, rank, world))

    def shard_export(self):
        """CUDA IPC handles of this context's shard (128 bytes) for a peer process"""
        buf = C.create_string_buffer(128)
        capi.check(self.lib.spb_index_shard_export(self.h, buf))
        return buf.raw

    def shard_attach_ipc(self, peer, handles):
        capi.check(self.lib.spb_index_shard_attach_ipc(self.h, peer, C.c_char_p(bytes(handles))))

    def shard_attach(self, peer, other):
        """attach the shard held by another Searcher of this process (same or peer-accessible device)"""
        capi.check(self.lib.spb_index_shard_attach_ctx(self.h, peer, other.h))

    def pairs_from_prefilter(self):
        capi.check(self.lib.spb_pairs_from_prefilter(self.h))

    def set_pairs(self, pair_off, pair_targets):
        pair_off = np.ascontiguousarray(pair_off, dtype=np.uint64)
        pair_targets = np.ascontiguousarray(pair_targets, dtype=np.uint32)
        capi.check(self.lib.spb_pairs_set(self.h, _u64(pair_off), pair_targets.ctypes.data_as(C.POINTER(C.c_uint32))))

    def _align_params(self, db_res_count, evalue, gap_open, gap_extend, comp_bias, gumbel, byte_lanes, word_lanes,
                      cov=0.0, cov_mode=0, min_seq_id=0.0, min_aln_len=0, seq_id_mode=0, sw_mode=0, max_accept=0, max_rejected=0):
        """cov / cov_mode / min_seq_id / min_aln_len / seq_id_mode: `-c`, `--cov-mode`, `--min-seq-id`, `--min-aln-len`,
        `--seq-id-mode` of the align module (Alignment.cpp:378,396,555-574); all 0 in the predictmatch search.
        sw_mode: 0 full alignment with backtrace (`-a 1` / `--alignment-mode 3`), 1 score + coverage (`--alignment-mode 2`), 2 score
        only (`--alignment-mode 1`) -- see sw_mode_for.  max_accept / max_rejected: `--max-accept` / `--max-rejected`, the ordered walk
        of each query's pair list (Alignment.cpp:353-404); 0 = no limit"""
        g = capi.Gumbel(*(gumbel or gumbel_for(self.sub_mat_name, gap_open, gap_extend)))
        return capi.AlignParams(gap_open, gap_extend, evalue, db_res_count, comp_bias, byte_lanes, word_lanes, g,
                                cov, cov_mode, min_seq_id, min_aln_len, seq_id_mode, sw_mode, max_accept, max_rejected)

    def align_resident(self, db_res_count, evalue=200.0, gap_open=16, gap_extend=2, comp_bias=1, gumbel=None,
                       byte_lanes=32, word_lanes=16, **criteria):
        p = self._align_params(db_res_count, evalue, gap_open, gap_extend, comp_bias, gumbel, byte_lanes, word_lanes, **criteria)
        capi.check(self.lib.spb_align_resident(self.h, C.byref(p)))
        R = capi.AlignResult()
        capi.check(self.lib.spb_align_stats(self.h, C.byref(R)))
        return self._aln_stats(R)

    def pairs_exhaustive(self, db_res_count, evalue=200.0, gap_open=16, gap_extend=2, comp_bias=1, gumbel=None,
                         byte_lanes=32, word_lanes=16, **criteria):
        """Exhaustive search (BASELINE config 5): score the resident queries against EVERY alignment target with the
        streaming register kernel; the pairs reaching the E-value threshold become the resident pair list.
        -> stats dict (n_scored, n_cells, n_kept, n_unfiltered, n_pairs, kernel_ms, total_ms, gcups)"""
        p = self._align_params(db_res_count, evalue, gap_open, gap_extend, comp_bias, gumbel, byte_lanes, word_lanes, **criteria)
        E = capi.ExhaustiveStats()
        capi.check(self.lib.spb_pairs_exhaustive(self.h, C.byref(p), C.byref(E)))
        st = {n: getattr(E, n) for n, _ in capi.ExhaustiveStats._fields_}
        st["gcups"] = (st["n_cells"] / (st["kernel_ms"] * 1e6)) if st["kernel_ms"] > 0 else 0.0
        return st

    def align_exhaustive(self, qres, qoff, db_res_count, **kw):
        """all queries x all targets -> (offsets, rows, cigar, stats) with the rows `align` writes for an all-pairs
        prefilter DB"""
        self.set_queries(qres, qoff)
        est = self.pairs_exhaustive(db_res_count, **kw)
        ast = self.align_resident(db_res_count, **kw)
        off, rows, cig, _ = self.align_fetch()
        return off, rows, cig, dict(exhaustive=est, align=ast)

    def align_fetch_view(self):
        """like align_fetch but zero-copy: returns an AlignResultView (close it when done)"""
        r = C.POINTER(capi.AlignResult)()
        capi.check(self.lib.spb_align_fetch(self.h, C.byref(r)))
        return AlignResultView(self.lib, r)

    def align_fetch_begin(self):
        """start copying the resident alignment result to pinned host memory on a second stream and return at once; the next
        batch can be uploaded and searched meanwhile.  -> AlignResultView whose views exist after .wait()"""
        r = C.POINTER(capi.AlignResult)()
        capi.check(self.lib.spb_align_fetch_begin(self.h, C.byref(r)))
        return AlignResultView(self.lib, r, pending=True)

    def align_fetch(self):
        r = C.POINTER(capi.AlignResult)()
        capi.check(self.lib.spb_align_fetch(self.h, C.byref(r)))
        try:
            R = r.contents
            Q = int(R.n_queries)
            off = _view(R.offsets, Q + 1, np.uint64).copy()
            n = int(R.n_rows)
            rows = _copy_struct_array(R.rows, n, capi.Aln)
            cig = C.string_at(R.cigar_ops, int(R.n_cigar)) if R.n_cigar else b""
            stats = self._aln_stats(R)
        finally:
            self.lib.spb_free_align_result(r)
        return off, rows, cig, stats

    def search(self, qres, qoff, db_res_count, fetch_prefilter=False, **kw):
        """Fused predictmatch search step: prefilter -> align with the hit lists kept in HBM.
        kw: kmer_thr, max_seqs, min_ungapped_score, bin_size, comp_bias, evalue, gap_open, gap_extend, gumbel"""
        pk = {k: kw[k] for k in ("kmer_thr", "max_seqs", "min_ungapped_score", "bin_size", "comp_bias", "id_base", "exact_kmer") if k in kw}
        ak = {k: kw[k] for k in ("evalue", "gap_open", "gap_extend", "comp_bias", "gumbel", "byte_lanes", "word_lanes", "cov", "cov_mode",
                                 "min_seq_id", "min_aln_len", "seq_id_mode", "sw_mode", "max_accept", "max_rejected") if k in kw}
        self.set_queries(qres, qoff)
        pst = self.prefilter_resident(**pk)
        pref = self.prefilter_fetch() if fetch_prefilter else None
        self.pairs_from_prefilter()
        ast = self.align_resident(db_res_count, **ak)
        off, rows, cig, _ = self.align_fetch()
        return off, rows, cig, dict(prefilter=pst, align=ast), pref

    # ---- prefilter ---------------------------------------------------------------------------
    def prefilter(self, qres, qoff, kmer_thr=112, max_seqs=300, min_ungapped_score=15, bin_size=2, comp_bias=1,
                  id_base=0, exact_kmer=0):
        """-> (offsets uint64[Q+1], hits structured array (id, score, diag), stats dict)"""
        qres = np.ascontiguousarray(qres, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        Q = len(qoff) - 1
        p = capi.PrefilterParams(kmer_thr, max_seqs, min_ungapped_score, bin_size, comp_bias, id_base, 0, exact_kmer)
        r = C.POINTER(capi.PrefilterResult)()
        capi.check(self.lib.spb_prefilter(self.h, _u8(qres), _u64(qoff), Q, C.byref(p), C.byref(r)))
        try:
            R = r.contents
            off = _view(R.offsets, Q + 1, np.uint64).copy()
            n = int(R.n_hits)
            hits = _copy_struct_array(R.hits, n, capi.Hit)
            stats = self._pref_stats(R)
        finally:
            self.lib.spb_free_prefilter_result(r)
        return off, hits, stats

    # ---- align -------------------------------------------------------------------------------
    def set_align_targets(self, tres, toff, keys=None):
        tres = np.ascontiguousarray(tres, dtype=np.uint8)
        toff = np.ascontiguousarray(toff, dtype=np.uint64)
        kp = None
        if keys is not None:
            keys = np.ascontiguousarray(keys, dtype=np.uint32)
            kp = keys.ctypes.data_as(C.POINTER(C.c_uint32))
        self.aT = len(toff) - 1
        self.a_toff = toff
        self.a_keys = keys
        capi.check(self.lib.spb_align_set_targets(self.h, _u8(tres), _u64(toff), self.aT, kp))

    def align(self, qres, qoff, pair_off, pair_targets, db_res_count, evalue=200.0, gap_open=16, gap_extend=2,
              comp_bias=1, gumbel=None, byte_lanes=32, word_lanes=16, **criteria):
        """-> (offsets uint64[Q+1], rows structured array, cigar bytes, stats); criteria: see _align_params"""
        qres = np.ascontiguousarray(qres, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        pair_off = np.ascontiguousarray(pair_off, dtype=np.uint64)
        pair_targets = np.ascontiguousarray(pair_targets, dtype=np.uint32)
        Q = len(qoff) - 1
        p = self._align_params(db_res_count, evalue, gap_open, gap_extend, comp_bias, gumbel, byte_lanes, word_lanes, **criteria)
        r = C.POINTER(capi.AlignResult)()
        capi.check(self.lib.spb_align(self.h, _u8(qres), _u64(qoff), Q, _u64(pair_off),
                                      pair_targets.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(p), C.byref(r)))
        try:
            R = r.contents
            off = _view(R.offsets, Q + 1, np.uint64).copy()
            n = int(R.n_rows)
            rows = _copy_struct_array(R.rows, n, capi.Aln)
            cig = C.string_at(R.cigar_ops, int(R.n_cigar)) if R.n_cigar else b""
            stats = self._aln_stats(R)
        finally:
            self.lib.spb_free_align_result(r)
        return off, rows, cig, stats


    # ---- nucleotide re-alignment branch (predictmatch --perform-nucl-aln 1) ----------------------------------
    def _set_nt(self, fn, seqs, start_codon):
        lens = np.fromiter((len(x) for x in seqs), dtype=np.uint64, count=len(seqs))
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        nt = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, dtype=np.uint8)
        sc = None if start_codon is None else np.ascontiguousarray(start_codon, dtype=np.uint8)
        capi.check(fn(self.h, _u8(np.ascontiguousarray(nt)), _u64(off), len(seqs), _u8(sc) if sc is not None else None))

    def set_nucl_targets(self, seqs, start_codon=None):
        """seqs: ASCII nucleotide ORFs (bytes) of the alignment targets, same order (the <db>_nucl_orf entries)"""
        self._set_nt(self.lib.spb_nucl_set_targets, seqs, start_codon)

    def set_nucl_queries(self, seqs, start_codon=None):
        self._set_nt(self.lib.spb_nucl_set_queries, seqs, start_codon)

    def nucl_rescore(self, db_res_count, rows=None, cigar=b"", gap_open=10, gap_extend=2, gumbel=None):
        """proteinaln2nucl (lib/mmseqs/src/util/proteinaln2nucl.cpp:84-175) on the device: rows=None re-scores the
        resident alignment result (row i belongs to row i of align_fetch), else the given rows (structured array of
        capi.Aln + their cigar ops).  -> (structured array of capi.NuclAln, kernel_ms)"""
        p = nucl_params(db_res_count, gap_open, gap_extend, gumbel)
        r = C.POINTER(capi.NuclAln)()
        ms = C.c_double(0)
        if rows is None:
            n = C.c_uint64(0)
            capi.check(self.lib.spb_nucl_rescore_resident(self.h, C.byref(p), C.byref(r), C.byref(n), C.byref(ms)))
            n = int(n.value)
        else:
            rows = np.ascontiguousarray(rows)
            n = len(rows)
            capi.check(self.lib.spb_nucl_rescore(self.h, C.byref(p), rows.ctypes.data_as(C.POINTER(capi.Aln)), n, bytes(cigar),
                                                 len(cigar), C.byref(r)))
        try:
            out = _copy_struct_array(r, n, capi.NuclAln)
        finally:
            self.lib.spb_free_nucl_result(self.h, r)
        return out, float(ms.value)


# ALP parameters of nucleotide.out with gap open 10 / extend 2 (what proteinaln2nucl uses in predictmatch); like the
# VTML40 set they are estimated by the reference at start-up (EvalueComputation.h:48-154, seed 42) and are an input here
GUMBEL_NUCL_10_2 = (0.63361984397325744, 0.42150719938226583, 0.68666204026925159, 0.18923575026627315, 0.71225621255786731,
                    2.9058945364387716, 0.68666204026925159, 0.18923575026627315, 0.71225621255786731, 2.9058945364387716,
                    0.7125673897037631, 2.8984262849372726)


def nucl_params(db_res_count, gap_open=10, gap_extend=2, gumbel=None):
    """NucleotideMatrix(nucleotide.out, 1.0, 0.0) + its letter mapping (M/commons/NucleotideMatrix.cpp:18-63)"""
    if gumbel is None:
        if (gap_open, gap_extend) != (10, 2):
            raise ValueError("no ALP parameter set for nucleotide.out %d/%d: pass gumbel=" % (gap_open, gap_extend))
        gumbel = GUMBEL_NUCL_10_2
    p = capi.NuclParams()
    p.gap_open, p.gap_extend, p.db_res_count = gap_open, gap_extend, int(db_res_count)
    p.gumbel = capi.Gumbel(*gumbel)
    groups = {"A": 0, "C": 1, "M": 1, "Y": 1, "H": 1, "T": 2, "U": 2, "W": 2, "G": 3, "K": 3, "B": 3, "D": 3, "V": 3, "R": 3, "S": 3}
    for c in range(256):
        p.letter[c] = groups.get(chr(c).upper(), 4) if c < 128 else 4
    for i in range(5):
        for j in range(5):
            p.sub[i * 5 + j] = 2 if (i == j and i < 4) else -3
    return p


def format_nucl_rows(nrows, rows, cigar, keys=None):
    """Matcher::resultToBuffer of the rows proteinaln2nucl writes: the protein backtrace with every run length tripled"""
    out = []
    for n, r in zip(nrows, rows):
        if int(n["err"]):
            raise capi.SpbError("Alignment contains unalignable character" if int(n["err"]) == 1 else
                                "nucleotide re-alignment: row outside its nucleotide entry (err %d)" % int(n["err"]))
        t = int(r["target"])
        ops = cigar[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])].decode()
        bt = "".join("%d%s" % (3 * int(c), o) for c, o in re.findall(r"(\d+)([MID])", compress_cigar(ops))) if ops else ""
        out.append("%d\t%d\t%s\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d\t%s" % (
            int(keys[t]) if keys is not None else t, int(n["bits"]), format_seq_id(n["seq_id"]), float(n["eval"]),
            int(n["q_start"]), int(n["q_end"]), int(n["q_len"]), int(n["t_start"]), int(n["t_end"]), int(n["t_len"]), bt))
    return out


def combine_prot_nucl_rows(prot_rows, nucl_rows):
    """combineprotnuclaln (src/util/CombineProtNuclAln.cpp:55-95) for one query entry, on text rows: per protein row the
    last nucleotide row of the same target key replaces the sequence identity and gives the updated E-value
    min(sqrt(protE * nuclE), nuclE), both re-read from their printed form and printed with %.3E (SSTR(double))."""
    import math
    out = []
    for pr in prot_rows:
        f = pr.split("\t")
        upd, sid = 3.40282346638528859811704183484516925e+38, 0.0   # FLT_MAX, 0
        for nr in nucl_rows:
            g = nr.split("\t")
            if int(g[0]) != int(f[0]):
                continue
            sid = float(g[2])
            pe, ne = float(f[3]), float(g[3])
            lp = math.log(pe) if pe > 0 else -math.inf
            ln = math.log(ne) if ne > 0 else -math.inf
            upd = math.exp(lp * 0.5 + ln * 0.5) if (lp * 0.5 + ln * 0.5) < ln else ne
        out.append("\t".join(f[:2] + ["%.3E" % sid, "%.3E" % upd] + f[4:]))
    return out


def format_orf_headers(orfs, nt_lens, keys=None):
    """Orf::writeOrfHeader (M/commons/Orf.cpp:440-457) as extractorfs calls it (M/util/extractorfs.cpp:93-99):
    contig key \t from(+|-)|from-to| [\t incompleteStart | incompleteEnd << 1]"""
    out = []
    for o in orfs:
        c = int(o["contig"])
        L = int(nt_lens[c])
        f, t = int(o["from_"]), int(o["to"])
        if int(o["strand"]) == 1:
            f, t = (L - 1) - f, (L - 1) - t
        flags = int(o["incomplete_start"]) | (int(o["incomplete_end"]) << 1)
        out.append("%d\t%d%s%d%s" % (int(keys[c]) if keys is not None else c, f, "+" if f < t else "-", abs(f - t),
                                    ("\t%d" % flags) if flags else ""))
    return out


def format_prefilter_rows(hits, keys=None):
    """QueryMatcher::prefilterHitToBuffer (M/prefiltering/QueryMatcher.h:114-126): key \\t score \\t diagonal"""
    return ["%d\t%d\t%d" % (int(keys[h["id"]]) if keys is not None else int(h["id"]), int(h["score"]), int(h["diag"]))
            for h in hits]


def compress_cigar(ops):
    """Matcher::compressAlignment (M/alignment/Matcher.cpp:211-230): always emits counts."""
    out = []
    state, count = "M", 0
    for ch in ops:
        if ch != state:
            out.append("%d%s" % (count, state))
            state, count = ch, 1
        else:
            count += 1
    out.append("%d%s" % (count, state))
    return "".join(out)


def format_seq_id(x):
    """Util::fastSeqIdToBuffer (M/commons/Util.cpp:278-307) incl. the 4-character "1.00" quirk."""
    x = np.float32(x)
    if x == np.float32(1.0):
        return "1.00"
    s = "0."
    if float(x) < 0.10:
        s += "0"
    if float(x) < 0.01:
        s += "0"
    return s + str(int(np.float32(x * np.float32(1000))))


def format_align_rows(rows, cigar, q_lens, t_lens, keys=None, backtrace=True):
    """Matcher::resultToBuffer (M/alignment/Matcher.cpp:323-370); backtrace=False: the ten columns of `--alignment-mode 3`
    without `-a`"""
    out = []
    for r in rows:
        t = int(r["target"])
        ops = cigar[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])].decode()
        row = "%d\t%d\t%s\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d" % (
            int(keys[t]) if keys is not None else t, int(r["bits"]), format_seq_id(r["seq_id"]), float(r["eval"]),
            int(r["q_start"]), int(r["q_end"]), int(q_lens[int(r["query"])]), int(r["t_start"]), int(r["t_end"]),
            int(t_lens[t]))
        out.append(row + "\t" + (compress_cigar(ops) if ops else "") if backtrace else row)
    return out
