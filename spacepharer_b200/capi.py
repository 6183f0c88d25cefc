"""ctypes binding of libspb200.so (include/spb200.h).  The library is built in-tree by `make`
(or __graft_entry__.build()); importing this module never falls back to a CPU implementation."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspb200.so")
ALPH = 21


class Gumbel(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "lambda_", "K", "a_I", "b_I", "alpha_I", "beta_I", "a_J", "b_J", "alpha_J", "beta_J", "sigma", "tau")]


class IndexStats(C.Structure):
    _fields_ = [("n_targets", C.c_uint64), ("n_residues", C.c_uint64), ("n_postings", C.c_uint64),
                ("n_kmers_nonempty", C.c_uint64), ("bytes_device", C.c_uint64), ("build_ms", C.c_double)]


class PrefilterParams(C.Structure):
    _fields_ = [("kmer_thr", C.c_int32), ("max_seqs", C.c_int32), ("min_diag_score", C.c_int32),
                ("bin_size", C.c_int32), ("comp_bias", C.c_int32), ("id_base", C.c_uint32),
                ("emit_trunc_keys", C.c_int32), ("exact_kmer", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("id", C.c_uint32), ("score", C.c_int32), ("diag", C.c_int32)]


class PrefilterResult(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("offsets", C.POINTER(C.c_uint64)), ("hits", C.POINTER(Hit)),
                ("n_hits", C.c_uint64), ("n_kmers", C.c_uint64), ("n_db_matches", C.c_uint64),
                ("n_events", C.c_uint64), ("n_slow_path", C.c_uint64), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double), ("algo_bytes", C.c_uint64), ("expand_ms", C.c_double),
                ("lookup_ms", C.c_double), ("bookkeep_ms", C.c_double)]


class AlignParams(C.Structure):
    _fields_ = [("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("eval_thr", C.c_double),
                ("db_res_count", C.c_uint64), ("comp_bias", C.c_int32), ("byte_lanes", C.c_int32),
                ("word_lanes", C.c_int32), ("gumbel", Gumbel),
                ("cov_thr", C.c_float), ("cov_mode", C.c_int32), ("seq_id_thr", C.c_float), ("aln_len_thr", C.c_int32),
                ("seq_id_mode", C.c_int32), ("sw_mode", C.c_int32), ("max_accept", C.c_int32), ("max_rejected", C.c_int32)]


class Aln(C.Structure):
    _fields_ = [("query", C.c_uint32), ("target", C.c_uint32), ("score", C.c_int32), ("bits", C.c_int32),
                ("eval", C.c_double), ("q_start", C.c_int32), ("q_end", C.c_int32), ("t_start", C.c_int32),
                ("t_end", C.c_int32), ("n_ident", C.c_int32), ("aln_len", C.c_int32), ("seq_id", C.c_float),
                ("cigar_off", C.c_uint32), ("cigar_len", C.c_uint32), ("reserved", C.c_uint32)]


class AlignResult(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("offsets", C.POINTER(C.c_uint64)), ("rows", C.POINTER(Aln)),
                ("n_rows", C.c_uint64), ("cigar_ops", C.POINTER(C.c_char)), ("n_cigar", C.c_uint64),
                ("n_pairs", C.c_uint64), ("n_cells", C.c_uint64), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double)]


class ExhaustiveStats(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("n_targets", C.c_uint64), ("n_scored", C.c_uint64), ("n_cells", C.c_uint64),
                ("n_kept", C.c_uint64), ("n_unfiltered", C.c_uint64), ("n_pairs", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double)]


class Orf(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("from_", C.c_uint32), ("to", C.c_uint32), ("strand", C.c_uint8),
                ("incomplete_start", C.c_uint8), ("incomplete_end", C.c_uint8), ("pad", C.c_uint8)]


class FramesResult(C.Structure):
    _fields_ = [("n_sequences", C.c_uint32), ("n_frames", C.c_uint32), ("n_residues", C.c_uint64),
                ("offsets", C.POINTER(C.c_uint64)), ("residues", C.POINTER(C.c_ubyte)), ("orfs", C.POINTER(Orf)),
                ("kernel_ms", C.c_double)]


# every symbol include/spb200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "spb_ctx_create", "spb_ctx_destroy", "spb_last_error", "spb_version", "spb_ctx_stream", "spb_set_matrices",
    "spb_ctx_launches", "spb_index_build", "spb_index_stats_get", "spb_index_export", "spb_similar_kmers", "spb_prefilter",
    "spb_free_prefilter_result", "spb_align_set_targets", "spb_align", "spb_free_align_result", "spb_evalue",
    "spb_bitscore", "spb_tantan_mask", "spb_queries_set", "spb_prefilter_resident", "spb_prefilter_stats",
    "spb_prefilter_fetch", "spb_pairs_set", "spb_pairs_from_prefilter", "spb_align_resident", "spb_align_stats",
    "spb_align_fetch", "spb_prefilter_device_result", "spb_prefilter_device_trunc_keys", "spb_prefilter_merge_shards",
    "spb_pairs_exhaustive", "spb_align_exhaustive", "spb_spacer_frames", "spb_free_frames_result",
    "spb_ctx_create_lane", "spb_align_fetch_begin", "spb_align_fetch_wait",
    "spb_nucl_set_targets", "spb_nucl_set_queries", "spb_nucl_rescore", "spb_nucl_rescore_resident", "spb_free_nucl_result",
    "spb_device_count", "spb_index_shard", "spb_index_shard_export", "spb_index_shard_attach_ipc", "spb_index_shard_attach_ctx",
]

_lib = None


class NuclParams(C.Structure):
    _fields_ = [("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("db_res_count", C.c_uint64), ("gumbel", Gumbel),
                ("letter", C.c_uint8 * 256), ("sub", C.c_int16 * 25)]


class NuclAln(C.Structure):
    _fields_ = [("raw_score", C.c_int32), ("bits", C.c_int32), ("eval", C.c_double), ("q_start", C.c_int32), ("q_end", C.c_int32),
                ("t_start", C.c_int32), ("t_end", C.c_int32), ("q_len", C.c_uint32), ("t_len", C.c_uint32),
                ("n_ident", C.c_uint32), ("aln_len", C.c_uint32), ("seq_id", C.c_float), ("err", C.c_uint32)]


class SpbError(RuntimeError):
    pass


def load():
    """Load libspb200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpbError("libspb200.so not built: run `make` (or __graft_entry__.build()) in the repo root")
    L = C.CDLL(LIB_PATH)
    vp, u8p, u64p, u32p = C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.spb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.spb_ctx_create_lane.argtypes = [vp, C.POINTER(vp)]
    L.spb_ctx_destroy.argtypes = [vp]
    L.spb_ctx_destroy.restype = None
    L.spb_last_error.restype = C.c_char_p
    L.spb_version.restype = C.c_char_p
    L.spb_ctx_stream.argtypes = [vp]
    L.spb_ctx_stream.restype = vp
    L.spb_ctx_launches.argtypes = [vp]
    L.spb_ctx_launches.restype = C.c_uint64
    L.spb_set_matrices.argtypes = [vp, C.POINTER(C.c_int16), C.POINTER(C.c_double), C.POINTER(C.c_int16),
                                   C.POINTER(C.c_int16), C.POINTER(C.c_double)]
    L.spb_index_build.argtypes = [vp, u8p, u64p, C.c_uint32, C.c_char_p, C.c_int]
    L.spb_index_stats_get.argtypes = [vp, C.POINTER(IndexStats)]
    L.spb_index_export.argtypes = [vp, u32p, u32p, C.POINTER(C.c_uint16)]
    L.spb_similar_kmers.argtypes = [vp, u8p, C.c_int, u32p, C.c_uint32, u32p]
    L.spb_prefilter.argtypes = [vp, u8p, u64p, C.c_uint32, C.POINTER(PrefilterParams),
                                C.POINTER(C.POINTER(PrefilterResult))]
    L.spb_free_prefilter_result.argtypes = [C.POINTER(PrefilterResult)]
    L.spb_free_prefilter_result.restype = None
    L.spb_align_set_targets.argtypes = [vp, u8p, u64p, C.c_uint32, u32p]
    L.spb_align.argtypes = [vp, u8p, u64p, C.c_uint32, u64p, u32p, C.POINTER(AlignParams),
                            C.POINTER(C.POINTER(AlignResult))]
    L.spb_free_align_result.argtypes = [C.POINTER(AlignResult)]
    L.spb_free_align_result.restype = None
    L.spb_evalue.argtypes = [C.POINTER(Gumbel), C.c_double, C.c_double, C.c_double]
    L.spb_evalue.restype = C.c_double
    L.spb_bitscore.argtypes = [C.POINTER(Gumbel), C.c_double]
    L.spb_bitscore.restype = C.c_double
    L.spb_queries_set.argtypes = [vp, u8p, u64p, C.c_uint32]
    L.spb_prefilter_resident.argtypes = [vp, C.POINTER(PrefilterParams)]
    L.spb_prefilter_stats.argtypes = [vp, C.POINTER(PrefilterResult)]
    L.spb_prefilter_fetch.argtypes = [vp, C.POINTER(C.POINTER(PrefilterResult))]
    L.spb_pairs_set.argtypes = [vp, u64p, u32p]
    L.spb_pairs_from_prefilter.argtypes = [vp]
    L.spb_align_resident.argtypes = [vp, C.POINTER(AlignParams)]
    L.spb_align_stats.argtypes = [vp, C.POINTER(AlignResult)]
    L.spb_align_fetch.argtypes = [vp, C.POINTER(C.POINTER(AlignResult))]
    L.spb_align_fetch_begin.argtypes = [vp, C.POINTER(C.POINTER(AlignResult))]
    L.spb_align_fetch_wait.argtypes = [C.POINTER(AlignResult)]
    L.spb_prefilter_device_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), u64p]
    L.spb_prefilter_merge_shards.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, vp, C.c_uint32, C.c_int32]
    L.spb_prefilter_device_trunc_keys.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.spb_index_shard.argtypes = [vp, C.c_int, C.c_int]
    L.spb_index_shard_export.argtypes = [vp, C.c_char_p]
    L.spb_index_shard_attach_ipc.argtypes = [vp, C.c_int, C.c_char_p]
    L.spb_index_shard_attach_ctx.argtypes = [vp, C.c_int, vp]
    L.spb_pairs_exhaustive.argtypes = [vp, C.POINTER(AlignParams), C.POINTER(ExhaustiveStats)]
    L.spb_align_exhaustive.argtypes = [vp, u8p, u64p, C.c_uint32, C.POINTER(AlignParams),
                                       C.POINTER(C.POINTER(AlignResult)), C.POINTER(ExhaustiveStats)]
    L.spb_spacer_frames.argtypes = [vp, u8p, u64p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(FramesResult))]
    L.spb_free_frames_result.argtypes = [C.POINTER(FramesResult)]
    L.spb_free_frames_result.restype = None
    L.spb_nucl_set_targets.argtypes = [vp, u8p, u64p, C.c_uint32, u8p]
    L.spb_nucl_set_queries.argtypes = [vp, u8p, u64p, C.c_uint32, u8p]
    L.spb_nucl_rescore.argtypes = [vp, C.POINTER(NuclParams), C.POINTER(Aln), C.c_uint64, C.c_char_p, C.c_uint64,
                                   C.POINTER(C.POINTER(NuclAln))]
    L.spb_nucl_rescore_resident.argtypes = [vp, C.POINTER(NuclParams), C.POINTER(C.POINTER(NuclAln)), u64p, C.POINTER(C.c_double)]
    L.spb_free_nucl_result.argtypes = [vp, C.POINTER(NuclAln)]
    L.spb_free_nucl_result.restype = None
    L.spb_tantan_mask.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), u8p, u64p, C.c_uint32, C.c_int, u64p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SpbError(load().spb_last_error().decode(errors="replace"))
