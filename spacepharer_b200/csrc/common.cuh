// common.cuh -- shared context, error plumbing and block-level primitives of libspb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/spb200.h"

#define SPB_X 20
#define SPB_N3 8000           // 20^3
#define SPB_NK 64000000u      // 20^6
#define SPB_CNT3_W 512        // score-count table width: scores -256 .. 255
#define SPB_SCORE_OFF 256
#define SPB_MAX_SHARDS 8

namespace spb {

void set_error(const char *fmt, ...);

#define SPB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            spb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define SPB_CHECK(cond, ...)                \
    do {                                    \
        if (!(cond)) {                      \
            spb::set_error(__VA_ARGS__);    \
            return 1;                       \
        }                                   \
    } while (0)

// growable device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;   // read-only view of another context's buffer (spb_ctx_create_lane): never freed or grown here
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (borrowed) {
            set_error("a lane context cannot rebuild the index / targets it borrows from its parent");
            return 1;
        }
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return 1;
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p && !borrowed) cudaFree(p);
        p = nullptr;
        cap = 0;
        borrowed = false;
    }
    DevBuf view() const {
        DevBuf v = *this;
        v.borrowed = true;
        return v;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// pinned host staging blocks handed out with results (grow-only pool owned by the context)
struct PinBlock {
    void *p = nullptr;
    size_t cap = 0;
    bool used = false;
};

struct Matrices {
    int16_t seed[SPB_ALPHABET * SPB_ALPHABET];
    int16_t ungapped[SPB_ALPHABET * SPB_ALPHABET];
    int16_t align[SPB_ALPHABET * SPB_ALPHABET];
    double pback_seed[SPB_ALPHABET];
    double pback_align[SPB_ALPHABET];
};

}  // namespace spb

struct spb_ctx {
    int device = 0;
    bool is_lane = false;         // spb_ctx_create_lane: index / targets / matrices are borrowed from the parent, read-only
    int n_sm = 148;               // multiprocessors of the device (persistent grids)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    uint64_t n_launches = 0;      // hand-written kernels launched so far (library kernels such as cub's are not counted)
    bool have_matrices = false;
    spb::Matrices hm;             // host copy
    spb::DevBuf d_mat;            // device copy of spb::Matrices

    // ---- index (masked targets) ----
    bool have_index = false;
    uint32_t T = 0;
    uint64_t n_res = 0;
    int span = 0;
    int pat[6] = {0, 0, 0, 0, 0, 0};
    spb::DevBuf d_tres;           // u8[n_res + pad] masked residues
    spb::DevBuf d_toff;           // u32[T+1]
    spb::DevBuf d_koff;           // u32[20^6 + 1] CSR offsets
    spb::DevBuf d_post;           // u64[n_post]: id << 16 | pos
    spb::DevBuf d_direct;         // ulonglong2[20^6] packed direct table: up to two postings inline, else len << 32 | offset (index.cu)
    spb::DevBuf d_kbits;          // u32[20^6 / 32] occupancy bit per k-mer (pre-check of the gather, L2-resident)
    // k-mer-range shards of the table (spb_index_shard): shard r holds k-mers [r * shard_R, (r + 1) * shard_R); own shard in
    // d_direct / d_post (re-based), the others peer-mapped.  n_shards == 1: shard 0 is the whole table.
    int n_shards = 1, shard_rank = 0;
    uint32_t shard_R = SPB_NK;
    const void *shard_direct[SPB_MAX_SHARDS] = {nullptr};
    const void *shard_post[SPB_MAX_SHARDS] = {nullptr};
    std::vector<void *> ipc_mapped;       // cudaIpcOpenMemHandle results to close
    uint64_t n_post = 0, n_kmers_nonempty = 0;
    spb::DevBuf d_row3;           // u32[8000][8000]: (score + 256) << 16 | 3-mer index
    spb::DevBuf d_cnt3;           // u16[8000][512]: #entries with score >= v
    double build_ms = 0;

    // ---- alignment targets (unmasked) ----
    bool have_aln_targets = false;
    uint32_t aT = 0;
    uint64_t a_res = 0;
    uint64_t a_maxlen = 0;        // longest target (sort key width of the register SW kernel)
    spb::DevBuf d_ares;           // u8
    spb::DevBuf d_aoff;           // u64[aT+1]
    std::vector<uint64_t> h_aoff; // host copy (work ordering, row order)
    std::vector<uint32_t> h_akeys;
    spb::DevBuf d_akeys;          // u32[aT] DB keys (last tie-break of the row order)

    // ---- resident query batch ----
    bool have_queries = false;
    uint32_t Q = 0;
    uint64_t q_res = 0;
    uint32_t q_maxlen = 0, q_minlen = 0;
    spb::DevBuf d_q;              // u8[q_res]
    spb::DevBuf d_qoff;           // u32[Q+1]
    spb::DevBuf d_qoff64;         // u64[Q+1]
    spb::DevBuf d_qorder;         // u32[Q] prefilter processing order (longest first)
    std::vector<uint32_t> h_qorder, h_qlen;   // host copies: that order, query lengths (exhaustive-search tiling)

    // ---- resident prefilter result (query order) ----
    bool have_pref = false;
    uint64_t n_hits = 0;
    spb::DevBuf d_hits;           // spb_hit[n_hits]
    spb::DevBuf d_hits_alt;       // second hit buffer: merged lists of the target split are built here, then swapped
    spb::DevBuf d_hitoff;         // u64[Q+1]
    bool have_tkeys = false;
    spb::DevBuf d_hit_tkeys;      // u64[n_hits] truncation keys (spb_prefilter_params.emit_trunc_keys), query order
    spb::DevBuf s_tk;             // arena of the same before the compaction
    spb_prefilter_result pref_stats;   // counters of the last run (pointers unused)

    // ---- resident alignment input / result ----
    bool have_pairs = false, have_aln = false;
    uint64_t n_pairs = 0;
    uint64_t pairs_max_per_query = 0;   // upper bound of the pairs of one query (row-order kernels)
    uint32_t pref_max_seqs = 0;         // max-seqs of the resident prefilter result
    spb::DevBuf d_pairoff;        // u64[Q+1]
    spb::DevBuf d_pairT;          // u32[n_pairs] (work order)
    spb::DevBuf d_pairRank;       // u32[n_pairs]: position of the pair in the caller's list of its query (pair_rank_mode == 1)
    int pair_rank_mode = 0;       // 0: work order = list order (pairs from the resident prefilter), 1: d_pairRank, 2: unknown (exhaustive)
    spb::DevBuf d_pairQ;          // u32[n_pairs]
    spb::DevBuf d_pairout;        // PairOut[n_pairs]
    spb::DevBuf d_acc;            // u32[n_acc] accepted pair indices
    uint64_t n_acc = 0, n_cigar = 0;
    spb::DevBuf d_cig;            // char[n_cigar]
    spb::DevBuf d_rows;           // spb_aln[n_acc] result rows, per query in the order of Matcher::compareHits
    spb::DevBuf d_rowoff;         // u64[Q+1]
    spb_align_result aln_stats;

    // ---- nucleotide ORFs of the re-alignment branch (nuclaln.cu; ASCII, same ids as the amino-acid entries) ----
    bool have_nt_targets = false, have_nt_queries = false;
    uint32_t ntT = 0, ntQ = 0;
    spb::DevBuf d_tnt, d_tntoff, d_tntsc;   // u8[], u64[ntT+1], u8[ntT] (amino-acid entry starts with '*')
    spb::DevBuf d_qnt, d_qntoff, d_qntsc;
    spb::DevBuf d_nrows;                    // spb_nucl_aln[n] last result

    // ---- pipelined result fetch (spb_align_fetch_begin / _wait): device staging copy + D2H on a second stream ----
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_staged = nullptr, ev_stage_free = nullptr;
    bool stage_busy = false;
    spb::DevBuf d_st_rows, d_st_cig, d_st_off;

    // ---- pinned result staging ----
    std::vector<spb::PinBlock> pins;
    void *pin_acquire(size_t bytes);
    void pin_release(void *p);

    // ---- prefilter tile arenas (prefilter.cu: expand -> lookup -> bookkeeping) ----
    spb::DevBuf p_recs;           // u32[chunkCap][256] slot records
    spb::DevBuf p_ctab;           // uint2[chunkCap] (query, first stream-order number)
    spb::DevBuf p_ccnt;           // uint2[chunkCap] (first hit record, hits) per chunk
    spb::DevBuf p_harena;         // u64[] hit records, chunk after chunk
    spb::DevBuf p_meta;           // QMeta[Q]
    spb::DevBuf p_heavy;          // HeavyQ[]: window tables of the queries whose slot records are filled chunk by chunk
    spb::DevBuf p_retry;          // u32[Q] queries of the next round
    spb::DevBuf p_team2;          // u32[Q] queries the warp-per-query bookkeeping hands to the CTA launch
    cudaStream_t aux_stream = nullptr;   // the CTA-per-query bookkeeping runs beside the warp-per-query one
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> ev_phase;   // five per tile

    // ---- scratch (grow-only, reused across calls) ----
    spb::DevBuf s_a, s_b, s_c, s_d, s_e, s_f, s_g, s_h, s_i, s_j, s_k, s_l, s_m, s_n, s_o, s_p, s_q, s_cub;
};

namespace spb {

// installs a query batch as the resident one (spb_queries_set; residues may already live on the device)
int queries_install(spb_ctx *c, const uint8_t *q_residues, bool on_device, const uint64_t *q_offsets, uint32_t Q);

// ---- E-value on the device (lib/alp/sls_pvalues.cpp:366-541) ----
__device__ __forceinline__ double dev_evalue(const spb_gumbel &g, double y, double n, double m) {
    const double const_val = 0.3989422804014326779;  // 1/sqrt(2 pi)
    double m_li_y = m - (g.a_I * y + g.b_I);
    double vi_y = fmax(2.0 * g.alpha_I / g.lambda, g.alpha_I * y + g.beta_I);
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = sqrt_vi_y == 0.0 ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = 0.5 * erfc(-sqrt(0.5) * m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = n - (g.a_J * y + g.b_J);
    double vj_y = fmax(2.0 * g.alpha_J / g.lambda, g.alpha_J * y + g.beta_J);
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = sqrt_vj_y == 0.0 ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = 0.5 * erfc(-sqrt(0.5) * n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    double c_y = fmax(2.0 * g.sigma / g.lambda, g.sigma * y + g.tau);
    double area = p1 * p2 + c_y * (P_m_F * P_n_F);
    return g.K * exp(-g.lambda * y) * area;
}

// ---------------------------------------------------------------------------------------------
// block-level primitives (blockDim.x threads, multiple of 32, <= 1024)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan over the block; returns exclusive prefix, *total = block sum. wsum: int[33] in smem.
__device__ __forceinline__ int block_excl_scan(int v, int *wsum, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_incl_scan(v);
    __syncthreads();
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? wsum[lane] : 0;
        int si = warp_incl_scan(s);
        if (lane < nw) wsum[lane] = si - s;
        if (lane == 31) wsum[32] = si;
    }
    __syncthreads();
    int r = inc - v + wsum[wid];
    *total = wsum[32];
    return r;
}

// in-place bitonic sort of n_pow2 keys (ascending); caller pads with max sentinels. Works on shared or global.
template <class K, class Less>
__device__ __forceinline__ void block_bitonic_sort(K *a, int n_pow2, Less less) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                // index of the lower element of pair t for stride j
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = ((i & k) == 0);
                K x = a[i], y = a[p];
                bool sw = up ? less(y, x) : less(x, y);
                if (sw) { a[i] = y; a[p] = x; }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace spb
