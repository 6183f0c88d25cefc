// index.cu -- HBM-resident target index: masked residue store, k-mer -> (id,pos) postings, 3-mer similarity rows.
//
// Replaces IndexBuilder::fillDatabase / IndexTable (M/prefiltering/IndexBuilder.cpp:53-256, IndexTable.h)
// and ExtendedSubstitutionMatrix::calcScoreMatrix (M/prefiltering/ExtendedSubstitutionMatrix.cpp:20-69).
//
// Layout (all resident for the life of the context):
//   tres   u8 [n_res]          masked residues, targets back to back (SequenceLookup)
//   toff   u32[T+1]
//   koff   u32[20^6+1]         CSR offsets, k-mer index = sum a_i 20^i (first residue lowest digit, Indexer.h:20-83)
//   post   u64[n_post]         (id << 16) | pos, sorted by (kmer, id); one posting per (kmer, target): smallest pos
//   direct u64x2[20^6]         packed per-k-mer entry (up to two postings inline, else len|offset) -- the table the
//                              lookup kernel of the prefilter gathers from
//   kbits  u32[20^6 / 32]      occupancy bit per k-mer (L2-resident pre-check of the gather)
//   row3   u32[8000][8000]     3-mer neighbours by score desc, ties in cartesian order: (score+256) << 16 | index
//   cnt3   u16[8000][512]      cnt3[x][v+256] = #neighbours of x with score >= v  (O(1) list lengths)
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace spb {

// ---------------------------------------------------------------------------------------------
// 3-mer similarity rows: one CTA per 3-mer x, bitonic sort of 8192 keys in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row3_kernel(const Matrices *__restrict__ mat, uint32_t *__restrict__ row3,
                                                   uint16_t *__restrict__ cnt3) {
    __shared__ uint32_t keys[8192];
    __shared__ int hist[SPB_CNT3_W];
    __shared__ short sub[SPB_ALPHABET * SPB_ALPHABET];
    const int x = blockIdx.x;
    for (int i = threadIdx.x; i < SPB_ALPHABET * SPB_ALPHABET; i += blockDim.x) sub[i] = mat->seed[i];
    for (int i = threadIdx.x; i < SPB_CNT3_W; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int x0 = x % 20, x1 = (x / 20) % 20, x2 = x / 400;
    for (int c = threadIdx.x; c < 8192; c += blockDim.x) {
        uint32_t key = 0xFFFFFFFFu;
        if (c < SPB_N3) {
            int y0 = c / 400, y1 = (c / 20) % 20, y2 = c % 20;  // cartesian order: y0 slowest
            int s = sub[x0 * SPB_ALPHABET + y0] + sub[x1 * SPB_ALPHABET + y1] + sub[x2 * SPB_ALPHABET + y2];
            key = ((uint32_t)(255 - s) << 13) | (uint32_t)c;      // score desc, then cartesian order (stable_sort)
            atomicAdd(&hist[s + SPB_SCORE_OFF], 1);
        }
        keys[c] = key;
    }
    block_bitonic_sort(keys, 8192, [](uint32_t a, uint32_t b) { return a < b; });
    for (int z = threadIdx.x; z < SPB_N3; z += blockDim.x) {
        uint32_t key = keys[z];
        int s = 255 - (int)(key >> 13);
        int c = key & 8191;
        int y0 = c / 400, y1 = (c / 20) % 20, y2 = c % 20;
        uint32_t idx = y0 + 20 * y1 + 400 * y2;
        row3[(size_t)x * SPB_N3 + z] = ((uint32_t)(s + SPB_SCORE_OFF) << 16) | idx;
    }
    // suffix sums: cnt[v] = #scores >= v
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int v = SPB_CNT3_W - 1; v >= 0; v--) {
            acc += hist[v];
            hist[v] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SPB_CNT3_W; i += blockDim.x) cnt3[(size_t)x * SPB_CNT3_W + i] = (uint16_t)hist[i];
}

// ---------------------------------------------------------------------------------------------
// k-mer extraction: one warp per target, lane = window start; key = kmer << 32 | id, value = pos.
// Written at the global residue position so that the input of the stable sort is ordered by (id, pos).
// IndexTable::addKmerCount / addSequence (IndexTable.h:130-170,341-394): skip windows with X, skip if the
// self score (diagonal truncated to char, IndexBuilder.cpp:8-19) is below the threshold.
// ---------------------------------------------------------------------------------------------
struct Pattern { int span; int pos[6]; };

__global__ void __launch_bounds__(256) extract_kmers_kernel(const uint8_t *__restrict__ tres, const uint32_t *__restrict__ toff,
                                                            uint32_t T, Pattern pat, int kmerThr, const Matrices *__restrict__ mat,
                                                            uint64_t *__restrict__ keys, uint16_t *__restrict__ vals,
                                                            unsigned long long *__restrict__ nValid) {
    __shared__ signed char selfScore[SPB_ALPHABET];
    if (threadIdx.x < SPB_ALPHABET) selfScore[threadIdx.x] = (signed char)mat->seed[threadIdx.x * SPB_ALPHABET + threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    unsigned long long localValid = 0;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warpsPerGrid) {
        const uint32_t o = toff[t];
        const int len = (int)(toff[t + 1] - o);
        const int nw = len - pat.span + 1;
        for (int w0 = 0; w0 < len; w0 += 32) {
            const int w = w0 + lane;
            if (w >= len) break;
            uint64_t key = ~0ull;
            if (w < nw && w < 65536) {
                uint32_t k = 0, mul = 1;
                int self = 0;
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    uint32_t a = tres[o + w + pat.pos[i]];
                    ok = ok && (a < SPB_X);
                    self += selfScore[a < SPB_ALPHABET ? a : SPB_X];
                    k += a * mul;
                    mul *= 20;
                }
                if (ok && !(kmerThr > 0 && self < kmerThr)) {
                    key = ((uint64_t)k << 32) | t;
                    localValid++;
                }
            }
            keys[o + w] = key;
            vals[o + w] = (uint16_t)w;
        }
    }
    // warp-aggregate the counter
    for (int d = 16; d > 0; d >>= 1) localValid += __shfl_down_sync(0xffffffffu, localValid, d);
    if (lane == 0 && localValid) atomicAdd(nValid, localValid);
}

// flag the first element of every (kmer, id) run among the M valid sorted entries
__global__ void flag_first_kernel(const uint64_t *__restrict__ keys, uint64_t M, uint32_t *__restrict__ flags) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void scatter_postings_kernel(const uint64_t *__restrict__ keys, const uint16_t *__restrict__ vals,
                                        const uint32_t *__restrict__ flags, const uint32_t *__restrict__ excl, uint64_t M,
                                        uint64_t *__restrict__ post, uint32_t *__restrict__ pk) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || !flags[i]) return;
    uint32_t d = excl[i];
    uint64_t key = keys[i];
    post[d] = ((key & 0xFFFFFFFFull) << 16) | vals[i];
    pk[d] = (uint32_t)(key >> 32);
}

// CSR offsets from the sorted k-mer column: koff[k] = first posting with kmer >= k
__global__ void fill_offsets_kernel(const uint32_t *__restrict__ pk, uint64_t nPost, uint32_t *__restrict__ koff,
                                    unsigned long long *__restrict__ nNonEmpty) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nPost) return;
    uint32_t from = (i == 0) ? 0u : pk[i - 1] + 1u;
    uint32_t to = (i == nPost) ? SPB_NK : pk[i];  // inclusive
    if (i < nPost && (i == 0 || pk[i - 1] != pk[i])) atomicAdd(nNonEmpty, 1ull);
    for (uint32_t k = from; k <= to && k <= SPB_NK; k++) koff[k] = (uint32_t)i;
}

// Packed direct table read by the warp-per-query prefilter: one 8-byte entry per k-mer.
//   x = 0                                  no target contains the k-mer
//   x = 1 << 63 | id << 16 | pos, y = 0    exactly one target: the posting itself (no second memory access)
//   x, y = 1 << 63 | posting               exactly two targets: both postings inline (ascending id, like the list)
//   x = len << 32 | offset                 len >= 3 postings at post[offset ..)
// kbits: one bit per k-mer (set = some target contains it), 8 MB, L2-resident: the lookup kernel asks it first, so the ~46 % of
// the similar k-mers that no target contains never touch the 1 GB table in DRAM.  SPB_NK is a multiple of 32 and of the block size.
__global__ void build_direct_kernel(const uint32_t *__restrict__ koff, const uint64_t *__restrict__ post,
                                    ulonglong2 *__restrict__ direct, uint32_t *__restrict__ kbits) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0;
    if (k < SPB_NK) {
        const uint32_t o0 = koff[k];
        len = koff[k + 1] - o0;
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        if (len == 1) v.x = (1ull << 63) | post[o0];
        else if (len == 2) { v.x = (1ull << 63) | post[o0]; v.y = (1ull << 63) | post[o0 + 1]; }
        else if (len > 2) v.x = ((uint64_t)len << 32) | o0;
        direct[k] = v;
    }
    const unsigned m = __ballot_sync(0xffffffffu, len != 0);
    if ((threadIdx.x & 31) == 0 && k < SPB_NK) kbits[k >> 5] = m;
}

__global__ void export_postings_kernel(const uint32_t *__restrict__ koff, const uint64_t *__restrict__ post,
                                       uint32_t *__restrict__ kmer, uint32_t *__restrict__ id, uint16_t *__restrict__ pos) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= SPB_NK) return;
    for (uint32_t z = koff[k]; z < koff[k + 1]; z++) {
        kmer[z] = k;
        id[z] = (uint32_t)(post[z] >> 16);
        pos[z] = (uint16_t)(post[z] & 0xFFFF);
    }
}

// copy of k-mer range [k0, k0 + n) of the direct table with the list offsets re-based to the shard's own posting slice
__global__ void shard_copy_kernel(const ulonglong2 *__restrict__ direct, uint32_t k0, uint32_t n, uint32_t postBase,
                                  ulonglong2 *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ulonglong2 v = direct[k0 + k];
    if (v.x != 0ull && !(v.x >> 63)) v.x = (v.x & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)v.x - postBase);
    out[k] = v;
}

static int parse_pattern(const char *p, Pattern *out) {
    int span = (int)strlen(p), k = 0;
    if (span < 6 || span > 32) return 1;
    for (int i = 0; i < span; i++) {
        if (p[i] == '1') {
            if (k >= 6) return 1;
            out->pos[k++] = i;
        } else if (p[i] != '0') {
            return 1;
        }
    }
    out->span = span;
    return k == 6 ? 0 : 1;
}

}  // namespace spb

using namespace spb;

extern "C" int spb_index_build(spb_ctx *c, const uint8_t *residues, const uint64_t *offsets, uint32_t T,
                               const char *pattern, int kmer_thr) {
    SPB_CHECK(!c || !c->is_lane, "spb_index_build: a lane context borrows its parent's index / targets");
    SPB_CHECK(c && residues && offsets && pattern, "spb_index_build: NULL argument");
    SPB_CHECK(c->have_matrices, "spb_index_build: call spb_set_matrices first");
    SPB_CUDA(cudaSetDevice(c->device));
    Pattern pat;
    SPB_CHECK(parse_pattern(pattern, &pat) == 0,
              "spb_index_build: spaced k-mer pattern '%s' must consist of 0/1 with exactly six 1 (k = 6 only)", pattern);
    const uint64_t N = offsets[T];
    SPB_CHECK(N < 0xFFFFFFF0ull, "spb_index_build: %llu residues exceed the 32-bit offset space of one shard; split the targets",
              (unsigned long long)N);
    SPB_CHECK(T < (1u << 28), "spb_index_build: more than 2^28 targets in one shard; split the targets");
    c->have_index = false;
    c->T = T;
    c->n_res = N;
    c->span = pat.span;
    for (int i = 0; i < 6; i++) c->pat[i] = pat.pos[i];
    cudaStream_t st = c->stream;
    SPB_CUDA(cudaEventRecord(c->ev0, st));

    // ---- residues + offsets
    if (c->d_tres.ensure(N + 64) || c->d_toff.ensure((size_t)(T + 1) * 4)) return 1;
    SPB_CUDA(cudaMemcpyAsync(c->d_tres.p, residues, N, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaMemsetAsync((uint8_t *)c->d_tres.p + N, SPB_X, 64, st));
    std::vector<uint32_t> off32((size_t)T + 1);
    for (uint32_t i = 0; i <= T; i++) {
        SPB_CHECK(offsets[i] <= N && (i == 0 || offsets[i] >= offsets[i - 1]), "spb_index_build: offsets not monotone at %u", i);
        off32[i] = (uint32_t)offsets[i];
    }
    SPB_CUDA(cudaMemcpyAsync(c->d_toff.p, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice, st));

    // ---- 3-mer rows
    if (c->d_row3.ensure((size_t)SPB_N3 * SPB_N3 * 4) || c->d_cnt3.ensure((size_t)SPB_N3 * SPB_CNT3_W * 2)) return 1;
    row3_kernel<<<SPB_N3, 256, 0, st>>>(c->d_mat.as<Matrices>(), c->d_row3.as<uint32_t>(), c->d_cnt3.as<uint16_t>());
    SPB_CUDA(cudaGetLastError()); c->n_launches++;

    // ---- postings
    const size_t Np = (size_t)(N > 0 ? N : 1);
    DevBuf keysA, keysB, valsA, valsB, cubTmp, counters;
    int rc = 0;
    auto cleanup = [&]() { keysA.release(); keysB.release(); valsA.release(); valsB.release(); cubTmp.release(); counters.release(); };
    if (keysA.ensure(Np * 8) || keysB.ensure(Np * 8) || valsA.ensure(Np * 2) || valsB.ensure(Np * 2) || counters.ensure(64)) {
        cleanup();
        return 1;
    }
#define IDX_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            spb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            cleanup();                                                                            \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)
    IDX_CUDA(cudaMemsetAsync(counters.p, 0, 64, st));
    IDX_CUDA(cudaMemsetAsync(keysA.p, 0xFF, Np * 8, st));
    unsigned long long *d_nValid = counters.as<unsigned long long>();
    unsigned long long *d_nNonEmpty = d_nValid + 1;
    if (T > 0) {
        int blocks = (int)((T + 7) / 8 < 148 * 16 ? (T + 7) / 8 : 148 * 16);
        extract_kmers_kernel<<<blocks, 256, 0, st>>>(c->d_tres.as<uint8_t>(), c->d_toff.as<uint32_t>(), T, pat, kmer_thr,
                                                     c->d_mat.as<Matrices>(), keysA.as<uint64_t>(), valsA.as<uint16_t>(), d_nValid);
        IDX_CUDA(cudaGetLastError()); c->n_launches++;
    }
    unsigned long long M = 0;
    IDX_CUDA(cudaMemcpyAsync(&M, d_nValid, 8, cudaMemcpyDeviceToHost, st));
    IDX_CUDA(cudaStreamSynchronize(st));
    uint64_t nPost = 0;
    if (c->d_koff.ensure(((size_t)SPB_NK + 2) * 4)) { cleanup(); return 1; }
    if (M > 0) {
        size_t tmpBytes = 0;
        // stable LSD radix sort on (kmer, id); invalid entries (all ones) sort last
        IDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysA.as<uint64_t>(), keysB.as<uint64_t>(), valsA.as<uint16_t>(),
                                                 valsB.as<uint16_t>(), (int64_t)N, 0, 64, st));
        if (cubTmp.ensure(tmpBytes)) { cleanup(); return 1; }
        IDX_CUDA(cub::DeviceRadixSort::SortPairs(cubTmp.p, tmpBytes, keysA.as<uint64_t>(), keysB.as<uint64_t>(), valsA.as<uint16_t>(),
                                                 valsB.as<uint16_t>(), (int64_t)N, 0, 64, st));
        // keysA is free now: reuse as flags (u32[M]) + exclusive sums (u32[M])
        uint32_t *flags = keysA.as<uint32_t>();
        uint32_t *excl = flags + M;
        unsigned gridM = (unsigned)((M + 255) / 256);
        flag_first_kernel<<<gridM, 256, 0, st>>>(keysB.as<uint64_t>(), M, flags);
        IDX_CUDA(cudaGetLastError()); c->n_launches++;
        size_t tmp2 = 0;
        IDX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, flags, excl, (int64_t)M, st));
        if (cubTmp.ensure(tmp2)) { cleanup(); return 1; }
        IDX_CUDA(cub::DeviceScan::ExclusiveSum(cubTmp.p, tmp2, flags, excl, (int64_t)M, st));
        uint32_t lastFlag = 0, lastExcl = 0;
        IDX_CUDA(cudaMemcpyAsync(&lastFlag, flags + (M - 1), 4, cudaMemcpyDeviceToHost, st));
        IDX_CUDA(cudaMemcpyAsync(&lastExcl, excl + (M - 1), 4, cudaMemcpyDeviceToHost, st));
        IDX_CUDA(cudaStreamSynchronize(st));
        nPost = (uint64_t)lastFlag + lastExcl;
        if (c->d_post.ensure((size_t)nPost * 8 + 64)) { cleanup(); return 1; }
        // k-mer column of the postings goes to valsA/keys scratch: reuse valsA (u16[N] = 2N bytes) is too small for u32; use cubTmp
        if (cubTmp.ensure((size_t)nPost * 4 + 64)) { cleanup(); return 1; }
        uint32_t *pk = cubTmp.as<uint32_t>();
        scatter_postings_kernel<<<gridM, 256, 0, st>>>(keysB.as<uint64_t>(), valsB.as<uint16_t>(), flags, excl, M,
                                                       c->d_post.as<uint64_t>(), pk);
        IDX_CUDA(cudaGetLastError()); c->n_launches++;
        unsigned gridP = (unsigned)((nPost + 1 + 255) / 256);
        fill_offsets_kernel<<<gridP, 256, 0, st>>>(pk, nPost, c->d_koff.as<uint32_t>(), d_nNonEmpty);
        IDX_CUDA(cudaGetLastError()); c->n_launches++;
    } else {
        if (c->d_post.ensure(64)) { cleanup(); return 1; }
        IDX_CUDA(cudaMemsetAsync(c->d_koff.p, 0, ((size_t)SPB_NK + 2) * 4, st));
    }
    if (c->d_direct.ensure((size_t)SPB_NK * 16) || c->d_kbits.ensure((size_t)(SPB_NK / 32) * 4 + 64)) { cleanup(); return 1; }
    build_direct_kernel<<<(SPB_NK + 255) / 256, 256, 0, st>>>(c->d_koff.as<uint32_t>(), c->d_post.as<uint64_t>(), c->d_direct.as<ulonglong2>(),
                                                              c->d_kbits.as<uint32_t>());
    IDX_CUDA(cudaGetLastError()); c->n_launches++;
    unsigned long long nonEmpty = 0;
    IDX_CUDA(cudaMemcpyAsync(&nonEmpty, d_nNonEmpty, 8, cudaMemcpyDeviceToHost, st));
    IDX_CUDA(cudaEventRecord(c->ev1, st));
    IDX_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    IDX_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
#undef IDX_CUDA
    cleanup();
    (void)rc;
    c->n_post = nPost;
    c->n_kmers_nonempty = nonEmpty;
    c->build_ms = ms;
    c->n_shards = 1;
    c->shard_rank = 0;
    c->shard_R = SPB_NK;
    for (int i = 0; i < SPB_MAX_SHARDS; i++) { c->shard_direct[i] = nullptr; c->shard_post[i] = nullptr; }
    c->shard_direct[0] = c->d_direct.p;
    c->shard_post[0] = c->d_post.p;
    c->have_index = true;
    return 0;
}

extern "C" int spb_index_shard(spb_ctx *c, int rank, int world) {
    SPB_CHECK(c && !c->is_lane, "spb_index_shard: needs a root context");
    SPB_CHECK(c->have_index && c->n_shards == 1, "spb_index_shard: build the (whole) index first");
    SPB_CHECK(world >= 1 && world <= SPB_MAX_SHARDS && rank >= 0 && rank < world, "spb_index_shard: rank %d / world %d out of range (max %d)",
              rank, world, SPB_MAX_SHARDS);
    if (world == 1) return 0;
    SPB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    // contiguous k-mer ranges, a multiple of 32 k-mers so that occupancy words never straddle shards
    const uint32_t R = (uint32_t)((((uint64_t)SPB_NK + world - 1) / world + 31) / 32 * 32);
    const uint32_t k0 = std::min<uint64_t>((uint64_t)R * rank, SPB_NK), k1 = (uint32_t)std::min<uint64_t>((uint64_t)R * (rank + 1), SPB_NK);
    const uint32_t n = k1 - k0;
    uint32_t h_off[2] = {0, 0};
    SPB_CUDA(cudaMemcpyAsync(&h_off[0], c->d_koff.as<uint32_t>() + k0, 4, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaMemcpyAsync(&h_off[1], c->d_koff.as<uint32_t>() + k1, 4, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    const uint64_t nPostShard = h_off[1] - h_off[0];
    // plain cudaMalloc allocations (IPC handles are per allocation)
    spb::DevBuf dshard, pshard;
    if (dshard.ensure(std::max<size_t>((size_t)n, 1) * 16) || pshard.ensure(std::max<uint64_t>(nPostShard, 1) * 8 + 64)) return 1;
    if (n) {
        shard_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(c->d_direct.as<ulonglong2>(), k0, n, h_off[0], dshard.as<ulonglong2>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    if (nPostShard)
        SPB_CUDA(cudaMemcpyAsync(pshard.p, c->d_post.as<uint64_t>() + h_off[0], nPostShard * 8, cudaMemcpyDeviceToDevice, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    // the whole table goes away: the CSR offsets, the postings and the direct table live on as this shard only
    c->d_direct.release();
    c->d_post.release();
    c->d_koff.release();
    c->d_direct = dshard;
    c->d_post = pshard;
    c->n_shards = world;
    c->shard_rank = rank;
    c->shard_R = R;
    for (int i = 0; i < SPB_MAX_SHARDS; i++) { c->shard_direct[i] = nullptr; c->shard_post[i] = nullptr; }
    c->shard_direct[rank] = c->d_direct.p;
    c->shard_post[rank] = c->d_post.p;
    return 0;
}

extern "C" int spb_index_shard_export(spb_ctx *c, void *handles128) {
    SPB_CHECK(c && handles128, "spb_index_shard_export: NULL argument");
    SPB_CHECK(c->have_index && c->n_shards > 1 && !c->is_lane, "spb_index_shard_export: call spb_index_shard first");
    SPB_CUDA(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h[2];
    SPB_CUDA(cudaIpcGetMemHandle(&h[0], c->d_direct.p));
    SPB_CUDA(cudaIpcGetMemHandle(&h[1], c->d_post.p));
    memcpy(handles128, h, 128);
    return 0;
}

extern "C" int spb_index_shard_attach_ipc(spb_ctx *c, int peer, const void *handles128) {
    SPB_CHECK(c && handles128, "spb_index_shard_attach_ipc: NULL argument");
    SPB_CHECK(c->have_index && c->n_shards > 1 && !c->is_lane, "spb_index_shard_attach_ipc: call spb_index_shard first");
    SPB_CHECK(peer >= 0 && peer < c->n_shards && peer != c->shard_rank, "spb_index_shard_attach_ipc: peer %d out of range", peer);
    SPB_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h[2];
    memcpy(h, handles128, 128);
    void *pd = nullptr, *pp = nullptr;
    SPB_CUDA(cudaIpcOpenMemHandle(&pd, h[0], cudaIpcMemLazyEnablePeerAccess));
    SPB_CUDA(cudaIpcOpenMemHandle(&pp, h[1], cudaIpcMemLazyEnablePeerAccess));
    c->ipc_mapped.push_back(pd);
    c->ipc_mapped.push_back(pp);
    c->shard_direct[peer] = pd;
    c->shard_post[peer] = pp;
    return 0;
}

extern "C" int spb_index_shard_attach_ctx(spb_ctx *c, int peer, spb_ctx *pc) {
    SPB_CHECK(c && pc, "spb_index_shard_attach_ctx: NULL argument");
    SPB_CHECK(c->have_index && c->n_shards > 1 && !c->is_lane, "spb_index_shard_attach_ctx: call spb_index_shard first");
    SPB_CHECK(pc->have_index && pc->n_shards == c->n_shards && pc->shard_rank == peer && peer != c->shard_rank,
              "spb_index_shard_attach_ctx: the peer context does not hold shard %d of %d", peer, c->n_shards);
    if (pc->device != c->device) {
        SPB_CUDA(cudaSetDevice(c->device));
        int can = 0;
        SPB_CUDA(cudaDeviceCanAccessPeer(&can, c->device, pc->device));
        SPB_CHECK(can, "spb_index_shard_attach_ctx: device %d cannot access device %d", c->device, pc->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(pc->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        SPB_CHECK(e == cudaSuccess, "spb_index_shard_attach_ctx: cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    }
    c->shard_direct[peer] = pc->d_direct.p;
    c->shard_post[peer] = pc->d_post.p;
    return 0;
}

extern "C" int spb_index_stats_get(spb_ctx *c, spb_index_stats *out) {
    SPB_CHECK(c && out, "spb_index_stats_get: NULL argument");
    SPB_CHECK(c->have_index, "spb_index_stats_get: no index built");
    out->n_targets = c->T;
    out->n_residues = c->n_res;
    out->n_postings = c->n_post;
    out->n_kmers_nonempty = c->n_kmers_nonempty;
    out->bytes_device = c->d_tres.cap + c->d_toff.cap + c->d_koff.cap + c->d_post.cap + c->d_direct.cap + c->d_kbits.cap + c->d_row3.cap + c->d_cnt3.cap;
    out->build_ms = c->build_ms;
    return 0;
}

extern "C" int spb_index_export(spb_ctx *c, uint32_t *kmer, uint32_t *id, uint16_t *pos) {
    SPB_CHECK(c && kmer && id && pos, "spb_index_export: NULL argument");
    SPB_CHECK(c->have_index, "spb_index_export: no index built");
    SPB_CHECK(c->n_shards == 1 && c->d_koff.p, "spb_index_export: the CSR view is gone once the table is sharded (spb_index_shard)");
    SPB_CUDA(cudaSetDevice(c->device));
    if (c->n_post == 0) return 0;
    DevBuf dk, di, dp;
    if (dk.ensure(c->n_post * 4) || di.ensure(c->n_post * 4) || dp.ensure(c->n_post * 2)) return 1;
    export_postings_kernel<<<(SPB_NK + 255) / 256, 256, 0, c->stream>>>(c->d_koff.as<uint32_t>(), c->d_post.as<uint64_t>(),
                                                                        dk.as<uint32_t>(), di.as<uint32_t>(), dp.as<uint16_t>());
    cudaError_t e = cudaGetLastError(); c->n_launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(kmer, dk.p, c->n_post * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(id, di.p, c->n_post * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos, dp.p, c->n_post * 2, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dk.release(); di.release(); dp.release();
    SPB_CHECK(e == cudaSuccess, "spb_index_export: %s", cudaGetErrorString(e));
    return 0;
}
