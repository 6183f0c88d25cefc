// prefilter.cu -- k-mer prefilter for a batch of query frames, one CTA per query.
//
// Replaces QueryMatcher::matchQuery and callees for AA/AA, --diag-score 1
// (M/prefiltering/QueryMatcher.cpp:84-307,324-381,459-505; KmerGenerator.cpp:107-216;
//  CacheFriendlyOperations.cpp:146-235,311-341; UngappedAlignment.cpp:27-51,216-233,295-403).
//
// The reference consumes a sequential hit stream ordered by (window, similar-k-mer rank, posting order) and runs
// order-dependent per-target byte state machines over it.  Here every hit carries (id, window, rank); once the hits
// of a target are sorted by (window, rank) every reference rule is a comparison between neighbours:
//   double hit      : hit i is an event iff its 8-bit diagonal equals that of the previous hit of the same target,
//                     or -- first hit of a target -- equals 0 (quirk B: the byte array starts zeroed)
//   event de-dup    : an event is kept iff the previous EVENT of the same target has a different 8-bit diagonal
//   keepMaxElement  : per target the first kept event (stream order) reaching the target's best capped score
//   truncation      : order key (score desc, id & (BIN-1), window, rank, id) == order of the reference's two stable
//                     counting sorts over the bin-major event list
// Windows w >= P-1 (the last window) never contribute (quirk A, CacheFriendlyOperations.cpp:43).
//
// Almost every index hit belongs to a target that is hit once and can never be an event.  The kernel therefore
// enumerates the hit stream twice (the second time out of L1/L2): pass 1 only marks hash(id) in two shared-memory
// bitmaps ("seen once", "seen twice"); pass 2 keeps a hit iff its target was seen twice or its 8-bit diagonal is 0.
// Only that short list (a few dozen records instead of ~600) is sorted and scored.  Hash collisions merely keep a
// few harmless extra hits.  The rank key is (a << 16 | b), the position in the nested 3-mer x 3-mer enumeration.
#include <cub/cub.cuh>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace spb {

struct Pattern6 { int span; int pos[6]; };

struct Rec {            // one index hit
    uint64_t hi;        // id << 32 | window << 16 | position_j
    uint64_t lo;        // rank: outer index << 16 | inner index
};
__device__ __forceinline__ bool rec_less(const Rec &a, const Rec &b) {
    const uint64_t ka = a.hi >> 16, kb = b.hi >> 16;   // (id, window)
    return ka != kb ? ka < kb : a.lo < b.lo;
}
__device__ __forceinline__ uint32_t rec_id(const Rec &r) { return (uint32_t)(r.hi >> 32); }
__device__ __forceinline__ uint32_t rec_diag16(const Rec &r) { return (uint32_t)(((r.hi >> 16) & 0xFFFFu) - (r.hi & 0xFFFFu)) & 0xFFFFu; }

struct PrefArgs {
    const uint8_t *q;
    const uint32_t *qoff;
    uint32_t Q;
    const uint32_t *qlist;   // optional list of query indices (global-memory variant), else NULL
    const uint8_t *tres;
    const uint32_t *toff;
    uint32_t T;
    // k-mer table: 16-byte entries + posting lists, in nShards contiguous k-mer ranges of shardR k-mers (one range when not
    // sharded); ranges of other GPUs are peer-mapped pointers (index.cu: spb_index_shard)
    const ulonglong2 *directS[SPB_MAX_SHARDS];
    const uint64_t *postS[SPB_MAX_SHARDS];
    uint32_t shardR;
    int nShards;
    const uint32_t *row3;
    const uint16_t *cnt3;
    const Matrices *mat;
    Pattern6 pat;
    int kmerThr, maxHits, minDiag, compBias, minSubUngapped;
    int exactKmer;           // --exact-kmer-matching: a window looks up only its own k-mer (prefilter_kernel only)
    uint32_t binMask, idBase;
    Rec *gScratch;           // [gridDim.x][cap] (global-memory variant)
    uint32_t cap;            // record capacity per CTA
    spb_hit *arena;
    uint64_t *arenaT;        // optional: truncation key of every arena hit (target split, tkey_* below), else NULL
    unsigned long long *cursor;  // [0] arena cursor, [1] kmers, [2] dbMatches, [3] events, [4] kept hits
    uint64_t arenaCap;
    uint64_t *outStart;      // per query
    uint32_t *outCount;      // per query
    uint32_t *status;        // per query: ST_*
    uint32_t *need;          // per query: records needed when status == ST_CAPACITY
    // expand / lookup / bookkeeping kernels
    const uint32_t *kbits;   // [20^6 / 32] occupancy bit per k-mer
    uint32_t *recs;          // [chunkCap][256] slot records: k-mer | window << 26
    uint2 *chunkTab;         // [chunkCap] (query, stream-order number of the chunk's first slot)
    uint2 *chunkHits;        // [chunkCap] (first hit record of the chunk in hitArena, number of hits | PX_OVER)
    uint64_t *hitArena;      // [hitArenaCap] hit records, chunk after chunk: id << 32 | k << 16 | 16-bit diagonal
    uint32_t hitArenaCap;
    struct QMeta *meta;      // [Q]
    struct HeavyQ *heavyTab; // window tables of the heavy queries of the tile
    unsigned int *heavyCount;
    uint32_t *heavyItems;    // work items of expand_heavy_kernel: heavy query << 8 | chunk group
    unsigned int *heavyItemCount;
    uint32_t heavyItemCap;
    uint32_t heavyK, heavyCap;   // queries with more similar k-mers than heavyK are filled by expand_heavy_kernel
    unsigned int *chunkCursor;   // chunks allocated in this tile
    uint32_t chunkCap;
    uint64_t *hitScratch;    // [teams in the grid][2 * hitCap + 2 * FCAP] survivor list | grouped by target bucket | per-target winners
    uint32_t hitCap;
    unsigned int *qCounter;  // dynamic query scheduler
    const unsigned int *nListPtr;  // device: entries of qlist (team launch reads it, no host round trip); NULL: nList
    uint32_t nList, listOff, listStride;   // query of slot s: qlist[listOff + s * listStride]
    uint32_t hitsSolo;       // TEAM = 1: queries with more hits go to the TEAM = 8 launch
    uint32_t *teamList;      // heavy queries (classify_kernel) for the TEAM = 8 launch
    unsigned int *teamCount;
    uint32_t *teamList2;     // queries the TEAM = 1 launch hands over at its end (truncation at max-seqs, > 256 winners)
    unsigned int *teamCount2;
};

// table entry of a k-mer: x = 0 none | x, y with bit 63: one or two inline postings | x = len << 32 | offset of a posting list
template <bool SHARDED = true>
__device__ __forceinline__ ulonglong2 kmer_entry(const PrefArgs &A, uint32_t kmer, bool streaming) {
    uint32_t o = 0, l = kmer;
    if (SHARDED && A.nShards > 1) { o = kmer / A.shardR; l = kmer - o * A.shardR; }
    const ulonglong2 *d = (SHARDED ? A.directS[o] : A.directS[0]) + l;
    if (!streaming) return __ldg(d);
    // a random 16-byte read: with the default policy L2 fetches the whole 128-byte line (117 B of DRAM traffic per entry,
    // tools/gather_probe.cu); the 64-byte fetch hint halves that.  The access rate, which bounds the gather, is the same.
    ulonglong2 r;
    asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(d));
    return r;
}
// posting lists of the shard that owns the k-mer (list offsets in the entries are relative to it)
template <bool SHARDED = true>
__device__ __forceinline__ const uint64_t *kmer_post_base(const PrefArgs &A, uint32_t kmer) {
    return (SHARDED && A.nShards > 1) ? A.postS[kmer / A.shardR] : A.postS[0];
}
// number of postings of an entry and posting z of it
__device__ __forceinline__ uint32_t entry_count(const ulonglong2 &e) {
    return (e.x >> 63) ? 1u + (uint32_t)(e.y >> 63) : (uint32_t)(e.x >> 32);
}
__device__ __forceinline__ uint64_t entry_posting(const ulonglong2 &e, const uint64_t *postBase, uint32_t z) {
    if (e.x >> 63) return (z ? e.y : e.x) & 0x7FFFFFFFFFFFFFFFull;
    return __ldg(postBase + (uint32_t)e.x + z);
}

// Truncation key exported with a hit when the target shards of a multi-GPU run are merged (spb_prefilter_merge_shards).
// The reference cuts a list at max-seqs in the order of its stable counting sorts (QueryMatcher.cpp:131-164,
// CacheFriendlyOperations.cpp:311-341): capped score desc, id & (BINSIZE-1) asc, stream order of the winning event,
// target id.  All four are shard independent (global ids via id_base), so the union of the shards' lists can be cut
// exactly like the unsplit list.  Layout (ascending = earlier):
//   [63] 1 = counter not saturated | [62:55] 255 - capped score | [54:44] bin | [43:16] similar-k-mer number | [15:0] 0
// Saturated candidates (bit 63 = 0) all share the highest capped score, so the flag does not disturb the order; it lets the
// merge detect the rescoreHits corner (QueryMatcher.cpp:486-505: the cut falls among saturated counters).
__device__ __forceinline__ uint64_t tkey_make(uint32_t capped, bool saturated, uint32_t bin, uint32_t k) {
    return ((uint64_t)(saturated ? 0u : 1u) << 63) | ((uint64_t)(255u - capped) << 55) | ((uint64_t)bin << 44) |
           ((uint64_t)(k & 0x0FFFFFFFu) << 16);
}

#define ST_OK 0u
#define ST_CAPACITY 1u
#define ST_ARENA 2u
#define ST_LAYOUT 3u

#define PF_THREADS 128
#define PF_WARPS (PF_THREADS / 32)
#define PF_MAXL 512          // longest query whose per-position arrays fit the static shared memory
#define PF_BM_BITS 32768     // bits per "seen" bitmap
#define PF_BM_WORDS (PF_BM_BITS / 32)
#define PF_SMEM_CAP 1024u    // records per CTA in shared memory (16 KB)

__device__ __forceinline__ uint32_t id_hash(uint32_t id) { return (id * 2654435761u) >> (32 - 15); }

// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:37-51,368-382)
__device__ __forceinline__ int ungapped_diag(const uint8_t *__restrict__ t, uint32_t tLen, const uint8_t *q, int L,
                                             const signed char *corr, const signed char *m40u, int diagonal) {
    const uint32_t minDist = (uint32_t)abs(diagonal);
    uint32_t n;
    int qi;
    const uint8_t *tt;
    if (diagonal >= 0 && minDist < (uint32_t)L) {
        n = min(tLen, (uint32_t)L - minDist);
        qi = (int)minDist;
        tt = t;
    } else if (diagonal < 0 && minDist < tLen) {
        n = min(tLen - minDist, (uint32_t)L);
        qi = 0;
        tt = t + minDist;
    } else {
        return 0;
    }
    int best = 0, s = 0;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t a = __ldg(tt + k);
        int curr = (a < SPB_ALPHABET) ? (int)m40u[q[qi + k] * SPB_ALPHABET + a] + (int)corr[qi + k] : 0;
        s = max(0, s + curr);
        best = max(best, s);
    }
    return best;
}

// UngappedAlignment::scoreSingleSequence incl. computeLongScore (UngappedAlignment.cpp:216-233,395-403)
__device__ __forceinline__ int ungapped_u16(const uint8_t *__restrict__ tres, const uint32_t *__restrict__ toff, uint32_t id,
                                            const uint8_t *q, int L, const signed char *corr, const signed char *m40u,
                                            uint32_t diag16) {
    const uint32_t o = __ldg(toff + id);
    const uint32_t tLen = __ldg(toff + id + 1) - o;
    const uint8_t *t = tres + o;
    // one loop over the candidate real diagonals (a single inlined copy of the scorer): the 16-bit diagonal itself, or
    // for sequences >= 32768 every diagonal congruent to it (computeLongScore)
    const bool isLong = (L >= 32768 || tLen >= 32768);
    const uint32_t nNeg = isLong ? 1u + tLen / 32768u : 0u, nPos = isLong ? (uint32_t)L / 65536u + 1u : 0u;
    const uint32_t nIter = isLong ? nNeg + nPos : 1u;
    int totalMax = 0;
    for (uint32_t it = 0; it < nIter; it++) {
        int real;
        if (!isLong) real = (int)(short)diag16;
        else if (it < nNeg) real = (int)(-(long long)(it + 1u) * 65536 + (long long)diag16);
        else real = (int)((it - nNeg) * 65536u + diag16);
        totalMax = max(totalMax, ungapped_diag(t, tLen, q, L, corr, m40u, real));
    }
    return totalMax;
}


// In-place stable compaction of a[0..n) keeping elements with flag(i, a[i], a[i-1]); returns the new length.
template <class Flag>
__device__ __forceinline__ int block_compact_prev(Rec *a, int n, int *wsum, Rec *carry, Flag flag) {
    int outN = 0;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        Rec r = {0, 0}, rp = {~0ull, ~0ull};
        int keep = 0;
        if (i < n) {
            r = a[i];
            if (i > 0) rp = (threadIdx.x == 0) ? *carry : a[i - 1];
            keep = flag(i, r, rp) ? 1 : 0;
        }
        int total;
        int ex = block_excl_scan(keep, wsum, &total);  // barriers inside: all reads above precede all writes below
        const int last = min(n, c0 + (int)blockDim.x) - 1;
        if (i == last) *carry = r;
        if (keep) a[outN + ex] = r;
        outN += total;
        __syncthreads();
    }
    return outN;
}

// One work unit = 32 consecutive neighbours `a` of the first 3-mer of window w (KmerGenerator::calculateArrayProduct,
// KmerGenerator.cpp:186-216: outer loop over row(x1) while s1 >= thr - best2, inner loop over row(x2) while
// s2 >= thr - s1).  f(kmer, a, b, valid) is called warp-wide for every (a, b) pair of the unit.
template <class F>
__device__ __forceinline__ uint32_t warp_unit_kmers(const uint32_t *__restrict__ row3, const uint16_t *__restrict__ cnt3,
                                                    int x1, int x2, int thr, int n1, int a0, int *s_excl, uint32_t *s_i1, F f) {
    const int lane = threadIdx.x & 31;
    const uint32_t *r1 = row3 + (size_t)x1 * SPB_N3;
    const uint32_t *r2 = row3 + (size_t)x2 * SPB_N3;
    const uint16_t *c2 = cnt3 + (size_t)x2 * SPB_CNT3_W;
    const int a = a0 + lane;
    uint32_t e1 = 0;
    int n2 = 0;
    if (a < n1) {
        e1 = __ldg(r1 + a);
        const int s1 = (int)(e1 >> 16) - SPB_SCORE_OFF;
        int cut2 = thr - s1;
        if (cut2 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) {
            cut2 = max(cut2, -SPB_SCORE_OFF);
            n2 = (int)__ldg(c2 + cut2 + SPB_SCORE_OFF);
        }
    }
    const int inc = warp_incl_scan(n2);
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    __syncwarp();
    s_excl[lane] = inc - n2;
    s_i1[lane] = e1 & 0xFFFFu;
    __syncwarp();
    for (int p0 = 0; p0 < total; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < total;
        uint32_t kmer = 0;
        int j = 0, b = 0;
        if (valid) {
            int lo = 0, hi = 31;   // largest j with s_excl[j] <= p
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_excl[mid] <= p) lo = mid; else hi = mid - 1;
            }
            j = lo;
            b = p - s_excl[lo];
            kmer = s_i1[lo] + 8000u * (__ldg(r2 + b) & 0xFFFFu);
        }
        f(kmer, (uint32_t)(a0 + j), (uint32_t)b, valid);
    }
    __syncwarp();
    return (uint32_t)total;
}

template <bool GLOBAL>
__global__ void __launch_bounds__(PF_THREADS) prefilter_kernel(PrefArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_q[PF_MAXL];
    __shared__ float s_cb[PF_MAXL];
    __shared__ signed char s_corr[PF_MAXL];
    __shared__ short s_thr[PF_MAXL];            // per window: similar-k-mer threshold, -1 = skipped
    __shared__ unsigned short s_n1[PF_MAXL];    // per window: number of outer neighbours
    __shared__ int s_ubase[PF_MAXL + 1];        // per window: first work unit
    __shared__ signed char s_m40u[SPB_ALPHABET * SPB_ALPHABET];
    __shared__ uint32_t s_seen1[PF_BM_WORDS], s_seen2[PF_BM_WORDS];
    __shared__ int s_wsum[33];
    __shared__ int s_excl[PF_WARPS][32];
    __shared__ uint32_t s_i1[PF_WARPS][32];
    __shared__ int s_hist[256];
    __shared__ Rec s_carry;
    __shared__ unsigned int s_nHits;
    __shared__ int s_minCorr, s_thrD;
    __shared__ unsigned long long s_base, s_stat[2];
    __shared__ uint32_t s_kbase[PF_MAXL + 1];   // first similar-k-mer number of every window (truncation keys only)

    const uint32_t qi = A.qlist ? A.qlist[blockIdx.x] : blockIdx.x;
    if (qi >= A.Q) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t qo = A.qoff[qi];
    const int L = (int)(A.qoff[qi + 1] - qo);
    const int P = L - A.pat.span + 1;
    if (P < 2) {   // no window, or only the (never counted) last one
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_OK; }
        return;
    }
    if (L > PF_MAXL) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_LAYOUT; }
        return;
    }
    Rec *rec = GLOBAL ? (A.gScratch + (size_t)blockIdx.x * A.cap) : reinterpret_cast<Rec *>(smem_raw);
    const uint32_t CAP = A.cap;

    // ---------------- phase 0/1: query, composition bias, thresholds, work units ----------------
    for (int i = tid; i < SPB_ALPHABET * SPB_ALPHABET; i += PF_THREADS) s_m40u[i] = (signed char)A.mat->ungapped[i];
    for (int i = tid; i < L; i += PF_THREADS) s_q[i] = A.q[qo + i];
    for (int i = tid; i < PF_BM_WORDS; i += PF_THREADS) { s_seen1[i] = 0; s_seen2[i] = 0; }
    for (int i = tid; i < 256; i += PF_THREADS) s_hist[i] = 0;
    if (tid == 0) { s_nHits = 0; s_minCorr = 0; s_stat[0] = s_stat[1] = 0; }
    __syncthreads();
    // SubstitutionMatrix::calcLocalAaBiasCorrection (SubstitutionMatrix.cpp:85-114): float accumulator, the division
    // and every pBack*score term in double; explicit _rn intrinsics keep the compiler from fusing into FMA.
    for (int i = tid; i < L; i += PF_THREADS) {
        float d = 0.0f;
        if (A.compBias) {
            const int lo = max(0, i - 20), hi = min(L, i + 20);
            const int16_t *row = A.mat->seed + s_q[i] * SPB_ALPHABET;
            int sum = 0;
            for (int j = lo; j < hi; j++) sum += row[s_q[j]];
            sum -= row[s_q[i]];
            d = (float)sum;
            d = __double2float_rn(__ddiv_rn((double)d, -1.0 * (double)(float)(hi - lo)));
            for (int a = 0; a < SPB_ALPHABET; a++)
                d = __double2float_rn(__dadd_rn((double)d, __dmul_rn(A.mat->pback_seed[a], (double)(float)row[a])));
        }
        s_cb[i] = d;
        // UngappedAlignment::createProfile (UngappedAlignment.cpp:322-328)
        const float t4 = __fdiv_rn(d, 4.0f);
        const double dd = (d < 0.0f) ? __dsub_rn((double)t4, 0.5) : __dadd_rn((double)t4, 0.5);
        const signed char cc = (signed char)__double2float_rn(dd);
        s_corr[i] = cc;
        if (cc < 0) atomicMin(&s_minCorr, (int)cc);
    }
    __syncthreads();
    for (int w = tid; w < P; w += PF_THREADS) {
        bool hasX = false;
        float b = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            hasX = hasX || (s_q[w + A.pat.pos[k]] >= SPB_X);
            b = __fadd_rn(b, s_cb[w + A.pat.pos[k]]);
        }
        // QueryMatcher::match (QueryMatcher.cpp:205-219)
        const short bias = (short)((b < 0.0f) ? __dsub_rn((double)b, 0.5) : __dadd_rn((double)b, 0.5));
        const short thr = (short)max(A.kmerThr - (int)bias, 0);
        int n1 = 0;
        if (A.exactKmer) {
            // takeOnlyBestKmer (Prefiltering.cpp:177, QueryMatcher.cpp:221-227): one "outer neighbour", the 3-mer itself
            n1 = (!hasX && w < P - 1) ? 1 : 0;
        } else if (!hasX && w < P - 1) {   // quirk A: the last window is never counted
            const int x1 = s_q[w + A.pat.pos[0]] + 20 * s_q[w + A.pat.pos[1]] + 400 * s_q[w + A.pat.pos[2]];
            const int x2 = s_q[w + A.pat.pos[3]] + 20 * s_q[w + A.pat.pos[4]] + 400 * s_q[w + A.pat.pos[5]];
            const int best2 = (int)(__ldg(A.row3 + (size_t)x2 * SPB_N3) >> 16) - SPB_SCORE_OFF;
            const int cut1 = (int)thr - best2;
            if (cut1 <= SPB_CNT3_W - SPB_SCORE_OFF - 1)
                n1 = (int)__ldg(A.cnt3 + (size_t)x1 * SPB_CNT3_W + max(cut1, -SPB_SCORE_OFF) + SPB_SCORE_OFF);
        }
        s_thr[w] = thr;
        s_n1[w] = (unsigned short)n1;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int w = 0; w < P; w++) { s_ubase[w] = acc; acc += ((int)s_n1[w] + 31) >> 5; }
        s_ubase[P] = acc;
    }
    __syncthreads();
    const int nUnits = s_ubase[P];
    const int ungBias = abs(A.minSubUngapped) + abs(min(s_minCorr, 0));
    const int capScore = 255 - ungBias;
    // similar k-mers before (w, a): number of inner neighbours of every outer neighbour a' < a (KmerGenerator.cpp:186-216)
    auto inner_count = [&](int w, int x1, int x2, int a) -> uint32_t {
        if (A.exactKmer) return 1u;
        const int s1 = (int)(__ldg(A.row3 + (size_t)x1 * SPB_N3 + a) >> 16) - SPB_SCORE_OFF;
        const int cut2 = (int)s_thr[w] - s1;
        if (cut2 > SPB_CNT3_W - SPB_SCORE_OFF - 1) return 0u;
        return (uint32_t)__ldg(A.cnt3 + (size_t)x2 * SPB_CNT3_W + max(cut2, -SPB_SCORE_OFF) + SPB_SCORE_OFF);
    };
    if (A.arenaT) {
        for (int w = wid; w < P; w += PF_WARPS) {
            const int x1 = s_q[w + A.pat.pos[0]] + 20 * s_q[w + A.pat.pos[1]] + 400 * s_q[w + A.pat.pos[2]];
            const int x2 = s_q[w + A.pat.pos[3]] + 20 * s_q[w + A.pat.pos[4]] + 400 * s_q[w + A.pat.pos[5]];
            uint32_t acc = 0;
            for (int a = lane; a < (int)s_n1[w]; a += 32) acc += inner_count(w, x1, x2, a);
            for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            if (lane == 0) s_kbase[w + 1] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            s_kbase[0] = 0;
            for (int w = 0; w < P; w++) s_kbase[w + 1] += s_kbase[w];
        }
        __syncthreads();
    }

    // ---------------- phase 2: two passes over the hit stream ----------------
    unsigned long long nk = 0, nm = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int u = wid; u < nUnits; u += PF_WARPS) {
            int lo = 0, hi = P - 1;   // window of unit u: largest w with s_ubase[w] <= u
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_ubase[mid] <= u) lo = mid; else hi = mid - 1;
            }
            const int w = lo;
            const int a0 = (u - s_ubase[w]) << 5;
            const int x1 = s_q[w + A.pat.pos[0]] + 20 * s_q[w + A.pat.pos[1]] + 400 * s_q[w + A.pat.pos[2]];
            const int x2 = s_q[w + A.pat.pos[3]] + 20 * s_q[w + A.pat.pos[4]] + 400 * s_q[w + A.pat.pos[5]];
            auto on_kmer = [&](uint32_t kmer, uint32_t a, uint32_t b, bool valid) {
                    uint32_t len = 0;
                    ulonglong2 e = make_ulonglong2(0ull, 0ull);
                    const uint64_t *pb = nullptr;
                    if (valid && ((__ldg(A.kbits + (kmer >> 5)) >> (kmer & 31)) & 1u)) {
                        e = kmer_entry(A, kmer, false);
                        pb = kmer_post_base(A, kmer);
                        len = entry_count(e);
                    }
                    if (pass == 0) {
                        nm += len;
                        for (uint32_t z = 0; z < len; z++) {
                            const uint32_t h = id_hash((uint32_t)(entry_posting(e, pb, z) >> 16));
                            const uint32_t m = 1u << (h & 31);
                            const uint32_t old = atomicOr(&s_seen1[h >> 5], m);
                            if (old & m) atomicOr(&s_seen2[h >> 5], m);
                        }
                    } else {
                        for (uint32_t z = 0; z < len; z++) {
                            const uint64_t pe = entry_posting(e, pb, z);   // id << 16 | pos
                            const uint32_t id = (uint32_t)(pe >> 16);
                            const uint32_t h = id_hash(id);
                            const bool twice = (s_seen2[h >> 5] >> (h & 31)) & 1u;
                            const bool diag0 = (((uint32_t)w - (uint32_t)(pe & 0xFFFFu)) & 0xFFu) == 0u;
                            if (twice || diag0) {
                                const unsigned int dst = atomicAdd(&s_nHits, 1u);
                                if (dst < CAP) {
                                    Rec r;
                                    r.hi = ((uint64_t)id << 32) | ((uint64_t)w << 16) | (pe & 0xFFFFull);
                                    r.lo = ((uint64_t)a << 16) | (uint64_t)b;
                                    rec[dst] = r;
                                }
                            }
                        }
                    }
                };
            uint32_t nKmers = 1;
            if (A.exactKmer) on_kmer((uint32_t)x1 + 8000u * (uint32_t)x2, 0u, 0u, lane == 0);
            else nKmers = warp_unit_kmers(A.row3, A.cnt3, x1, x2, (int)s_thr[w], (int)s_n1[w], a0, s_excl[wid], s_i1[wid], on_kmer);
            if (pass == 0 && lane == 0) nk += nKmers;
        }
        __syncthreads();
    }
    for (int d = 16; d > 0; d >>= 1) nm += __shfl_down_sync(0xffffffffu, nm, d);
    if (lane == 0) { atomicAdd(&s_stat[0], nk); atomicAdd(&s_stat[1], nm); }
    __syncthreads();
    int n = (int)s_nHits;
    // later stages keep events A[0..EV) (Rec), final keys B[0..EV) (u64), cnt u8[EV] and ord u16/u32[EV] in the buffer
    const uint32_t EV = CAP / 2;
    // with truncation keys the upper half of the final-key array holds them: twice the room
    const uint32_t finNeed = (uint32_t)next_pow2(max(A.maxHits, 2)) * (A.arenaT ? 2u : 1u);
    if ((uint32_t)n > CAP || finNeed > EV) {
        if (tid == 0) {
            A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY;
            A.need[qi] = max((uint32_t)n, 2u * finNeed);
        }
        return;
    }

    // ---------------- phase 3: sort the kept hits by (id, window, rank) ----------------
    {
        const int np2 = next_pow2(max(n, 2));
        for (int i = n + tid; i < np2; i += PF_THREADS) { rec[i].hi = ~0ull; rec[i].lo = ~0ull; }
        __syncthreads();
        block_bitonic_sort(rec, np2, [](const Rec &a, const Rec &b) { return rec_less(a, b); });
    }

    // ---------------- phase 4: double-hit events, then de-duplication of consecutive equal events ----------------
    int ne = block_compact_prev(rec, n, s_wsum, &s_carry, [&](int i, const Rec &r, const Rec &rp) {
        const uint32_t d8 = rec_diag16(r) & 0xFFu;
        const bool same = (i > 0) && (rec_id(rp) == rec_id(r));
        return same ? (d8 == (rec_diag16(rp) & 0xFFu)) : (d8 == 0u);
    });
    int nkept = block_compact_prev(rec, ne, s_wsum, &s_carry, [&](int i, const Rec &r, const Rec &rp) {
        const bool same = (i > 0) && (rec_id(rp) == rec_id(r));
        return !same || ((rec_diag16(r) & 0xFFu) != (rec_diag16(rp) & 0xFFu));
    });
    if ((uint32_t)nkept > EV) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY; A.need[qi] = 2u * (uint32_t)nkept; }
        return;
    }
    Rec *evA = rec;
    uint64_t *finB = reinterpret_cast<uint64_t *>(rec + EV);
    uint8_t *cnt = reinterpret_cast<uint8_t *>(finB + EV);
    uint32_t *ord = reinterpret_cast<uint32_t *>(cnt + ((EV + 15) & ~15u));   // 16 + 8 + 1 + 4 bytes per slot <= 32

    // ---------------- phase 5: ungapped diagonal scores (capped like the uint8 counter) ----------------
    for (int i = tid; i < nkept; i += PF_THREADS) {
        const Rec r = evA[i];
        const int s = ungapped_u16(A.tres, A.toff, rec_id(r), s_q, L, s_corr, s_m40u, rec_diag16(r));
        cnt[i] = (uint8_t)min(s, capScore);
    }
    __syncthreads();

    // ---------------- phase 6: score >= minDiag and, per target, the first event reaching the target's maximum ----------------
    int nc = 0;
    for (int c0 = 0; c0 < nkept; c0 += PF_THREADS) {
        const int i = c0 + tid;
        int keep = 0;
        Rec r = {0, 0};
        uint8_t best = 0;
        if (i < nkept) {
            r = evA[i];
            const uint32_t id = rec_id(r);
            const bool head = (i == 0) || (rec_id(evA[i - 1]) != id);
            if (head) {
                int bi = i;
                best = cnt[i];
                for (int j = i + 1; j < nkept && rec_id(evA[j]) == id; j++)
                    if (cnt[j] > best) { best = cnt[j]; bi = j; }
                if ((int)best >= A.minDiag) { keep = 1; r = evA[bi]; }
            }
        }
        int total;
        const int ex = block_excl_scan(keep, s_wsum, &total);
        if (keep) { evA[nc + ex] = r; cnt[nc + ex] = best; }
        nc += total;
        __syncthreads();
    }

    // ---------------- phase 7: top-maxHits selection (QueryMatcher.cpp:131-164) ----------------
    const int maxHits = A.maxHits;
    int ns = nc;
    bool rescale = false;
    int maxSelf = 1;
    if (nc >= maxHits) {
        // computeScoreThreshold (QueryMatcher.h:199-209)
        for (int i = tid; i < nc; i += PF_THREADS) atomicAdd(&s_hist[cnt[i]], 1);
        __syncthreads();
        if (tid == 0) {
            int found = 0, thr = 0;
            for (int s = 255; s >= 1; s--) {
                found += s_hist[s];
                if (found >= maxHits) { thr = s; break; }
            }
            s_thrD = max(thr, A.minDiag);
        }
        __syncthreads();
        rescale = (s_thrD >= capScore);
    }
    if (rescale) {
        // rescoreHits (QueryMatcher.cpp:486-505): every remaining candidate is saturated; rescale exact scores to 0..255
        {
            int s = 0, best = 0;
            for (int k = 0; k < L; k++) {
                s = max(0, s + (int)s_m40u[s_q[k] * SPB_ALPHABET + s_q[k]] + (int)s_corr[k]);
                best = max(best, s);
            }
            maxSelf = min(max(best - capScore, 1), 65535);
        }
        const float fltMax = (float)maxSelf;
        int m = 0;
        for (int c0 = 0; c0 < nc; c0 += PF_THREADS) {
            const int i = c0 + tid;
            int keep = 0;
            Rec r = {0, 0};
            uint8_t nc8 = 0;
            if (i < nc && (int)cnt[i] >= capScore) {
                r = evA[i];
                unsigned int nsx = (unsigned int)ungapped_u16(A.tres, A.toff, rec_id(r), s_q, L, s_corr, s_m40u, rec_diag16(r));
                nsx -= (unsigned int)capScore;
                const float sc = (float)min(nsx, 65535u);
                nc8 = (uint8_t)__dadd_rn((double)__fmul_rn(__fdiv_rn(sc, fltMax), 255.0f), 0.5);
                keep = 1;
            }
            int total;
            const int ex = block_excl_scan(keep, s_wsum, &total);
            if (keep) { evA[m + ex] = r; cnt[m + ex] = nc8; }
            m += total;
            __syncthreads();
        }
        nc = m;
    }
    if (nc > maxHits) {
        // order of the reference's stable counting sorts: score desc, bin, (window, rank), id
        const uint32_t binMask = A.binMask, idBase = A.idBase;
        const int np2 = next_pow2(nc);
        for (int i = tid; i < np2; i += PF_THREADS) ord[i] = i < nc ? (uint32_t)i : 0xFFFFFFFFu;
        __syncthreads();
        block_bitonic_sort(ord, np2, [&](uint32_t a, uint32_t b) {
            if (a == 0xFFFFFFFFu || b == 0xFFFFFFFFu) return a != 0xFFFFFFFFu && b == 0xFFFFFFFFu;
            const uint8_t ca = cnt[a], cb = cnt[b];
            if (ca != cb) return ca > cb;
            const Rec ra = evA[a], rb = evA[b];
            const uint32_t ba = (rec_id(ra) + idBase) & binMask, bb = (rec_id(rb) + idBase) & binMask;
            if (ba != bb) return ba < bb;
            const uint32_t wa = (uint32_t)(ra.hi >> 16) & 0xFFFFu, wb = (uint32_t)(rb.hi >> 16) & 0xFFFFu;
            if (wa != wb) return wa < wb;
            if (ra.lo != rb.lo) return ra.lo < rb.lo;
            return rec_id(ra) < rec_id(rb);
        });
        ns = maxHits;
    } else {
        ns = nc;
    }

    // ---------------- phase 8: exact scores, final order (score desc, id asc), output ----------------
    for (int k = tid; k < ns; k += PF_THREADS) {
        const int i = (nc > maxHits) ? (int)ord[k] : k;
        const Rec r = evA[i];
        const uint32_t id = rec_id(r);
        int sc = cnt[i];
        if (rescale) {
            sc = capScore + (int)((unsigned int)cnt[i] * (unsigned int)maxSelf / 255u);
        } else if (sc >= capScore) {
            sc = ungapped_u16(A.tres, A.toff, id, s_q, L, s_corr, s_m40u, rec_diag16(r));   // getResult, QueryMatcher.cpp:371-374
        }
        // (score desc, id asc): 16-bit inverted score | 32-bit id | 16-bit diagonal (scores < 64 * PF_MAXL fit 16 bits)
        if (!A.arenaT) {
            finB[k] = ((uint64_t)(0xFFFFu - (uint32_t)min(sc, 0xFFFF)) << 48) | ((uint64_t)id << 16) | rec_diag16(r);
        } else {
            // truncation key (its low 16 bits carry the diagonal until the output) in the upper half of finB; the sort key
            // carries the slot instead of the diagonal.  After rescoreHits the counters are no longer comparable across
            // shards: key 0 (saturated) makes the merge report the corner.
            uint64_t tk = 0;
            if (!rescale) {
                const int w = (int)((r.hi >> 16) & 0xFFFFu);
                const int a = (int)(r.lo >> 16);
                const int x1 = s_q[w + A.pat.pos[0]] + 20 * s_q[w + A.pat.pos[1]] + 400 * s_q[w + A.pat.pos[2]];
                const int x2 = s_q[w + A.pat.pos[3]] + 20 * s_q[w + A.pat.pos[4]] + 400 * s_q[w + A.pat.pos[5]];
                uint32_t kn = s_kbase[w] + (uint32_t)(r.lo & 0xFFFFu);
                for (int aa = 0; aa < a; aa++) kn += inner_count(w, x1, x2, aa);
                tk = tkey_make(cnt[i], (int)cnt[i] >= capScore, (id + A.idBase) & A.binMask, kn);
            }
            finB[EV / 2 + k] = tk | rec_diag16(r);
            finB[k] = ((uint64_t)(0xFFFFu - (uint32_t)min(sc, 0xFFFF)) << 48) | ((uint64_t)id << 16) | (uint64_t)k;
        }
    }
    {
        const int np2 = next_pow2(max(ns, 2));
        for (int i = ns + tid; i < np2; i += PF_THREADS) finB[i] = ~0ull;
        __syncthreads();
        block_bitonic_sort(finB, np2, [](uint64_t a, uint64_t b) { return a < b; });
    }
    if (tid == 0) {
        s_base = atomicAdd(A.cursor, (unsigned long long)ns);
        atomicAdd(A.cursor + 1, s_stat[0]);
        atomicAdd(A.cursor + 2, s_stat[1]);
        atomicAdd(A.cursor + 3, (unsigned long long)nkept);
        atomicAdd(A.cursor + 4, (unsigned long long)n);
    }
    __syncthreads();
    const unsigned long long base = s_base;
    if (base + (unsigned long long)ns > A.arenaCap) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_ARENA; }
        return;
    }
    for (int k = tid; k < ns; k += PF_THREADS) {
        const uint64_t f = finB[k];
        spb_hit h;
        h.id = (uint32_t)(f >> 16);
        h.score = (int32_t)(0xFFFFu - (uint32_t)(f >> 48));
        uint32_t dg = (uint32_t)(f & 0xFFFFull);
        if (A.arenaT) {
            const uint64_t tk = finB[EV / 2 + dg];
            dg = (uint32_t)(tk & 0xFFFFull);
            A.arenaT[base + k] = tk & ~0xFFFFull;
        }
        h.diag = (int32_t)(short)(uint16_t)dg;
        A.arena[base + k] = h;
    }
    if (tid == 0) { A.outStart[qi] = base; A.outCount[qi] = (uint32_t)ns; A.status[qi] = ST_OK; }
}

// =========================================================================================================
// The production path: three throughput-shaped kernels per tile of queries ("expand -> gather -> bookkeeping").
//
//  expand_kernel    one warp per query: composition bias, window thresholds, and the similar k-mers of the query written
//                   as flat 4-byte SLOT records (k-mer | window << 26) in the reference's enumeration order (window, outer
//                   neighbour a, inner neighbour b: KmerGenerator.cpp:107-216).  Outer neighbours of equal score share their
//                   inner count, so the list of a window is a handful of dense rectangles (rows x inner count) whose sizes
//                   come straight out of the score-count table cnt3 -- no 3-mer row is scanned to find them.  A query owns a
//                   run of CHUNKS of 256 slots (the tail padded with an invalid k-mer) allocated with one atomic.
//  lookup_kernel    the gather, one thread per slot over all chunks of the tile -- every query, light or heavy, is the same
//                   uniform work, so there is no tail and no team launch for long k-mer lists.  A slot asks the 8 MB occupancy
//                   bitmap (L2-resident) first; only k-mers some target contains read their 16-byte table entry (streaming
//                   load, does not displace the bitmap / hit records in L2).  Hits (id << 32 | k << 16 | 16-bit diagonal, k =
//                   slot number inside the query = stream order) are compacted per chunk into the chunk's own region.
//  bookkeep_kernel  the order-dependent rules per query (TEAM = 1: a warp; TEAM = 8: a CTA for queries with many hits): pass
//                   over the query's hit chunks marking targets seen once / twice in shared-memory bitmaps, second pass keeping
//                   the records of targets seen twice (or with 8-bit diagonal 0, quirk B), grouped sort by (target, stream
//                   order), double-hit rule, de-duplication, ungapped score, per-target winner, final order, output.
// Anything that does not fit (query longer than 64, >= 65536 k-mers, chunk region / group / candidate overflow, the
// all-saturated rescoring corner) is left to prefilter_kernel via status = ST_CAPACITY; a query that found the slot arena of
// its tile full is run again in a later tile (ST_RETRY).  Results are identical on every path.
// =========================================================================================================
#define PW_WARPS 4           // independent warps per CTA in the TEAM = 1 launch
#define PW_TEAM 8            // warps per query in the team launch
#define PW_MAXL 64
#define PW_SCAP 256          // records a warp sorts at once
#define PW_FCAP_SOLO 512     // per-target winners of one query (slots behind the grouped records)
#define PW_FCAP_TEAM 4096
#define PW_BIG 1024          // TEAM = 8: records the CTA sorts at once in the final stage
#define PX_CH 256            // slots per chunk = threads per CTA of the lookup kernel
#define PX_OVER 0xFFFFFFFFu   // chunkHits[c].y: the hit arena was full
#define PX_INVALID 0xFFFFFFFFu
#define ST_TEAM 4u
#define ST_RETRY 5u          // slot arena of the tile was full: the query runs in a later tile
#define ST_PENDING 6u        // expanded, waiting for its bookkeeping

struct QMeta {               // per query, expand -> bookkeeping
    uint32_t chunk0, nChunks, K;
    uint32_t nHits;          // classify_kernel: hits of the query = sum of its chunk counts
    int32_t minCorr;
    uint32_t heavy;          // 1 + index of the query's window table (HeavyQ) when its slot records are filled chunk by chunk, else 0
    signed char corr[PW_MAXL];
};

// Window table of a query with very many similar k-mers (K is heavy-tailed: median 340, mean 1076, max 38 161 per frame): its
// slot records are not written by the warp that counted them -- that one warp would still be filling long after every
// other query of the tile is done -- but chunk by chunk by expand_heavy_kernel, one warp per 256-slot chunk.
struct HeavyQ {
    uint32_t kstart[PW_MAXL];            // slots before window w (kstart[nWin] = K)
    uint16_t x1[PW_MAXL], x2[PW_MAXL];
    short thr[PW_MAXL], best1[PW_MAXL], cut1[PW_MAXL];
    uint32_t nWin, chunk0, K;
};
#define PX_HEAVY_GROUP 8u     // chunks per work item of expand_heavy_kernel

struct ExpSmem {
    uint32_t kst[PW_MAXL + 1];            // slots before every window
    float cb[PW_MAXL];
    uint16_t x1[PW_MAXL], x2[PW_MAXL];   // first / second 3-mer of every window
    short thr[PW_MAXL];
    short best1[PW_MAXL], cut1[PW_MAXL]; // best outer score / lowest outer score of the window (best1 = -32768: no k-mers)
    uint8_t q[PW_MAXL];
    uint32_t v_qi;
};

__device__ __noinline__ int warp_incl_scan_ool(int v) { return warp_incl_scan(v); }

// One window of a query: the rectangles of its similar-k-mer list, 32 outer score values per round (lane = score value,
// descending = ascending outer rank).  EXPAND = false: returns the number of k-mers.  EXPAND = true: writes the slot records
// from `dst` on (dst[k] for the k-th k-mer of the window) and returns the same count.
// The records of a round are filled FLAT: the lanes that own a rectangle leave its descriptor (first slot, first outer rank,
// inner count, reciprocal) in the warp's shared-memory table, then lane l takes slots l, l + 32, ... of the whole round and
// walks the table forward to the rectangle its slot lies in -- every lane is busy whatever the rectangle sizes (they average
// 50 slots, a quarter of them fewer than 12), and the two slots of a trip keep four independent row loads in flight.
struct ExpRound {
    uint4 desc[32];          // (first slot, first outer rank, inner count, magic reciprocal) per rectangle of the round
    uint32_t end[32];        // one past its last slot
};

template <bool EXPAND>
__device__ __forceinline__ uint32_t window_kmers(const uint32_t *__restrict__ row3, const uint16_t *__restrict__ cnt3, int x1, int x2,
                                                 int thr, int best1, int cut1, uint32_t w, uint32_t *dst, ExpRound *R,
                                                 uint32_t lo = 0u, uint32_t hi = 0xFFFFFFFFu) {     // only slots [lo, hi) of the window are written
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int vlo = max(cut1, -SPB_SCORE_OFF);
    const uint16_t *c1 = cnt3 + (size_t)x1 * SPB_CNT3_W;
    const uint16_t *c2 = cnt3 + (size_t)x2 * SPB_CNT3_W;
    const uint32_t *row1 = row3 + (size_t)x1 * SPB_N3;
    const uint32_t *row2 = row3 + (size_t)x2 * SPB_N3;
    uint32_t done = 0;
    for (int v0 = best1; v0 >= vlo; v0 -= 32) {
        const int v = v0 - lane;
        uint32_t rows = 0, c = 0, a0 = 0;
        if (v >= vlo) {
            const int idx = v + SPB_SCORE_OFF;                    // 0 .. 511
            const uint32_t ge = __ldg(c1 + idx);                   // outer neighbours with score >= v
            a0 = (idx + 1 < SPB_CNT3_W) ? (uint32_t)__ldg(c1 + idx + 1) : 0u;   // ... with score > v: rank of the first of score v
            rows = ge - a0;
            const int cut2 = thr - v;
            if (rows && cut2 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) c = __ldg(c2 + max(cut2, -SPB_SCORE_OFF) + SPB_SCORE_OFF);
        }
        const uint32_t cnt = rows * c;
        const uint32_t inc = (uint32_t)warp_incl_scan_ool((int)cnt);
        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
        if (EXPAND && tot && done < hi && done + tot > lo) {
            const unsigned m = __ballot_sync(0xffffffffu, cnt != 0);
            if (cnt) {
                const int r = __popc(m & lt);
                // magic reciprocal of the inner count: p / c == umulhi(p, M) for p < 2^16, 1 < c < 2^16 (M = floor(2^32 / c) + 1)
                R->desc[r] = make_uint4(done + inc - cnt, a0, c, (c > 1) ? (0xFFFFFFFFu / c + 1u) : 0u);
                R->end[r] = done + inc;
            }
            __syncwarp();
            const uint32_t end = min(done + tot, hi);
            int r0 = 0, r1 = 0;
            for (uint32_t s0 = max(done, lo) + (uint32_t)lane; s0 < end; s0 += 64) {
                const uint32_t s1 = s0 + 32;
                const bool two = s1 < end;
                while (s0 >= R->end[r0]) r0++;
                r1 = max(r1, r0);
                if (two) { while (s1 >= R->end[r1]) r1++; }
                const uint4 d0 = R->desc[r0], d1 = R->desc[r1];
                const uint32_t p0 = s0 - d0.x, p1 = s1 - d1.x;
                const uint32_t ar0 = (d0.z > 1) ? __umulhi(p0, d0.w) : p0, ar1 = (d1.z > 1) ? __umulhi(p1, d1.w) : p1;
                const uint32_t b0 = p0 - ar0 * d0.z, b1 = p1 - ar1 * d1.z;
                uint32_t i1a = __ldg(row1 + d0.y + ar0), i2a = __ldg(row2 + b0), i1b = 0, i2b = 0;
                if (two) { i1b = __ldg(row1 + d1.y + ar1); i2b = __ldg(row2 + b1); }
                dst[s0] = ((i1a & 0xFFFFu) + 8000u * (i2a & 0xFFFFu)) | (w << 26);
                if (two) dst[s1] = ((i1b & 0xFFFFu) + 8000u * (i2b & 0xFFFFu)) | (w << 26);
                r0 = r1;
            }
            __syncwarp();      // the next round rewrites the table
        }
        done += tot;
    }
    return done;
}

__global__ void __launch_bounds__(PW_WARPS * 32, 10) expand_kernel(PrefArgs A) {
    __shared__ ExpSmem s_e[PW_WARPS];
    __shared__ ExpRound s_r[PW_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ExpSmem &S = s_e[wid];
    for (;;) {
        if (lane == 0) S.v_qi = atomicAdd(A.qCounter, 1u);
        __syncwarp();
        const uint32_t slot = S.v_qi;
        __syncwarp();
        if (slot >= A.nList) break;
        const uint32_t qi = A.qlist[A.listOff + (size_t)slot * A.listStride];
        const uint32_t qo = A.qoff[qi];
        const int L = (int)(A.qoff[qi + 1] - qo);
        const int P = L - A.pat.span + 1;
        if (P < 2) {   // no window, or only the (never counted) last one
            if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_OK; }
            continue;
        }
        if (L > PW_MAXL) {
            if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY; A.need[qi] = 1; }
            continue;
        }
        QMeta &M = A.meta[qi];
        // query, composition bias (SubstitutionMatrix.cpp:85-114: float accumulator, the division and every pBack * score
        // term in double; explicit _rn intrinsics keep the compiler from fusing into FMA), ungapped-profile correction
        for (int i = lane; i < L; i += 32) S.q[i] = A.q[qo + i];
        __syncwarp();
        int minCorr = 0;
        for (int i = lane; i < L; i += 32) {
            float d = 0.0f;
            if (A.compBias) {
                const int lo = max(0, i - 20), hi = min(L, i + 20);
                const int16_t *row = A.mat->seed + S.q[i] * SPB_ALPHABET;
                int sum = 0;
                for (int j = lo; j < hi; j++) sum += row[S.q[j]];
                sum -= row[S.q[i]];
                d = (float)sum;
                d = __double2float_rn(__ddiv_rn((double)d, -1.0 * (double)(float)(hi - lo)));
                for (int a = 0; a < SPB_ALPHABET; a++)
                    d = __double2float_rn(__dadd_rn((double)d, __dmul_rn(A.mat->pback_seed[a], (double)(float)row[a])));
            }
            S.cb[i] = d;
            // UngappedAlignment::createProfile (UngappedAlignment.cpp:322-328)
            const float t4 = __fdiv_rn(d, 4.0f);
            const double dd = (d < 0.0f) ? __dsub_rn((double)t4, 0.5) : __dadd_rn((double)t4, 0.5);
            const signed char cc = (signed char)__double2float_rn(dd);
            M.corr[i] = cc;
            minCorr = min(minCorr, (int)cc);
        }
        for (int d = 16; d > 0; d >>= 1) minCorr = min(minCorr, __shfl_xor_sync(0xffffffffu, minCorr, d));
        __syncwarp();
        for (int w = lane; w < P; w += 32) {
            bool hasX = false;
            float b = 0.0f;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                hasX = hasX || (S.q[w + A.pat.pos[k]] >= SPB_X);
                b = __fadd_rn(b, S.cb[w + A.pat.pos[k]]);
            }
            // QueryMatcher::match (QueryMatcher.cpp:205-219)
            const short bias = (short)((b < 0.0f) ? __dsub_rn((double)b, 0.5) : __dadd_rn((double)b, 0.5));
            const short thr = (short)max(A.kmerThr - (int)bias, 0);
            int x1 = 0, x2 = 0, best1 = -32768, cut1 = 0;
            if (!hasX && w < P - 1) {   // quirk A: the last window is never counted
                x1 = S.q[w + A.pat.pos[0]] + 20 * S.q[w + A.pat.pos[1]] + 400 * S.q[w + A.pat.pos[2]];
                x2 = S.q[w + A.pat.pos[3]] + 20 * S.q[w + A.pat.pos[4]] + 400 * S.q[w + A.pat.pos[5]];
                const int best2 = (int)(__ldg(A.row3 + (size_t)x2 * SPB_N3) >> 16) - SPB_SCORE_OFF;
                cut1 = (int)thr - best2;
                if (cut1 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) best1 = (int)(__ldg(A.row3 + (size_t)x1 * SPB_N3) >> 16) - SPB_SCORE_OFF;
            }
            S.thr[w] = thr;
            S.x1[w] = (unsigned short)x1;
            S.x2[w] = (unsigned short)x2;
            S.best1[w] = (short)best1;
            S.cut1[w] = (short)cut1;
        }
        __syncwarp();
        // similar k-mers of the query: count, allocate chunks, write
        uint32_t K = 0;
        for (int w = 0; w < P - 1; w++) {
            if (lane == 0) S.kst[w] = K;
            if (S.best1[w] != -32768)
                K += window_kmers<false>(A.row3, A.cnt3, S.x1[w], S.x2[w], S.thr[w], S.best1[w], S.cut1[w], (uint32_t)w, nullptr, nullptr);
        }
        if (lane == 0) S.kst[P - 1] = K;
        __syncwarp();
        if (K > 65535u) {
            if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY; A.need[qi] = 4; }
            continue;
        }
        const uint32_t nCh = (K + PX_CH - 1) / PX_CH;
        uint32_t base = 0;
        if (lane == 0 && nCh) base = atomicAdd(A.chunkCursor, nCh);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + nCh > A.chunkCap) {
            // the chunks of the claim that lie inside the arena are visited by the lookup kernel: make them empty
            for (uint32_t ch = base; ch < min(base + nCh, A.chunkCap); ch++) {
                for (int i = lane; i < PX_CH; i += 32) A.recs[(size_t)ch * PX_CH + i] = PX_INVALID;
                if (lane == 0) A.chunkTab[ch] = make_uint2(qi, 0u);
            }
            if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_RETRY; }
            continue;
        }
        uint32_t *dst = A.recs + (size_t)base * PX_CH;
        uint32_t heavy = 0;
        if (K > A.heavyK) {
            if (lane == 0) heavy = atomicAdd(A.heavyCount, 1u) + 1u;
            heavy = __shfl_sync(0xffffffffu, heavy, 0);
            if (heavy > A.heavyCap) heavy = 0;        // table full: filled here after all
        }
        if (heavy) {
            HeavyQ &H = A.heavyTab[heavy - 1];
            for (int w = lane; w < P; w += 32) {
                H.kstart[w] = S.kst[w];
                if (w < P - 1) { H.x1[w] = S.x1[w]; H.x2[w] = S.x2[w]; H.thr[w] = S.thr[w]; H.best1[w] = S.best1[w]; H.cut1[w] = S.cut1[w]; }
            }
            const uint32_t nG = (nCh + PX_HEAVY_GROUP - 1) / PX_HEAVY_GROUP;
            uint32_t ib = 0;
            if (lane == 0) {
                H.nWin = (uint32_t)(P - 1); H.chunk0 = base; H.K = K;
                ib = atomicAdd(A.heavyItemCount, nG);
            }
            ib = __shfl_sync(0xffffffffu, ib, 0);
            for (uint32_t g = (uint32_t)lane; g < nG; g += 32) A.heavyItems[ib + g] = ((heavy - 1) << 8) | g;
        } else {
            for (int w = 0; w < P - 1; w++)
                if (S.best1[w] != -32768)
                    window_kmers<true>(A.row3, A.cnt3, S.x1[w], S.x2[w], S.thr[w], S.best1[w], S.cut1[w], (uint32_t)w, dst + S.kst[w], &s_r[wid]);
        }
        for (uint32_t k = K + (uint32_t)lane; k < nCh * PX_CH; k += 32) dst[k] = PX_INVALID;
        for (uint32_t i = (uint32_t)lane; i < nCh; i += 32) A.chunkTab[base + i] = make_uint2(qi, i * PX_CH);
        if (lane == 0) { M.chunk0 = base; M.nChunks = nCh; M.K = K; M.minCorr = minCorr; M.heavy = heavy; A.status[qi] = ST_PENDING; }
    }
}

// Slot records of the heavy queries: one warp per group of 8 chunks (2048 slots; the work items were listed by expand_kernel).
// The warp finds the windows that overlap its stretch in the query's window table and fills only its own slots of them --
// every work item is at most 32 trips, so the queries with tens of thousands of similar k-mers no longer form the tail of the
// stage.
__global__ void __launch_bounds__(PW_WARPS * 32, 10) expand_heavy_kernel(PrefArgs A) {
    __shared__ ExpRound s_r[PW_WARPS];
    const int wid = threadIdx.x >> 5;
    const uint32_t nItems = min(*A.heavyItemCount, A.heavyItemCap);
    const uint32_t nWarps = gridDim.x * PW_WARPS;
    for (uint32_t it = blockIdx.x * PW_WARPS + (uint32_t)wid; it < nItems; it += nWarps) {
        const uint32_t item = A.heavyItems[it];
        const HeavyQ &H = A.heavyTab[item >> 8];
        const uint32_t K = H.K;
        const uint32_t lo = (item & 0xFFu) * PX_HEAVY_GROUP * PX_CH, hi = min(lo + PX_HEAVY_GROUP * PX_CH, K);
        uint32_t *dst = A.recs + (size_t)H.chunk0 * PX_CH;
        const int nW = (int)H.nWin;
        for (int w = 0; w < nW; w++) {
            const uint32_t ks = H.kstart[w], ke = H.kstart[w + 1];
            if (ke <= lo || ke == ks) continue;
            if (ks >= hi) break;
            window_kmers<true>(A.row3, A.cnt3, H.x1[w], H.x2[w], H.thr[w], H.best1[w], H.cut1[w], (uint32_t)w, dst + ks, &s_r[wid],
                               lo > ks ? lo - ks : 0u, hi - ks);
        }
    }
}

// The gather.  A CTA takes a PAIR of chunks (512 slots, two per thread) per trip, grid-stride over the chunks the tile
// allocated (the count is read from device memory: no host round trip after expand_kernel).  Per slot: occupancy bit (L2) ->
// 16-byte table entry (DRAM, streaming load) -> 0, 1 or 2 inline postings or a posting list; the slot records of the next
// pair are already in flight.  The hits of the pair are compacted in slot order (two CTA-wide scans sharing ONE barrier) and
// appended to the CTA's private stretch of the tile's hit arena -- every thread advances the same private cursor, so there is
// no atomic and no second barrier on the path; chunkHits[c] = (first record, count).  k-mer frequencies are far from uniform
// (a 6-mer of common residues occurs dozens of times: 5 % of the looked-up k-mers hold 38 % of the hits in lists of three or
// more), so a chunk holds anything from 0 to thousands of hits.  A pair that finds its stretch full stores count = PX_OVER
// and its queries run again in the next round.
__device__ __forceinline__ uint64_t hit_record(uint64_t pe, uint64_t kw, uint32_t w) {
    return ((pe >> 16) << 32) | kw | (uint64_t)((w - (uint32_t)(pe & 0xFFFFull)) & 0xFFFFu);
}

template <bool SHARDED>
__global__ void __launch_bounds__(PX_CH, 6) lookup_kernel(PrefArgs A) {
    __shared__ uint32_t s_w[2][2][PX_CH / 32];     // [trip parity][chunk of the pair][warp]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nCh = min(*A.chunkCursor, A.chunkCap);
    uint64_t *__restrict__ arena = A.hitArena;
    const uint32_t nPairs = (nCh + 1) >> 1;
    // private stretch of the hit arena
    const uint32_t stretch = A.hitArenaCap / gridDim.x;
    uint32_t cur = blockIdx.x * stretch;
    const uint32_t curEnd = cur + stretch;
    uint32_t rec[2] = {PX_INVALID, PX_INVALID};
    if (blockIdx.x < nPairs) {
#pragma unroll
        for (int u = 0; u < 2; u++)
            if (2 * blockIdx.x + u < nCh) rec[u] = __ldcs(A.recs + (size_t)(2 * blockIdx.x + u) * PX_CH + tid);
    }
    int par = 0;
    for (uint32_t pr = blockIdx.x; pr < nPairs; pr += gridDim.x, par ^= 1) {
        const uint32_t c0 = 2 * pr;
        uint32_t kmer[2], w[2], n[2];
        ulonglong2 e[2];
        bool occ[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            kmer[u] = rec[u] & 0x03FFFFFFu;
            w[u] = rec[u] >> 26;
            occ[u] = kmer[u] < SPB_NK && ((__ldg(A.kbits + (kmer[u] >> 5)) >> (kmer[u] & 31)) & 1u);
        }
        // slot records of the next pair: in flight while this one waits for its table entries
        const uint32_t prN = pr + gridDim.x;
#pragma unroll
        for (int u = 0; u < 2; u++)
            rec[u] = (prN < nPairs && 2 * prN + u < nCh) ? __ldcs(A.recs + (size_t)(2 * prN + u) * PX_CH + tid) : PX_INVALID;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            e[u] = make_ulonglong2(0ull, 0ull);
            if (occ[u]) e[u] = kmer_entry<SHARDED>(A, kmer[u], true);
        }
        uint32_t inc[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            n[u] = entry_count(e[u]);
            inc[u] = (uint32_t)warp_incl_scan((int)n[u]);
            if (lane == 31) s_w[par][u][wid] = inc[u];
        }
        __syncthreads();     // the only barrier of the trip (the warp totals are double-buffered by trip parity)
        // prefix over the warps of the CTA: lanes 0..7 hold the warp totals, two warp reductions give (before me, all)
        uint32_t off[2], total[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t v = (lane < PX_CH / 32) ? s_w[par][u][lane] : 0u;
            total[u] = __reduce_add_sync(0xffffffffu, v);
            off[u] = __reduce_add_sync(0xffffffffu, lane < wid ? v : 0u) + inc[u] - n[u];
        }
        const uint32_t tt = total[0] + total[1];
        const bool fits = tt <= curEnd - cur;
        if (tid == 0) {
            A.chunkHits[c0] = make_uint2(cur, fits ? total[0] : PX_OVER);
            if (c0 + 1 < nCh) A.chunkHits[c0 + 1] = make_uint2(cur + total[0], fits ? total[1] : PX_OVER);
        }
        if (!fits) continue;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (n[u]) {
                uint64_t *__restrict__ dst = arena + (size_t)(cur + (u ? total[0] : 0u) + off[u]);
                const uint64_t kw = (uint64_t)(A.chunkTab[c0 + u].y + (uint32_t)tid) << 16;
                if (e[u].x >> 63) {
                    dst[0] = hit_record(e[u].x, kw, w[u]);
                    if (e[u].y >> 63) dst[1] = hit_record(e[u].y, kw, w[u]);
                } else {
                    // list of three or more postings: four independent loads in flight per trip
                    const uint64_t *__restrict__ pl = kmer_post_base<SHARDED>(A, kmer[u]) + (uint32_t)e[u].x;
                    for (uint32_t z = 0; z < n[u]; z += 4) {
                        uint64_t pe[4];
#pragma unroll
                        for (int v = 0; v < 4; v++) pe[v] = (z + v < n[u]) ? __ldg(pl + z + v) : 0ull;
#pragma unroll
                        for (int v = 0; v < 4; v++)
                            if (z + v < n[u]) dst[z + v] = hit_record(pe[v], kw, w[u]);
                    }
                }
            }
        }
        cur += tt;
    }
}

// After the gather: hits per query (sum of its chunk counts) and the split of the tile into the queries a single warp finishes
// (stay ST_PENDING) and the heavy ones a CTA takes (listed, ST_TEAM) -- the two bookkeeping launches then run side by side on
// two streams.  A chunk that found the hit arena full sends its query to the next round (ST_RETRY).
__global__ void classify_kernel(PrefArgs A) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.nList) return;
    const uint32_t qi = A.qlist[A.listOff + (size_t)s * A.listStride];
    if (A.status[qi] != ST_PENDING) return;
    QMeta &M = A.meta[qi];
    uint32_t sum = 0;
    bool over = false;
    for (uint32_t ch = 0; ch < M.nChunks; ch++) {
        const uint32_t n = A.chunkHits[M.chunk0 + ch].y;
        over = over || (n == PX_OVER);
        sum += (n == PX_OVER) ? 0u : n;
    }
    M.nHits = sum;
    if (over) {
        A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_RETRY;
    } else if (sum > A.hitsSolo) {
        A.teamList[atomicAdd(A.teamCount, 1u)] = qi;
        A.status[qi] = ST_TEAM;
    }
}

template <int BM1, int BM2>
struct QuerySmem {
    uint32_t seen1[BM1], seen2[BM2], seen3[BM2];
    uint8_t q[PW_MAXL];
    signed char corr[PW_MAXL];
    uint32_t v_qi, v_K, v_slow, v_why, v_nHits, v_nSurv, v_nFin, v_cut, v_chunk0, v_nCh;
    int v_minCorr;
};
// Shared memory is budgeted per warp: 8 CTAs x 4 warps must fit the 164 KB carve-out, which leaves 64 KB of L1.  The capped
// per-record scores (u8[256] per warp) live behind the bucket counters in seen1, which is dead by then.
struct WorkSmem {
    uint64_t surv[PW_SCAP];
};

__device__ __noinline__ void warp_bitonic_u64(uint64_t *a, int n_pow2) {
    const int lane = threadIdx.x & 31;
    if (n_pow2 <= 32) {
        // up to one element per lane (the common case: a spacer frame has ~11 winners and a few dozen survivor records):
        // the whole network runs in registers over shuffles -- no shared-memory round trips, no barriers between stages
        __syncwarp();
        uint64_t v = (lane < n_pow2) ? a[lane] : ~0ull;
#pragma unroll 1
        for (int k = 2; k <= n_pow2; k <<= 1) {      // lanes >= n_pow2 hold the maximum and sort among themselves: harmless
#pragma unroll 1
            for (int j = k >> 1; j > 0; j >>= 1) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, v, j);
                const bool takeMin = (((lane & k) == 0) == ((lane & j) == 0));
                v = takeMin ? (o < v ? o : v) : (o > v ? o : v);
            }
        }
        if (lane < n_pow2) a[lane] = v;
        __syncwarp();
        return;
    }
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncwarp();
            for (int t = lane; t < (n_pow2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = ((i & k) == 0);
                const uint64_t x = a[i], y = a[p];
                if (up ? (y < x) : (x < y)) { a[i] = y; a[p] = x; }
            }
        }
    }
    __syncwarp();
}

// out-of-line copies for the bookkeeping kernels: their instruction footprint must stay small (32 warps per SM run different
// phases of different queries; a fully inlined kernel thrashes the instruction cache)
__device__ __noinline__ int ungapped_u16_ool(const uint8_t *__restrict__ tres, const uint32_t *__restrict__ toff, uint32_t id,
                                             const uint8_t *q, int L, const signed char *corr, const signed char *m40u, uint32_t diag16) {
    return ungapped_u16(tres, toff, id, q, L, corr, m40u, diag16);
}

// CTA-wide bitonic sort of (key, payload) pairs by key (ascending); n_pow2 entries, padded by the caller
__device__ __forceinline__ void block_bitonic_pairs(uint64_t *k, uint64_t *v, int n_pow2) {
    for (int kk = 2; kk <= n_pow2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = ((i & kk) == 0);
                const uint64_t x = k[i], y = k[p];
                if (up ? (y < x) : (x < y)) {
                    k[i] = y; k[p] = x;
                    const uint64_t vx = v[i];
                    v[i] = v[p]; v[p] = vx;
                }
            }
        }
    }
    __syncthreads();
}

// per-phase warp-cycle counters (development builds only: make PHASE=1; read back with SPB_DEBUG)
#ifdef SPB_PHASE_TIMING
#define PH_MARK(i) do { const long long _t = clock64(); phase[i] += _t - phT; phT = _t; } while (0)
#else
#define PH_MARK(i) do { } while (0)
#endif

template <int TEAM>
__global__ void __launch_bounds__((TEAM == 1 ? PW_WARPS : TEAM) * 32, TEAM == 1 ? 8 : 4) bookkeep_kernel(PrefArgs A) {
    constexpr bool SOLO = (TEAM == 1);
    // seen1: hashed (target, 8-bit diagonal) pairs seen once; seen2: hashed targets with a repeated pair.  TEAM = 1: 16384 /
    // 4096 bits -- with ~600 hits per frame a dozen pairs collide in seen1 and the coarser seen2 adds a handful of false
    // survivors, so the survivor set usually fits the register sort (<= 32); TEAM = 8: 32768 / 32768 bits.
    constexpr int H1 = SOLO ? 14 : 15, H2 = SOLO ? 12 : 15;
    constexpr int BM1 = (1 << H1) / 32, BM2 = (1 << H2) / 32;
    constexpr int NQS = SOLO ? PW_WARPS : 1;
    constexpr int NWS = SOLO ? PW_WARPS : TEAM;
    __shared__ QuerySmem<BM1, BM2> s_qs[NQS];
    __shared__ WorkSmem s_ws[NWS];
    __shared__ signed char s_m40u[SPB_ALPHABET * SPB_ALPHABET];
    __shared__ uint64_t s_big[SOLO ? 1 : 2 * PW_BIG];     // TEAM = 8: CTA-wide final sort (keys | payload)
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tw = SOLO ? 0 : wid;                  // warp index inside the team
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < SPB_ALPHABET * SPB_ALPHABET; i += blockDim.x) s_m40u[i] = (signed char)A.mat->ungapped[i];
    __syncthreads();
    QuerySmem<BM1, BM2> &S = s_qs[SOLO ? wid : 0];
    WorkSmem &W = s_ws[wid];
    uint8_t *const wcnt = reinterpret_cast<uint8_t *>(S.seen1) + 768 + (SOLO ? 0 : tw * PW_SCAP);   // [PW_SCAP], group phase
    static_assert(768 + (SOLO ? 1 : TEAM) * PW_SCAP <= BM1 * 4, "aliased buffers fit");
    auto team_sync = [&]() { if (SOLO) __syncwarp(); else __syncthreads(); };
    constexpr uint32_t FCAP = SOLO ? PW_FCAP_SOLO : PW_FCAP_TEAM;
    // survivor list | the same grouped by target bucket | per-target winners (final key, truncation key)
    const size_t slice = 2 * (size_t)A.hitCap + 2 * (size_t)FCAP;
    uint64_t *grouped = A.hitScratch + (size_t)(SOLO ? blockIdx.x * PW_WARPS + wid : blockIdx.x) * slice;
    uint64_t *finList = grouped + 2 * (size_t)A.hitCap;
    const uint32_t HCAP = A.hitCap;
    unsigned long long statK = 0, statM = 0, statE = 0, statS = 0;
#ifdef SPB_PHASE_TIMING
    long long phase[6] = {0, 0, 0, 0, 0, 0}, phT = clock64();
#endif
    // f(record, valid) is called warp-wide for the hit records of the chunks this warp owns (TEAM = 8: chunks round-robin)
    // (the lookup kernel writes the hits of chunks 2p and 2p + 1 back to back: the warp-per-query launch walks such a pair as one
    // stretch -- a chunk holds ~145 hits, so two chunks fill three 128-hit trips instead of four)
    auto for_hits = [&](auto f) {
        const uint32_t nCh = S.v_nCh, c0 = S.v_chunk0;
        for (uint32_t ch = (uint32_t)tw; ch < nCh; ch += TEAM) {
            const uint2 cb = __ldcg(A.chunkHits + c0 + ch);
            uint32_t n = cb.y;
            if (SOLO && ((c0 + ch) & 1u) == 0u && ch + 1 < nCh) {
                n += __ldcg(A.chunkHits + c0 + ch + 1).y;
                ch++;
            }
            const uint64_t *src = A.hitArena + (size_t)cb.x;
            for (uint32_t i0 = 0; i0 < n; i0 += 128) {
                uint64_t r[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)(u * 32 + lane);
                    r[u] = (i < n) ? __ldcg(src + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) f(r[u], i0 + (uint32_t)(u * 32 + lane) < n);
            }
        }
    };

    for (;;) {
        PH_MARK(5);
        // ---------------- next query ----------------
        if (tw == 0 && lane == 0) S.v_qi = atomicAdd(A.qCounter, 1u);
        team_sync();
        const uint32_t slot = S.v_qi;
        if (slot >= (A.nListPtr ? *A.nListPtr : A.nList)) break;
        const uint32_t qi = A.qlist[A.listOff + (size_t)slot * A.listStride];
        if (A.status[qi] != (SOLO ? ST_PENDING : ST_TEAM)) {   // decided by expand_kernel (no window, handed over, retry)
            team_sync();
            continue;
        }
        const uint32_t qo = A.qoff[qi];
        const int L = (int)(A.qoff[qi + 1] - qo);
        const QMeta &M = A.meta[qi];
        for (int i = tw * 32 + lane; i < BM1; i += TEAM * 32) S.seen1[i] = 0;
        for (int i = tw * 32 + lane; i < BM2; i += TEAM * 32) S.seen2[i] = 0;
        if (tw == 0) {
            for (int i = lane; i < L; i += 32) { S.q[i] = A.q[qo + i]; S.corr[i] = M.corr[i]; }
            const uint32_t nCh = M.nChunks, c0 = M.chunk0, sum = M.nHits;
            if (lane == 0) {
                const uint32_t why = ST_OK, reason = 0;
                S.v_K = M.K; S.v_nHits = sum; S.v_nSurv = 0; S.v_nFin = 0; S.v_slow = why; S.v_why = reason;
                S.v_minCorr = M.minCorr; S.v_chunk0 = c0; S.v_nCh = nCh;
            }
        }
        team_sync();
        PH_MARK(0);
        const uint32_t K = S.v_K;
        const uint32_t nHits = S.v_nHits;
        const bool live = (S.v_slow == ST_OK);
        // ---------------- pass 1 over the hit stream: (target, 8-bit diagonal) pairs seen once; targets with a repeated pair ----------------
        // An event needs the previous hit of its target on the same 8-bit diagonal (or, first hit, diagonal 0: quirk B), so
        // only targets in which some (target, diagonal) pair occurs twice -- or a diagonal 0 -- can produce one; ALL records of
        // such a target are kept below (hits on other diagonals in between break the adjacency).  Random second hits of a
        // target lie on another diagonal 255 times out of 256, so the survivor set is the true diagonal matches plus a few
        // hash collisions instead of every target hit twice.
        if (live) {
            for_hits([&](uint64_t r, bool valid) {
                if (valid) {
                    const uint32_t x = (uint32_t)(r >> 32) * 2654435761u;
                    const uint32_t d8 = (uint32_t)r & 0xFFu;
                    const uint32_t h = ((x + d8 * 0x9E3779B9u) * 0x85EBCA6Bu) >> (32 - H1);
                    const uint32_t bm = 1u << (h & 31);
                    const uint32_t old = atomicOr(&S.seen1[h >> 5], bm);
                    if ((old & bm) || d8 == 0u) {
                        const uint32_t h2 = x >> (32 - H2);
                        atomicOr(&S.seen2[h2 >> 5], 1u << (h2 & 31));
                    }
                }
            });
        }
        team_sync();
        PH_MARK(1);
        // ---------------- survivors: the records of the marked targets ----------------
        // Pass 2 over the hit stream appends them to a flat list in the team's scratch slice (TEAM = 1 also keeps the first
        // PW_SCAP in the warp's sort buffer); everything after that reads the list, a small fraction of the hits.
        //  * up to PW_SCAP survivors: one group, sorted by one warp -- after a second (target, diagonal) test in shared memory
        //    with other hash functions when they exceed a register sort, since most of them are hash collisions of pass 1
        //    (those grow with the square of the hits);
        //  * more: the same second test runs over the list into a second target bitmap, then the records that pass both are
        //    counted per hash bucket (64 buckets of targets), the buckets are merged into the fewest contiguous groups of at
        //    most PW_SCAP records and the list is scattered group by group behind itself; one warp per group.
        uint32_t *const cnt64 = S.seen1;            // dead after pass 1: [0,64) counts, [64,128) starts, [128,192) fill cursors
        uint64_t *const list = grouped, *const bucketed = grouped + HCAP;
        int nS0 = 0;
        if (live) {
            for_hits([&](uint64_t r, bool valid) {
                const uint32_t h2 = ((uint32_t)(r >> 32) * 2654435761u) >> (32 - H2);
                const bool keep = valid && ((S.seen2[h2 >> 5] >> (h2 & 31)) & 1u);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    uint32_t base = (uint32_t)nS0;
                    if (!SOLO) {
                        if (lane == 0) base = atomicAdd(&S.v_nSurv, (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                    }
                    const uint32_t dst = base + (uint32_t)__popc(m & lt);
                    if (keep) {
                        if (SOLO && dst < PW_SCAP) W.surv[dst] = r;
                        if ((!SOLO || dst >= PW_SCAP) && dst < HCAP) __stcg(list + dst, r);
                    }
                    nS0 += __popc(m);
                }
            });
        }
        team_sync();
        uint32_t nSurv = 0, G = 1;
        bool scatter = false;
        if (live) {
            const uint32_t nL = SOLO ? (uint32_t)nS0 : S.v_nSurv;     // survivors of pass 2
            nSurv = nL;
            if (nL > HCAP) {
                if (tw == 0 && lane == 0) { S.v_slow = ST_CAPACITY; S.v_why = 5; }       // the scratch slice cannot hold them
            } else if (nL <= PW_SCAP) {
                // one group, held by the first warp of the team
                nS0 = 0;
                if (tw == 0) {
                    nS0 = (int)nL;
                    if (!SOLO) for (int i = lane; i < nS0; i += 32) W.surv[i] = __ldcg(list + i);
                    if (nS0 > 32) {
                        for (int i = lane; i < BM1; i += 32) S.seen1[i] = 0;
                        for (int i = lane; i < BM2; i += 32) S.seen2[i] = 0;
                        __syncwarp();
                        for (int i = lane; i < nS0; i += 32) {
                            const uint64_t r = W.surv[i];
                            const uint32_t x = (uint32_t)(r >> 32) * 0x7FEB352Du;
                            const uint32_t d8 = (uint32_t)r & 0xFFu;
                            const uint32_t h = ((x ^ (d8 * 0x846CA68Bu)) * 0x2C1B3C6Du) >> (32 - H1);
                            const uint32_t bm = 1u << (h & 31);
                            const uint32_t old = atomicOr(&S.seen1[h >> 5], bm);
                            if ((old & bm) || d8 == 0u) {
                                const uint32_t h2 = x >> (32 - H2);
                                atomicOr(&S.seen2[h2 >> 5], 1u << (h2 & 31));
                            }
                        }
                        __syncwarp();
                        int outN = 0;
                        for (int c0i = 0; c0i < nS0; c0i += 32) {
                            const int i = c0i + lane;
                            uint64_t r = 0;
                            bool keep = false;
                            if (i < nS0) {
                                r = W.surv[i];
                                const uint32_t h2 = ((uint32_t)(r >> 32) * 0x7FEB352Du) >> (32 - H2);
                                keep = (S.seen2[h2 >> 5] >> (h2 & 31)) & 1u;
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, keep);
                            __syncwarp();
                            if (keep) W.surv[outN + __popc(m & lt)] = r;
                            outN += __popc(m);
                            __syncwarp();
                        }
                        nS0 = outN;
                    }
                    nSurv = (uint32_t)nS0;
                }
            } else {
                // second (target, diagonal) test over the list, then bucket counts of the records that pass both
                const uint32_t nLr = (nL + 31u) & ~31u;
                if (SOLO) for (int i = lane; i < PW_SCAP; i += 32) __stcg(list + i, W.surv[i]);   // the head of the list sits in the sort buffer
                for (int i = tw * 32 + lane; i < BM1; i += TEAM * 32) S.seen1[i] = 0;
                for (int i = tw * 32 + lane; i < BM2; i += TEAM * 32) S.seen3[i] = 0;
                team_sync();
                for (uint32_t i = (uint32_t)(tw * 32 + lane); i < nLr; i += TEAM * 32) {
                    if (i < nL) {
                        const uint64_t r = __ldcg(list + i);
                        const uint32_t x = (uint32_t)(r >> 32) * 0xC2B2AE35u;
                        const uint32_t d8 = (uint32_t)r & 0xFFu;
                        const uint32_t h = ((x ^ (d8 * 0x27D4EB2Fu)) * 0x165667B1u) >> (32 - H1);
                        const uint32_t bm = 1u << (h & 31);
                        const uint32_t old = atomicOr(&S.seen1[h >> 5], bm);
                        if ((old & bm) || d8 == 0u) {
                            const uint32_t h3 = x >> (32 - H2);
                            atomicOr(&S.seen3[h3 >> 5], 1u << (h3 & 31));
                        }
                    }
                }
                team_sync();
                for (int i = tw * 32 + lane; i < 192; i += TEAM * 32) cnt64[i] = 0;
                team_sync();
                for (uint32_t i = (uint32_t)(tw * 32 + lane); i < nLr; i += TEAM * 32) {
                    if (i < nL) {
                        const uint32_t id = (uint32_t)(__ldcg(list + i) >> 32);
                        const uint32_t h3 = (id * 0xC2B2AE35u) >> (32 - H2);
                        if ((S.seen3[h3 >> 5] >> (h3 & 31)) & 1u) atomicAdd(&cnt64[(id * 0x85EBCA6Bu) >> 26], 1u);
                    }
                }
                team_sync();
                // group count: the smallest power of two whose merged buckets all fit one sort (every warp computes the same)
                const uint32_t c0 = cnt64[2 * lane], c1 = cnt64[2 * lane + 1];
                // level l uses 64 >> l groups; the largest merged bucket grows with l, so the last level that fits wins
                uint32_t sum = c0 + c1;
                int lvl = -1;
#pragma unroll 1
                for (int l = 0; l < 7; l++) {
                    if (l >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1 << (l - 2));
                    uint32_t m = (l == 0) ? max(c0, c1) : sum;
#pragma unroll 1
                    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
                    if (m <= PW_SCAP) lvl = l;
                    if (l == 6) nSurv = m;
                }
                G = 64u >> max(lvl, 0);
                team_sync();
                if (lvl < 0) {
                    if (tw == 0 && lane == 0) { S.v_slow = ST_CAPACITY; S.v_why = 6; }   // one bucket alone overflows a sort
                } else {
                    if (tw == 0) {   // bucket starts
                        const uint32_t e0 = (uint32_t)warp_incl_scan_ool((int)(c0 + c1)) - (c0 + c1);
                        cnt64[64 + 2 * lane] = e0;
                        cnt64[64 + 2 * lane + 1] = e0 + c0;
                    }
                    team_sync();
                    for (uint32_t i = (uint32_t)(tw * 32 + lane); i < nLr; i += TEAM * 32) {
                        if (i < nL) {
                            const uint64_t r = __ldcg(list + i);
                            const uint32_t id = (uint32_t)(r >> 32);
                            const uint32_t h3 = (id * 0xC2B2AE35u) >> (32 - H2);
                            if ((S.seen3[h3 >> 5] >> (h3 & 31)) & 1u) {
                                const uint32_t bk = (id * 0x85EBCA6Bu) >> 26;
                                __stcg(bucketed + cnt64[64 + bk] + atomicAdd(&cnt64[128 + bk], 1u), r);
                            }
                        }
                    }
                    scatter = true;
                }
            }
        }
        team_sync();
        PH_MARK(2);
        const int minCorr = S.v_minCorr;
        const int ungBias = abs(A.minSubUngapped) + abs(min(minCorr, 0));
        const int capScore = 255 - ungBias;
        unsigned long long evAcc = 0;
        if (live && S.v_slow == ST_OK) {
            for (uint32_t g = (uint32_t)tw; g < G; g += TEAM) {
                // ---- the group's records ----
                int nS = nS0;
                if (scatter) {
                    const uint32_t per = 64u / G;
                    const uint32_t lo = cnt64[64 + g * per];
                    const uint32_t hi = ((g + 1) * per >= 64u) ? nSurv : cnt64[64 + (g + 1) * per];
                    nS = (int)(hi - lo);
                    for (int i = lane; i < nS; i += 32) W.surv[i] = __ldcg(bucketed + lo + i);
                }
                // ---- sort by (target, stream order) ----
                if (nS > 1) {
                    const int np2 = next_pow2(nS);
                    for (int i = nS + lane; i < np2; i += 32) W.surv[i] = ~0ull;
                    warp_bitonic_u64(W.surv, np2);
                }
                // ---- double-hit events, then de-duplication of consecutive equal events ----
                int nkept = nS;
                for (int pass = 0; pass < 2; pass++) {
                    int outN = 0;
                    uint64_t carryR = 0;
                    for (int c0 = 0; c0 < nkept; c0 += 32) {
                        const int i = c0 + lane;
                        uint64_t r = 0;
                        if (i < nkept) r = W.surv[i];
                        uint64_t rp = __shfl_up_sync(0xffffffffu, r, 1);
                        if (lane == 0) rp = carryR;
                        bool keep = false;
                        if (i < nkept) {
                            const bool same = (i > 0) && ((uint32_t)(rp >> 32) == (uint32_t)(r >> 32));
                            const bool eq8 = (((uint32_t)r ^ (uint32_t)rp) & 0xFFu) == 0u;
                            keep = (pass == 0) ? (same ? eq8 : (((uint32_t)r & 0xFFu) == 0u)) : (!same || !eq8);
                        }
                        carryR = __shfl_sync(0xffffffffu, r, 31);
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        __syncwarp();
                        if (keep) W.surv[outN + __popc(m & lt)] = r;
                        outN += __popc(m);
                        __syncwarp();
                    }
                    nkept = outN;
                }
                evAcc += (unsigned long long)nkept;
                // ---- ungapped diagonal scores; per target the first event reaching the target's maximum ----
                for (int i = lane; i < nkept; i += 32) {
                    const uint64_t r = W.surv[i];
                    const int sc = ungapped_u16_ool(A.tres, A.toff, (uint32_t)(r >> 32), S.q, L, S.corr, s_m40u, (uint32_t)r & 0xFFFFu);
                    wcnt[i] = (uint8_t)min(sc, capScore);
                }
                __syncwarp();
                for (int c0 = 0; c0 < nkept; c0 += 32) {
                    const int i = c0 + lane;
                    bool keep = false;
                    uint64_t f = 0, t = 0;
                    if (i < nkept) {
                        const uint64_t r0 = W.surv[i];
                        const uint32_t id = (uint32_t)(r0 >> 32);
                        const bool head = (i == 0) || ((uint32_t)(W.surv[i - 1] >> 32) != id);
                        if (head) {
                            int bi = i;
                            uint8_t best = wcnt[i];
                            for (int j = i + 1; j < nkept && (uint32_t)(W.surv[j] >> 32) == id; j++)
                                if (wcnt[j] > best) { best = wcnt[j]; bi = j; }
                            if ((int)best >= A.minDiag) {
                                keep = true;
                                const uint32_t dg = (uint32_t)W.surv[bi] & 0xFFFFu;
                                int sc = best;     // saturated counters get their exact score (getResult, QueryMatcher.cpp:371-374)
                                if (sc >= capScore) sc = ungapped_u16_ool(A.tres, A.toff, id, S.q, L, S.corr, s_m40u, dg);
                                f = ((uint64_t)(0xFFFFu - (uint32_t)min(sc, 0xFFFF)) << 48) | ((uint64_t)id << 16) | dg;
                                // truncation order of the reference's stable counting sorts: capped score desc, bin, stream order, id
                                const uint32_t kk = (uint32_t)(W.surv[bi] >> 16) & 0xFFFFu;
                                t = ((uint64_t)(255u - best) << 55) | ((uint64_t)((id + A.idBase) & A.binMask) << 44) | ((uint64_t)kk << 28) | id;
                            }
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (m) {
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&S.v_nFin, (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const uint32_t dst = base + (uint32_t)__popc(m & lt);
                        if (keep && dst < FCAP) { __stcg(finList + 2 * dst, f); __stcg(finList + 2 * dst + 1, t); }
                    }
                }
                __syncwarp();
            }
        }
        team_sync();
        PH_MARK(3);
        // ---------------- final order (score desc, id asc), output ----------------
        if (SOLO) {
            const uint32_t nc = S.v_nFin;
            uint32_t why = S.v_slow;
            uint32_t reason = S.v_why;
            // truncation at max-seqs (and anything beyond one warp sort) is finished by the 8-warp team launch
            if (why == ST_OK && (nc >= (uint32_t)A.maxHits || nc > PW_SCAP)) why = ST_TEAM;
            if (why != ST_OK) {
                if (lane == 0) {
                    A.outStart[qi] = 0; A.outCount[qi] = 0; A.need[qi] = reason;
                    if (why == ST_TEAM) A.teamList2[atomicAdd(A.teamCount2, 1u)] = qi;
                    A.status[qi] = why;
                }
            } else {
                statK += K;
                statM += nHits;
                statS += nSurv;
                const int np2 = next_pow2(max((int)nc, 2));
                // with truncation keys the sort key carries the winner's slot instead of the diagonal (ids are unique, so the
                // low 16 bits never decide the order); diagonal and key are read back from the slot afterwards
                const bool wantT = A.arenaT != nullptr;
                for (int k = lane; k < np2; k += 32) {
                    uint64_t f = (k < (int)nc) ? __ldcg(finList + 2 * k) : ~0ull;
                    if (wantT && k < (int)nc) f = (f & ~0xFFFFull) | (uint64_t)k;
                    W.surv[k] = f;
                }
                if (nc > 1) warp_bitonic_u64(W.surv, np2);
                else __syncwarp();
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(A.cursor, (unsigned long long)nc);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + (unsigned long long)nc > A.arenaCap) {
                    if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_ARENA; }
                } else {
                    for (int k = lane; k < (int)nc; k += 32) {
                        const uint64_t f = W.surv[k];
                        spb_hit h;
                        h.id = (uint32_t)(f >> 16);
                        h.score = (int32_t)(0xFFFFu - (uint32_t)(f >> 48));
                        uint32_t dg = (uint32_t)(f & 0xFFFFull);
                        if (wantT) {
                            const uint64_t t = __ldcg(finList + 2 * dg + 1);
                            dg = (uint32_t)(__ldcg(finList + 2 * dg) & 0xFFFFull);
                            const uint32_t capped = 255u - (uint32_t)(t >> 55);
                            A.arenaT[base + k] = tkey_make(capped, (int)capped >= capScore, (uint32_t)(t >> 44) & 0x7FFu,
                                                           (uint32_t)(t >> 28) & 0xFFFFu);
                        }
                        h.diag = (int32_t)(short)(uint16_t)dg;
                        A.arena[base + k] = h;
                    }
                    if (lane == 0) { A.outStart[qi] = base; A.outCount[qi] = nc; A.status[qi] = ST_OK; }
                }
            }
        } else {
            // whole CTA.  QueryMatcher.cpp:131-164: with >= max-seqs candidates the score histogram fixes the cut score; above
            // max-seqs the candidates are taken in the order of the reference's stable counting sorts (truncation key).
            uint64_t *const bigK = s_big, *const bigV = s_big + PW_BIG;
            const uint32_t nc = S.v_nFin;
            const uint32_t maxHits = (uint32_t)A.maxHits;
            const int tid = threadIdx.x;
            if (tid == 0) {
                S.v_nSurv = 0;                    // reused: selected candidates
                if (S.v_slow == ST_OK && nc > FCAP) { S.v_slow = ST_CAPACITY; S.v_why = 7; }
            }
            __syncthreads();
            const bool ok0 = (S.v_slow == ST_OK);
            const bool trunc = ok0 && nc >= maxHits;
            if (trunc) {
                uint32_t *const histS = S.seen2;  // dead after the survivor passes; 1024 words in the team kernel
                for (int i = tid; i < 256; i += blockDim.x) histS[i] = 0;
                __syncthreads();
                for (uint32_t k = tid; k < nc; k += blockDim.x) atomicAdd(&histS[255u - (uint32_t)(__ldcg(finList + 2 * k + 1) >> 55)], 1u);
                __syncthreads();
                if (tid == 0) {     // computeScoreThreshold (QueryMatcher.h:199-209)
                    uint32_t found = 0, thr = 0;
                    for (int sc = 255; sc >= 1; sc--) {
                        found += histS[sc];
                        if (found >= maxHits) { thr = (uint32_t)sc; break; }
                    }
                    const uint32_t thrD = max(thr, (uint32_t)A.minDiag);
                    S.v_cut = thrD;
                    // all remaining candidates saturated -> rescoreHits (QueryMatcher.cpp:486-505) lives in prefilter_kernel
                    if ((int)thrD >= capScore) { S.v_slow = ST_CAPACITY; S.v_why = 8; }
                }
                __syncthreads();
            }
            const bool ok1 = (S.v_slow == ST_OK);
            uint32_t ns = nc;                     // rows to output
            if (ok1) {
                if (nc > maxHits) {
                    const uint32_t thrD = S.v_cut;
                    for (uint32_t k = tid; k < nc; k += blockDim.x) {
                        const uint64_t t = __ldcg(finList + 2 * k + 1);
                        if (255u - (uint32_t)(t >> 55) >= thrD) {
                            const uint32_t dst = atomicAdd(&S.v_nSurv, 1u);
                            if (dst < PW_BIG) { bigK[dst] = t; bigV[dst] = __ldcg(finList + 2 * k); }
                        }
                    }
                    __syncthreads();
                    const uint32_t nSel = S.v_nSurv;
                    if (nSel > PW_BIG) {
                        if (tid == 0) { S.v_slow = ST_CAPACITY; S.v_why = 7; }
                    } else {
                        const int np2 = next_pow2(max((int)nSel, 2));
                        for (int i = (int)nSel + tid; i < np2; i += blockDim.x) { bigK[i] = ~0ull; bigV[i] = ~0ull; }
                        block_bitonic_pairs(bigK, bigV, np2);
                        ns = maxHits;
                        const int nq2 = next_pow2(max((int)ns, 2));
                        // the first max-seqs in truncation order, re-sorted by the final key (truncation key as payload)
                        for (int i = tid; i < nq2; i += blockDim.x) {
                            const uint64_t t = bigK[i], f = bigV[i];
                            bigK[i] = (i < (int)ns) ? f : ~0ull;
                            bigV[i] = (i < (int)ns) ? t : ~0ull;
                        }
                        block_bitonic_pairs(bigK, bigV, nq2);
                    }
                } else if (nc > PW_BIG) {
                    if (tid == 0) { S.v_slow = ST_CAPACITY; S.v_why = 7; }
                } else {
                    const int np2 = next_pow2(max((int)nc, 2));
                    for (int i = tid; i < np2; i += blockDim.x) {
                        bigK[i] = (i < (int)nc) ? __ldcg(finList + 2 * i) : ~0ull;
                        bigV[i] = (i < (int)nc) ? __ldcg(finList + 2 * i + 1) : ~0ull;
                    }
                    block_bitonic_pairs(bigK, bigV, np2);
                }
            }
            __syncthreads();
            if (S.v_slow != ST_OK) {
                if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = S.v_slow; A.need[qi] = S.v_why; }
            } else {
                if (tid == 0) s_base = atomicAdd(A.cursor, (unsigned long long)ns);
                __syncthreads();
                const unsigned long long base = s_base;
                if (base + (unsigned long long)ns > A.arenaCap) {
                    if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_ARENA; }
                } else {
                    for (uint32_t k = tid; k < ns; k += blockDim.x) {
                        const uint64_t f = bigK[k];
                        spb_hit h;
                        h.id = (uint32_t)(f >> 16);
                        h.score = (int32_t)(0xFFFFu - (uint32_t)(f >> 48));
                        h.diag = (int32_t)(short)(uint16_t)(f & 0xFFFFull);
                        A.arena[base + k] = h;
                        if (A.arenaT) {
                            const uint64_t t = bigV[k];
                            const uint32_t capped = 255u - (uint32_t)(t >> 55);
                            A.arenaT[base + k] = tkey_make(capped, (int)capped >= capScore, (uint32_t)(t >> 44) & 0x7FFu,
                                                           (uint32_t)(t >> 28) & 0xFFFFu);
                        }
                    }
                    if (tid == 0) {
                        A.outStart[qi] = base; A.outCount[qi] = ns; A.status[qi] = ST_OK;
                        statK += K;
                        statM += nHits;
                        statS += nSurv;
                    }
                }
            }
            __syncthreads();
        }
        team_sync();     // every lane has read v_slow above (the TEAM = 1 hand-over branch has no barrier of its own)
        if (tw == 0 && lane == 0) S.v_slow = (A.status[qi] == ST_OK) ? ST_OK : ST_CAPACITY;
        team_sync();
        if (S.v_slow == ST_OK) statE += evAcc;     // events of the groups this warp processed
        PH_MARK(4);
    }
#ifdef SPB_PHASE_TIMING
    if (lane == 0)
        for (int i = 0; i < 6; i++) atomicAdd(A.cursor + (SOLO ? 40 : 48) + i, (unsigned long long)phase[i]);
#endif
    // counters of the queries finished on this path (the others are counted by the kernel that finishes them)
    if (lane == 0) {
        if (statK) atomicAdd(A.cursor + 1, statK);
        if (statM) atomicAdd(A.cursor + 2, statM);
        if (statE) atomicAdd(A.cursor + 3, statE);
        if (statS) atomicAdd(A.cursor + 4, statS);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Target-split merge (multi-GPU): the hit lists of one query coming from the target shards are concatenated and put
// into the final order (score desc, global id asc).  Per-target rules are shard-local, so below max-seqs the union IS
// the unsplit list.  At or above max-seqs the union is cut like the unsplit list: every shard list is a prefix of its
// candidates in truncation order (tkey_make above), so the first max-seqs of the union in (truncation key, id) order are
// the reference's selection (QueryMatcher.cpp:131-164).  Without keys such a query is reported (err[0] = query + 1); a
// cut that falls among saturated counters is the rescoreHits corner and is reported too (err[1]).
// One CTA per query of the owned range; segment (s, q) of the shard-major receive buffer starts at segOff[s * nq + q].
// Dynamic shared memory: u64 keysF[MG_CAP] | u64 keysT[MG_CAP] | u16 ord[MG_CAP].
// ---------------------------------------------------------------------------------------------------------
#define MG_CAP 4096
#define MG_SMEM (MG_CAP * (8 + 8 + 2))
__global__ void __launch_bounds__(256) merge_shards_kernel(const spb_hit *__restrict__ recv, const uint64_t *__restrict__ recvT,
                                                           const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ segOff,
                                                           uint32_t nq, uint32_t nShards, uint32_t qLo,
                                                           const uint64_t *__restrict__ outOff, spb_hit *__restrict__ out,
                                                           uint32_t maxSeqs, unsigned int *__restrict__ err) {
    extern __shared__ __align__(16) unsigned char mg_smem[];
    uint64_t *keysF = reinterpret_cast<uint64_t *>(mg_smem);
    uint64_t *keysT = keysF + MG_CAP;
    uint16_t *ord = reinterpret_cast<uint16_t *>(keysT + MG_CAP);
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const uint64_t o = outOff[qLo + q];
    const uint32_t nOut = (uint32_t)(outOff[qLo + q + 1] - o);
    uint32_t total = 0;
    for (uint32_t s = 0; s < nShards; s++) total += cnt[(size_t)s * nq + q];
    if (total == 0) return;
    const bool cut = total >= maxSeqs;
    if ((cut && recvT == nullptr) || total > MG_CAP) {
        if (threadIdx.x == 0) atomicMax(&err[0], qLo + q + 1);
        return;
    }
    uint32_t base = 0;
    for (uint32_t s = 0; s < nShards; s++) {
        const uint32_t n = cnt[(size_t)s * nq + q];
        const uint64_t so = segOff[(size_t)s * nq + q];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const spb_hit h = recv[so + i];
            keysF[base + i] = ((uint64_t)(0xFFFFu - (uint32_t)min(max(h.score, 0), 0xFFFF)) << 48) | ((uint64_t)h.id << 16) |
                              (uint64_t)((uint32_t)h.diag & 0xFFFFu);
            if (cut) keysT[base + i] = recvT[so + i];
        }
        base += n;
    }
    uint32_t ns = total;
    if (total > maxSeqs) {
        // the first max-seqs in (truncation key, id) order
        const int np2 = next_pow2((int)total);
        for (int i = threadIdx.x; i < np2; i += blockDim.x) ord[i] = i < (int)total ? (uint16_t)i : (uint16_t)0xFFFFu;
        __syncthreads();
        block_bitonic_sort(ord, np2, [&](uint16_t a, uint16_t b) {
            if (a == 0xFFFFu || b == 0xFFFFu) return a != 0xFFFFu && b == 0xFFFFu;
            const uint64_t ta = keysT[a], tb = keysT[b];
            if (ta != tb) return ta < tb;
            return (uint32_t)(keysF[a] >> 16) < (uint32_t)(keysF[b] >> 16);
        });
        ns = maxSeqs;
        // keysT is dead after the selection except for the corner test: keep the last selected key, reuse the array
        const uint64_t lastT = keysT[ord[ns - 1]];
        __syncthreads();
        if ((lastT >> 63) == 0) {
            if (threadIdx.x == 0) atomicMax(&err[1], qLo + q + 1);
            return;
        }
        for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) keysT[i] = keysF[ord[i]];
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) keysF[i] = keysT[i];
        __syncthreads();
    } else if (cut) {
        // exactly max-seqs candidates: nothing is dropped, but the reference still takes the rescoreHits path when even the
        // lowest counter is saturated (thr >= 255 - bias)
        __syncthreads();
        bool allSat = true;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) allSat = allSat && ((keysT[i] >> 63) == 0);
        if (__syncthreads_and(allSat)) {
            if (threadIdx.x == 0) atomicMax(&err[1], qLo + q + 1);
            return;
        }
    }
    const int np2 = next_pow2(max((int)ns, 2));
    __syncthreads();
    for (int i = (int)ns + threadIdx.x; i < np2; i += blockDim.x) keysF[i] = ~0ull;
    block_bitonic_sort(keysF, np2, [](uint64_t a, uint64_t b) { return a < b; });
    if (ns != nOut) {   // the offsets were built from min(total, maxSeqs)
        if (threadIdx.x == 0) atomicMax(&err[2], qLo + q + 1);
        return;
    }
    for (uint32_t i = threadIdx.x; i < ns; i += blockDim.x) {
        const uint64_t f = keysF[i];
        spb_hit h;
        h.id = (uint32_t)(f >> 16);
        h.score = (int32_t)(0xFFFFu - (uint32_t)(f >> 48));
        h.diag = (int32_t)(short)(uint16_t)(f & 0xFFFFull);
        out[o + i] = h;
    }
}

// rows per query of the full batch after the merge: min(sum over shards, max-seqs) inside the owned range, 0 outside
__global__ void merge_totals_kernel(const uint32_t *__restrict__ cnt, uint32_t nq, uint32_t nShards, uint32_t qLo, uint32_t Q,
                                    uint32_t maxSeqs, uint64_t *__restrict__ totals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > Q) return;
    uint64_t t = 0;
    if (i >= qLo && i < qLo + nq && i < Q)
        for (uint32_t s = 0; s < nShards; s++) t += cnt[(size_t)s * nq + (i - qLo)];
    totals[i] = min(t, (uint64_t)maxSeqs);
}

__global__ void widen_u32_kernel(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[i] = 0;
}

// gather per-query slices of the arena into query order
__global__ void gather_hits_kernel(const spb_hit *__restrict__ arena, const uint64_t *__restrict__ start, const uint64_t *__restrict__ outOff,
                                   uint32_t Q, spb_hit *__restrict__ out) {
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const uint64_t s = start[qi], o = outOff[qi];
    const uint32_t n = (uint32_t)(outOff[qi + 1] - o);
    for (uint32_t k = threadIdx.x & 31; k < n; k += 32) out[o + k] = arena[s + k];
}

__global__ void gather_keys_kernel(const uint64_t *__restrict__ arena, const uint64_t *__restrict__ start, const uint64_t *__restrict__ outOff,
                                   uint32_t Q, uint64_t *__restrict__ out) {
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const uint64_t s = start[qi], o = outOff[qi];
    const uint32_t n = (uint32_t)(outOff[qi + 1] - o);
    for (uint32_t k = threadIdx.x & 31; k < n; k += 32) out[o + k] = arena[s + k];
}

__global__ void count_to_u64_kernel(const uint32_t *__restrict__ cnt, uint32_t Q, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Q) out[i] = cnt[i];
    if (i == Q) out[i] = 0;
}

// status scan: list of queries needing the global-memory variant, max need, error flags
__global__ void status_scan_kernel(const uint32_t *__restrict__ status, const uint32_t *__restrict__ need, uint32_t Q,
                                   uint32_t *__restrict__ list, unsigned int *__restrict__ ctr, uint32_t listStatus,
                                   uint32_t *__restrict__ retryList) {
    // ctr: [0] entries of `list` (queries whose status is listStatus), [1] maxNeed, [2] arena overflow,
    //      [3] layout error (query index + 1), [4] unprocessed, [5] entries of `retryList` (ST_RETRY; NULL: none expected)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    const uint32_t s = status[i];
    if (s == listStatus) {
        list[atomicAdd(&ctr[0], 1u)] = i;
        atomicMax(&ctr[1], need[i]);
    } else if (s == ST_ARENA) {
        atomicOr(&ctr[2], 1u);
    } else if (s == ST_LAYOUT) {
        atomicMax(&ctr[3], i + 1);
    } else if (s == ST_RETRY && retryList) {
        retryList[atomicAdd(&ctr[5], 1u)] = i;
    } else if (s != ST_OK) {
        atomicMax(&ctr[4], i + 1);
    }
}

// test hook: the similar-k-mer list of one 6-mer through the enumerator of expand_kernel (one warp)
__global__ void similar_kmers_test_kernel(const uint32_t *row3, const uint16_t *cnt3, const uint8_t *k6, int thr, uint32_t *out,
                                          uint32_t cap, uint32_t *n) {
    const int x1 = k6[0] + 20 * k6[1] + 400 * k6[2];
    const int x2 = k6[3] + 20 * k6[4] + 400 * k6[5];
    const int best1 = (int)(row3[(size_t)x1 * SPB_N3] >> 16) - SPB_SCORE_OFF;
    const int best2 = (int)(row3[(size_t)x2 * SPB_N3] >> 16) - SPB_SCORE_OFF;
    const int cut1 = thr - best2;
    __shared__ ExpRound s_r;
    uint32_t K = 0;
    if (cut1 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) {
        K = window_kmers<false>(row3, cnt3, x1, x2, thr, best1, cut1, 0u, nullptr, nullptr);
        if (K <= cap) window_kmers<true>(row3, cnt3, x1, x2, thr, best1, cut1, 0u, out, &s_r);
    }
    if (threadIdx.x == 0) *n = K;
}

}  // namespace spb

using namespace spb;

extern "C" int spb_similar_kmers(spb_ctx *c, const uint8_t *kmer6, int thr, uint32_t *out, uint32_t cap, uint32_t *n) {
    SPB_CHECK(c && kmer6 && out && n, "spb_similar_kmers: NULL argument");
    SPB_CHECK(c->have_index, "spb_similar_kmers: build the index first");
    SPB_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < 6; i++) SPB_CHECK(kmer6[i] < 20, "spb_similar_kmers: residue out of range");
    DevBuf dk, dout, dn;
    if (dk.ensure(8) || dout.ensure((size_t)cap * 4 + 4) || dn.ensure(4)) return 1;
    cudaError_t e = cudaMemcpyAsync(dk.p, kmer6, 6, cudaMemcpyHostToDevice, c->stream);
    similar_kmers_test_kernel<<<1, 32, 0, c->stream>>>(c->d_row3.as<uint32_t>(), c->d_cnt3.as<uint16_t>(), dk.as<uint8_t>(), thr,
                                                        dout.as<uint32_t>(), cap, dn.as<uint32_t>());
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(n, dn.p, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && *n > cap) {
        dk.release(); dout.release(); dn.release();
        spb::set_error("spb_similar_kmers: %u similar k-mers exceed the capacity %u of the output array", *n, cap);
        return 1;
    }
    if (e == cudaSuccess && *n) e = cudaMemcpy(out, dout.p, (size_t)*n * 4, cudaMemcpyDeviceToHost);
    dk.release(); dout.release(); dn.release();
    SPB_CHECK(e == cudaSuccess, "spb_similar_kmers: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// resident pipeline
// ------------------------------------------------------------------------------------------------------
extern "C" int spb_queries_set(spb_ctx *c, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q) {
    return spb::queries_install(c, q_residues, false, q_offsets, Q);
}

// residues on the host (spb_queries_set) or already on the device (frames produced by spb_spacer_frames); offsets on the host
int spb::queries_install(spb_ctx *c, const uint8_t *q_residues, bool on_device, const uint64_t *q_offsets, uint32_t Q) {
    SPB_CHECK(c && q_offsets, "spb_queries_set: NULL argument");
    SPB_CUDA(cudaSetDevice(c->device));
    const uint64_t NQ = q_offsets[Q];
    SPB_CHECK(NQ < 0xFFFFFFF0ull, "spb_queries_set: query batch too large for 32-bit offsets; split the batch");
    SPB_CHECK(Q == 0 || q_residues, "spb_queries_set: NULL residues");
    std::vector<uint32_t> off32((size_t)Q + 1);
    uint32_t maxL = 0;
    for (uint32_t i = 0; i <= Q; i++) {
        SPB_CHECK(q_offsets[i] <= NQ && (i == 0 || q_offsets[i] >= q_offsets[i - 1]), "spb_queries_set: offsets not monotone at %u", i);
        off32[i] = (uint32_t)q_offsets[i];
        if (i) maxL = std::max(maxL, (uint32_t)(q_offsets[i] - q_offsets[i - 1]));
    }
    // processing order of the prefilter: by length, longest first (counting sort, stable)
    std::vector<uint32_t> order((size_t)Q);
    {
        std::vector<uint32_t> start((size_t)maxL + 2, 0);
        for (uint32_t i = 0; i < Q; i++) start[maxL - (off32[i + 1] - off32[i]) + 1]++;
        for (size_t l = 1; l < start.size(); l++) start[l] += start[l - 1];
        for (uint32_t i = 0; i < Q; i++) order[start[maxL - (off32[i + 1] - off32[i])]++] = i;
    }
    if (c->d_q.ensure(NQ + 64) || c->d_qoff.ensure(((size_t)Q + 1) * 4) || c->d_qoff64.ensure(((size_t)Q + 1) * 8) ||
        c->d_qorder.ensure((size_t)Q * 4 + 4)) return 1;
    if (Q) SPB_CUDA(cudaMemcpyAsync(c->d_qorder.p, order.data(), (size_t)Q * 4, cudaMemcpyHostToDevice, c->stream));
    if (NQ) SPB_CUDA(cudaMemcpyAsync(c->d_q.p, q_residues, NQ, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(c->d_qoff.p, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(c->d_qoff64.p, q_offsets, ((size_t)Q + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    c->h_qlen.resize(Q);
    for (uint32_t i = 0; i < Q; i++) c->h_qlen[i] = off32[i + 1] - off32[i];
    c->h_qorder.swap(order);
    c->Q = Q;
    c->q_res = NQ;
    c->q_maxlen = maxL;
    c->q_minlen = Q ? *std::min_element(c->h_qlen.begin(), c->h_qlen.end()) : 0u;
    c->have_queries = true;
    c->have_pref = false;
    c->have_pairs = false;
    c->have_aln = false;
    return 0;
}

extern "C" int spb_prefilter_resident(spb_ctx *c, const spb_prefilter_params *p) {
    SPB_CHECK(c && p, "spb_prefilter_resident: NULL argument");
    SPB_CHECK(c->have_index, "spb_prefilter: build the index first (spb_index_build)");
    SPB_CHECK(c->have_queries, "spb_prefilter: no resident query batch (spb_queries_set)");
    SPB_CHECK(p->bin_size >= 2 && p->bin_size <= 2048 && (p->bin_size & (p->bin_size - 1)) == 0,
              "spb_prefilter: bin_size %d must be a power of two in 2..2048", p->bin_size);
    SPB_CHECK(p->max_seqs >= 1, "spb_prefilter: max_seqs must be >= 1");
    SPB_CHECK(p->min_diag_score >= 0 && p->min_diag_score <= 255, "spb_prefilter: min_diag_score out of range");
    SPB_CHECK(!p->emit_trunc_keys || p->max_seqs <= 4096, "spb_prefilter: emit_trunc_keys needs max_seqs <= 4096 (got %d)", p->max_seqs);
    SPB_CHECK(c->q_maxlen <= PF_MAXL, "spb_prefilter: query of %u residues exceeds the supported query length %d of the spacer path",
              c->q_maxlen, PF_MAXL);
    SPB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint32_t Q = c->Q;
    c->have_pref = false;
    c->have_tkeys = false;
    memset(&c->pref_stats, 0, sizeof(c->pref_stats));
    c->pref_stats.n_queries = Q;
    if (c->d_hitoff.ensure(((size_t)Q + 1) * 8)) return 1;
    if (Q == 0) {
        SPB_CUDA(cudaMemsetAsync(c->d_hitoff.p, 0, 8, st));
        c->n_hits = 0;
        c->have_pref = true;
        return 0;
    }
    PrefArgs A;
    memset(&A, 0, sizeof(A));
    A.q = c->d_q.as<uint8_t>();
    A.qoff = c->d_qoff.as<uint32_t>();
    A.Q = Q;
    A.tres = c->d_tres.as<uint8_t>();
    A.toff = c->d_toff.as<uint32_t>();
    A.T = c->T;
    A.nShards = c->n_shards;
    A.shardR = c->shard_R;
    for (int i = 0; i < SPB_MAX_SHARDS; i++) {
        A.directS[i] = reinterpret_cast<const ulonglong2 *>(c->shard_direct[i]);
        A.postS[i] = reinterpret_cast<const uint64_t *>(c->shard_post[i]);
        SPB_CHECK(i >= c->n_shards || (A.directS[i] && A.postS[i]), "spb_prefilter: shard %d of the k-mer table is not attached "
                  "(spb_index_shard_attach_ipc / _ctx)", i);
    }
    A.kbits = c->d_kbits.as<uint32_t>();
    A.row3 = c->d_row3.as<uint32_t>();
    A.cnt3 = c->d_cnt3.as<uint16_t>();
    A.mat = c->d_mat.as<Matrices>();
    A.pat.span = c->span;
    for (int i = 0; i < 6; i++) A.pat.pos[i] = c->pat[i];
    A.kmerThr = p->kmer_thr;
    A.maxHits = (int)std::min<uint64_t>((uint64_t)p->max_seqs, std::max<uint64_t>(c->T, 1));
    A.minDiag = p->min_diag_score;
    A.compBias = p->comp_bias;
    A.minSubUngapped = 0;
    for (int i = 0; i < SPB_ALPHABET * SPB_ALPHABET; i++) A.minSubUngapped = std::min(A.minSubUngapped, (int)c->hm.ungapped[i]);
    A.binMask = (uint32_t)p->bin_size - 1;
    A.idBase = p->id_base;
    A.exactKmer = p->exact_kmer ? 1 : 0;

    // scratch: s_a start u64[Q], s_b count u32[Q], s_c status u32[Q], s_d need u32[Q], s_e slow list u32[Q], s_f counters, s_g arena
    if (c->s_a.ensure((size_t)Q * 8) || c->s_b.ensure((size_t)Q * 4) || c->s_c.ensure((size_t)Q * 4) || c->s_d.ensure((size_t)Q * 4) ||
        c->s_e.ensure((size_t)Q * 4) || c->s_f.ensure(512)) return 1;
    uint64_t arenaCap = std::min<uint64_t>((uint64_t)Q * (uint64_t)A.maxHits, std::max<uint64_t>((uint64_t)Q * 24, 1u << 20));
    float kernelMs = 0;
    double phaseMs[4] = {0, 0, 0, 0};   // expand, lookup, bookkeeping TEAM = 1, TEAM = 8 (summed over tiles)
    uint64_t nSlow = 0;
    unsigned long long h_cursor[24];
    unsigned int dbgWhy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int h_ctr[8];
    SPB_CUDA(cudaEventRecord(c->ev2, st));
    for (int attempt = 0;; attempt++) {
        if (c->s_g.ensure(arenaCap * sizeof(spb_hit) + 64)) return 1;
        if (p->emit_trunc_keys && c->s_tk.ensure(arenaCap * sizeof(uint64_t) + 64)) return 1;
        A.arena = c->s_g.as<spb_hit>();
        A.arenaT = p->emit_trunc_keys ? c->s_tk.as<uint64_t>() : nullptr;
        A.arenaCap = arenaCap;
        A.cursor = c->s_f.as<unsigned long long>();
        A.outStart = c->s_a.as<uint64_t>();
        A.outCount = c->s_b.as<uint32_t>();
        A.status = c->s_c.as<uint32_t>();
        A.need = c->s_d.as<uint32_t>();
        A.qlist = nullptr;
        A.gScratch = nullptr;
        A.cap = PF_SMEM_CAP;
        unsigned int *d_ctr = reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 8);
        SPB_CUDA(cudaMemsetAsync(c->s_f.p, 0, 512, st));
        SPB_CUDA(cudaMemsetAsync(c->s_c.p, 0xFF, (size_t)Q * 4, st));
        const size_t smem = (size_t)PF_SMEM_CAP * sizeof(Rec);
        // stage 1: expand -> lookup -> bookkeeping, tile by tile (a tile = every nTiles-th query of the longest-first order, so
        // all tiles have the same length mix; its slot / hit arenas are reused by the next tile in stream order).  The
        // TEAM = 1 bookkeeping lists the queries it hands to the TEAM = 8 launch on the device -- no host round trip inside a
        // tile; one status scan + read-back per round.  Queries that found the slot arena full (ST_RETRY) form the next round.
        int soloPerSm = 8;
        if (const char *e = getenv("SPB_SOLO_CTAS")) soloPerSm = std::max(1, std::min(8, atoi(e)));   // tuning aids
        if (const char *e = getenv("SPB_SOLO_CARVEOUT"))    // percent of the 228 KB given to shared memory (the rest is L1)
            cudaFuncSetAttribute(bookkeep_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
        int teamPerSm = 4;
        if (const char *e = getenv("SPB_TEAM_CTAS")) teamPerSm = std::max(1, std::min(4, atoi(e)));
        int lookupPerSm = 6, expandPerSm = 10;    // r2d sweep: 6 / 8 / 10 expand CTAs per SM -> 0.319 / 0.290 / 0.274 ms
        if (const char *e = getenv("SPB_LOOKUP_CTAS")) lookupPerSm = std::max(1, std::min(16, atoi(e)));
        if (const char *e = getenv("SPB_EXPAND_CTAS")) expandPerSm = std::max(1, std::min(16, atoi(e)));
        const int soloBlocks = c->n_sm * soloPerSm, teamBlocks = c->n_sm * teamPerSm;
        A.hitsSolo = 4096;     // r2h sweep (both launches side by side): 1024 / 2048 / 4096 -> 0.63 / 0.47 / 0.43 ms of bookkeeping
        if (const char *e = getenv("SPB_HITS_SOLO")) A.hitsSolo = (uint32_t)std::max(1, atoi(e));
        const uint32_t hitCapSolo = std::max<uint32_t>(A.hitsSolo, PW_SCAP), hitCapTeam = 32768;
        const size_t sliceSolo = 2 * (size_t)hitCapSolo + 2 * PW_FCAP_SOLO, sliceTeam = 2 * (size_t)hitCapTeam + 2 * PW_FCAP_TEAM;
        uint32_t tileQ = 49152;
        if (const char *e = getenv("SPB_PF_TILEQ")) tileQ = (uint32_t)std::max(256, atoi(e));
        // chunks of 256 slots a tile may allocate: 12 per query is 2.5 x the mean at -s 5.7; grown when a round had to retry
        uint32_t chunkCap = (uint32_t)std::min<uint64_t>(1u << 19, (uint64_t)std::min(Q, tileQ) * 12 + 1024);
        if (const char *e = getenv("SPB_PF_CHUNKS")) chunkCap = (uint32_t)std::max(256, atoi(e));
        uint32_t hitsPerChunk = 256;
        if (const char *e = getenv("SPB_PF_HITS_PER_CHUNK")) hitsPerChunk = (uint32_t)std::max(1, atoi(e));
        const size_t soloScratch = (size_t)soloBlocks * PW_WARPS * sliceSolo;      // the two bookkeeping launches run concurrently
        if (c->s_j.ensure((soloScratch + (size_t)teamBlocks * sliceTeam) * sizeof(uint64_t)) ||
            c->s_k.ensure((size_t)Q * 4 + 64) || c->p_team2.ensure((size_t)Q * 4 + 64) || c->p_retry.ensure((size_t)Q * 4 + 64) || c->p_meta.ensure((size_t)Q * sizeof(QMeta))) return 1;
        A.hitScratch = c->s_j.as<uint64_t>();
        A.meta = c->p_meta.as<QMeta>();
        // queries with more similar k-mers than this are filled chunk by chunk (expand_heavy_kernel); 65535 switches that off
        A.heavyK = 2048;
        if (const char *e = getenv("SPB_EXPAND_HEAVY")) A.heavyK = (uint32_t)std::max(1, std::min(65535, atoi(e)));
        A.heavyCap = std::min<uint32_t>(std::min(Q, tileQ), 16384u);
        A.heavyItemCap = A.heavyCap * 32u;     // K <= 65535: at most 256 chunks = 32 groups per query
        if (c->p_heavy.ensure((size_t)A.heavyCap * sizeof(HeavyQ) + (size_t)A.heavyItemCap * 4)) return 1;
        A.heavyTab = c->p_heavy.as<HeavyQ>();
        A.heavyItems = reinterpret_cast<uint32_t *>(c->p_heavy.as<unsigned char>() + (size_t)A.heavyCap * sizeof(HeavyQ));
        A.teamList = c->s_k.as<uint32_t>();
        A.teamList2 = c->p_team2.as<uint32_t>();
        if (!c->aux_stream) {
            SPB_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
            SPB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            SPB_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
        // tile counters: [0] expand scheduler, [1] TEAM = 1 scheduler, [2] TEAM = 8 scheduler, [3] chunks allocated, [4] team list,
        // [5] scheduler of the second TEAM = 8 launch, [6] its list (hand-overs of the TEAM = 1 launch)
        unsigned int *tc = reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 16);
        const uint32_t *curList = c->d_qorder.as<uint32_t>();   // longest queries first (tail of a launch)
        uint32_t curN = Q;
        if (A.exactKmer) {
            // --exact-kmer-matching: one k-mer per window -- nothing to expand; every query goes straight to the CTA-per-query
            // kernel of stage 2, which enumerates the window's own k-mer instead of its similar k-mers
            curN = 0;
            SPB_CUDA(cudaMemcpyAsync(c->s_e.p, curList, (size_t)Q * 4, cudaMemcpyDeviceToDevice, st));
            memset(h_ctr, 0, sizeof(h_ctr));
            h_ctr[0] = Q;
        }
        unsigned int nTilesRun = 0;
        bool arenaOverflow = false;
        for (int round = 0; curN; round++) {
            // hit arena: 0.56 hits per slot at GenBank-phage size, i.e. ~145 per chunk; 256 per chunk leaves 1.75 x
            const uint32_t hitArenaCap = (uint32_t)std::min<uint64_t>(0xFFFFFF00ull, (uint64_t)chunkCap * hitsPerChunk);
            if (c->p_recs.ensure((size_t)chunkCap * PX_CH * 4) || c->p_ctab.ensure((size_t)chunkCap * sizeof(uint2)) ||
                c->p_ccnt.ensure((size_t)chunkCap * sizeof(uint2)) || c->p_harena.ensure((size_t)hitArenaCap * 8)) return 1;
            A.recs = c->p_recs.as<uint32_t>();
            A.chunkTab = c->p_ctab.as<uint2>();
            A.chunkHits = c->p_ccnt.as<uint2>();
            A.hitArena = c->p_harena.as<uint64_t>();
            A.chunkCap = chunkCap;
            A.hitArenaCap = hitArenaCap;
            const uint32_t nTiles = (curN + tileQ - 1) / tileQ;
            SPB_CUDA(cudaEventRecord(c->ev0, st));
            for (uint32_t t = 0; t < nTiles; t++) {
                while (c->ev_phase.size() < (size_t)(nTilesRun + 1) * 5) {
                    cudaEvent_t ev;
                    SPB_CUDA(cudaEventCreate(&ev));
                    c->ev_phase.push_back(ev);
                }
                cudaEvent_t *ep = c->ev_phase.data() + (size_t)nTilesRun * 5;
                SPB_CUDA(cudaMemsetAsync(tc, 0, 48, st));
                A.qlist = curList;
                A.listOff = t;
                A.listStride = nTiles;
                A.nList = (curN - t + nTiles - 1) / nTiles;
                A.nListPtr = nullptr;
                A.chunkCursor = tc + 3;
                A.heavyCount = tc + 7;
                A.heavyItemCount = tc + 8;
                A.teamCount = tc + 4;
                A.teamCount2 = tc + 6;
                A.qCounter = tc + 0;
                SPB_CUDA(cudaEventRecord(ep[0], st));
                expand_kernel<<<c->n_sm * expandPerSm, PW_WARPS * 32, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
                if (A.heavyK < 65535u) {
                    expand_heavy_kernel<<<c->n_sm * expandPerSm, PW_WARPS * 32, 0, st>>>(A);
                    SPB_CUDA(cudaGetLastError()); c->n_launches++;
                }
                SPB_CUDA(cudaEventRecord(ep[1], st));
                if (c->n_shards > 1) lookup_kernel<true><<<c->n_sm * lookupPerSm, PX_CH, 0, st>>>(A);   // persistent: lookupPerSm CTAs per SM
                else lookup_kernel<false><<<c->n_sm * lookupPerSm, PX_CH, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
                SPB_CUDA(cudaEventRecord(ep[2], st));
                classify_kernel<<<(A.nList + 255) / 256, 256, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
                SPB_CUDA(cudaEventRecord(c->ev_fork, st));
                // heavy queries: CTA per query on the second stream, next to the warp-per-query launch
                {
                    PrefArgs T8 = A;
                    T8.qCounter = tc + 2;
                    T8.qlist = c->s_k.as<uint32_t>();
                    T8.listOff = 0;
                    T8.listStride = 1;
                    T8.nListPtr = tc + 4;
                    T8.hitCap = hitCapTeam;
                    T8.hitScratch = c->s_j.as<uint64_t>() + soloScratch;
                    SPB_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
                    bookkeep_kernel<PW_TEAM><<<teamBlocks, PW_TEAM * 32, 0, c->aux_stream>>>(T8);
                    SPB_CUDA(cudaGetLastError()); c->n_launches++;
                    SPB_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
                }
                A.qCounter = tc + 1;
                A.hitCap = hitCapSolo;
                bookkeep_kernel<1><<<soloBlocks, PW_WARPS * 32, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
                SPB_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
                SPB_CUDA(cudaEventRecord(ep[3], st));
                // what the warp launch handed over at its end (lists cut at max-seqs, > 256 winners): usually nothing
                A.qCounter = tc + 5;
                A.qlist = c->p_team2.as<uint32_t>();
                A.listOff = 0;
                A.listStride = 1;
                A.nListPtr = tc + 6;
                A.hitCap = hitCapTeam;
                A.hitScratch = c->s_j.as<uint64_t>() + soloScratch;
                bookkeep_kernel<PW_TEAM><<<teamBlocks, PW_TEAM * 32, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
                A.hitScratch = c->s_j.as<uint64_t>();
                SPB_CUDA(cudaEventRecord(ep[4], st));
                nTilesRun++;
            }
            A.qlist = nullptr;
            A.nListPtr = nullptr;
            SPB_CUDA(cudaEventRecord(c->ev1, st));
            SPB_CUDA(cudaMemsetAsync(d_ctr, 0, 32, st));
            status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_e.as<uint32_t>(), d_ctr, ST_CAPACITY,
                                                                c->p_retry.as<uint32_t>());
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, 32, cudaMemcpyDeviceToHost, st));
            SPB_CUDA(cudaStreamSynchronize(st));
            float ms = 0;
            SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            kernelMs += ms;
            SPB_CHECK(h_ctr[3] == 0, "spb_prefilter: query %u is longer than the supported %d residues", h_ctr[3] - 1, PF_MAXL);
            SPB_CHECK(h_ctr[4] == 0, "spb_prefilter: query %u was not processed", h_ctr[4] - 1);
            arenaOverflow = h_ctr[2] != 0;
            if (arenaOverflow) break;
            SPB_CHECK(h_ctr[5] < curN || chunkCap < (1u << 21), "spb_prefilter: the slot arena cannot hold a single query");
            curList = c->p_retry.as<uint32_t>();
            curN = h_ctr[5];
            if (curN) {   // next round: larger arenas (the skew of this DB / sensitivity is above the default headroom)
                chunkCap = (uint32_t)std::min<uint64_t>(1u << 21, (uint64_t)chunkCap * 2);
                hitsPerChunk = std::min<uint32_t>(hitsPerChunk * 2, 2048);
            }
        }
        for (unsigned int t = 0; t < nTilesRun; t++) {
            cudaEvent_t *ep = c->ev_phase.data() + (size_t)t * 5;
            float e0 = 0, e1 = 0, e2 = 0, e3 = 0;
            SPB_CUDA(cudaEventElapsedTime(&e0, ep[0], ep[1]));
            SPB_CUDA(cudaEventElapsedTime(&e1, ep[1], ep[2]));
            SPB_CUDA(cudaEventElapsedTime(&e2, ep[2], ep[3]));
            SPB_CUDA(cudaEventElapsedTime(&e3, ep[3], ep[4]));
            phaseMs[0] += e0; phaseMs[1] += e1; phaseMs[2] += e2; phaseMs[3] += e3;
        }
#ifdef SPB_PHASE_TIMING
        if (getenv("SPB_DEBUG")) {
            unsigned long long ph[16];
            cudaMemcpy(ph, c->s_f.as<unsigned long long>() + 40, sizeof(ph), cudaMemcpyDeviceToHost);
            for (int t = 0; t < 2; t++) {
                double tot = 0;
                for (int i = 0; i < 6; i++) tot += (double)ph[8 * t + i];
                fprintf(stderr, "phase cycles TEAM=%d: setup %.1f%% lookup %.1f%% survivors %.1f%% groups %.1f%% final %.1f%% idle/sched %.1f%% (total %.3g warp-cycles)\n",
                        t ? 8 : 1, 100 * ph[8 * t] / tot, 100 * ph[8 * t + 1] / tot, 100 * ph[8 * t + 2] / tot, 100 * ph[8 * t + 3] / tot,
                        100 * ph[8 * t + 4] / tot, 100 * ph[8 * t + 5] / tot, tot);
            }
        }
#endif
        uint32_t nList = h_ctr[0], maxNeed = 0, capG = 2048;
        float ms = 0;
        nSlow = nList;
        if (getenv("SPB_DEBUG") && nList) {
            std::vector<uint32_t> hl(nList), hn(Q);
            cudaMemcpy(hl.data(), c->s_e.p, (size_t)nList * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hn.data(), A.need, (size_t)Q * 4, cudaMemcpyDeviceToHost);
            memset(dbgWhy, 0, sizeof(dbgWhy));
            for (uint32_t i = 0; i < nList; i++) dbgWhy[std::min<uint32_t>(hn[hl[i]], 7)]++;
        }
        // stage 2: queries the warp kernel handed over (long, many hits, truncation at max-seqs): one CTA each,
        // records in shared memory
        if (nList && !arenaOverflow) {
            if (c->s_i.ensure((size_t)Q * 4)) return 1;
            SPB_CUDA(cudaEventRecord(c->ev0, st));
            A.qlist = c->s_e.as<uint32_t>();
            prefilter_kernel<false><<<nList, PF_THREADS, smem, st>>>(A);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cudaEventRecord(c->ev1, st));
            SPB_CUDA(cudaMemsetAsync(d_ctr, 0, 32, st));
            status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_i.as<uint32_t>(), d_ctr, ST_CAPACITY, nullptr);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, 32, cudaMemcpyDeviceToHost, st));
            SPB_CUDA(cudaStreamSynchronize(st));
            SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            kernelMs += ms;
            SPB_CHECK(h_ctr[3] == 0, "spb_prefilter: query %u is longer than the supported %d residues", h_ctr[3] - 1, PF_MAXL);
            SPB_CHECK(h_ctr[4] == 0, "spb_prefilter: query %u was not processed", h_ctr[4] - 1);
            arenaOverflow = h_ctr[2] != 0;
            nList = h_ctr[0];
            maxNeed = h_ctr[1];
            if (nList) SPB_CUDA(cudaMemcpyAsync(c->s_e.p, c->s_i.p, (size_t)nList * 4, cudaMemcpyDeviceToDevice, st));
        }
        // stage 3: queries whose kept-hit list does not fit shared memory run again on global-memory scratch
        while (nList && !arenaOverflow) {
            while (capG < maxNeed) capG <<= 1;
            SPB_CHECK(capG <= (1u << 28), "spb_prefilter: hit list of a query exceeds 2^28 records");
            const uint32_t chunkMax = (uint32_t)std::max<uint64_t>(1, ((uint64_t)4 << 30) / ((uint64_t)capG * sizeof(Rec)));
            if (c->s_h.ensure((size_t)std::min(chunkMax, nList) * capG * sizeof(Rec))) return 1;
            SPB_CUDA(cudaEventRecord(c->ev0, st));
            for (uint32_t done = 0; done < nList; done += chunkMax) {
                const uint32_t chunk = std::min(chunkMax, nList - done);
                A.qlist = c->s_e.as<uint32_t>() + done;
                A.gScratch = c->s_h.as<Rec>();
                A.cap = capG;
                prefilter_kernel<true><<<chunk, PF_THREADS, 0, st>>>(A);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
            }
            SPB_CUDA(cudaEventRecord(c->ev1, st));
            // which of them still need more room (their status stays ST_CAPACITY)?
            if (c->s_i.ensure((size_t)Q * 4)) return 1;
            SPB_CUDA(cudaMemsetAsync(d_ctr, 0, 32, st));
            status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_i.as<uint32_t>(), d_ctr, ST_CAPACITY, nullptr);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, 32, cudaMemcpyDeviceToHost, st));
            SPB_CUDA(cudaStreamSynchronize(st));
            SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            kernelMs += ms;
            SPB_CHECK(h_ctr[3] == 0 && h_ctr[4] == 0, "spb_prefilter: a query failed on the global-memory path");
            arenaOverflow = h_ctr[2] != 0;
            nList = h_ctr[0];
            if (nList) {
                SPB_CUDA(cudaMemcpyAsync(c->s_e.p, c->s_i.p, (size_t)nList * 4, cudaMemcpyDeviceToDevice, st));
                maxNeed = std::max(h_ctr[1], capG * 2);
            }
        }
        if (!arenaOverflow) break;
        SPB_CHECK(arenaCap < (uint64_t)Q * (uint64_t)A.maxHits && attempt < 8, "spb_prefilter: result arena overflow could not be resolved");
        arenaCap = std::min<uint64_t>((uint64_t)Q * (uint64_t)A.maxHits, arenaCap * 4);
    }
    SPB_CUDA(cudaMemcpyAsync(h_cursor, c->s_f.p, 192, cudaMemcpyDeviceToHost, st));

    // ---- compact the arena into query order: offsets = exclusive scan of the counts
    if (c->s_i.ensure(((size_t)Q + 1) * 8)) return 1;
    count_to_u64_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(c->s_b.as<uint32_t>(), Q, c->s_i.as<uint64_t>());
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    size_t tmpBytes = 0;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, c->s_i.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), (int)(Q + 1), st));
    if (c->s_cub.ensure(tmpBytes)) return 1;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, c->s_i.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), (int)(Q + 1), st));
    uint64_t nHits = 0;
    SPB_CUDA(cudaMemcpyAsync(&nHits, c->d_hitoff.as<uint64_t>() + Q, 8, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    if (c->d_hits.ensure(std::max<uint64_t>(nHits, 1) * sizeof(spb_hit))) return 1;
    if (nHits) {
        gather_hits_kernel<<<(Q + 3) / 4, 128, 0, st>>>(c->s_g.as<spb_hit>(), c->s_a.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), Q,
                                                        c->d_hits.as<spb_hit>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    if (p->emit_trunc_keys) {
        if (c->d_hit_tkeys.ensure(std::max<uint64_t>(nHits, 1) * sizeof(uint64_t))) return 1;
        if (nHits) {
            gather_keys_kernel<<<(Q + 3) / 4, 128, 0, st>>>(c->s_tk.as<uint64_t>(), c->s_a.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), Q,
                                                            c->d_hit_tkeys.as<uint64_t>());
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
        }
        c->have_tkeys = true;
    }
    SPB_CUDA(cudaEventRecord(c->ev3, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    float totalMs = 0;
    SPB_CUDA(cudaEventElapsedTime(&totalMs, c->ev2, c->ev3));
    c->n_hits = nHits;
    c->pref_max_seqs = (uint32_t)A.maxHits;
    if (getenv("SPB_DEBUG")) {
        unsigned int tcs[8];
        cudaMemcpy(tcs, reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 16), 32, cudaMemcpyDeviceToHost);
        fprintf(stderr, "spb_prefilter: %u queries; last tile: %u chunks, %u queries taken by the 8-warp team launch; %llu handed to the "
                "CTA kernel (long query %u, k-mers %u, chunk/scratch %u, bucket %u, max-seqs %u); expand %.3f lookup %.3f "
                "bookkeeping %.3f + %.3f ms\n", Q, tcs[3], tcs[4], (unsigned long long)nSlow, dbgWhy[1], dbgWhy[4], dbgWhy[5],
                dbgWhy[6], dbgWhy[7], phaseMs[0], phaseMs[1], phaseMs[2], phaseMs[3]);
    }
    spb_prefilter_result &R = c->pref_stats;
    R.n_hits = nHits;
    R.n_kmers = h_cursor[1];
    R.n_db_matches = h_cursor[2];
    R.n_events = h_cursor[3];
    R.n_slow_path = nSlow;
    R.kernel_ms = kernelMs;
    R.total_ms = totalMs;
    R.expand_ms = phaseMs[0];
    R.lookup_ms = phaseMs[1];
    R.bookkeep_ms = phaseMs[2] + phaseMs[3];
    // algorithmic bytes of one pass over the hit stream: per similar k-mer one 3-mer row entry (4 B) and two
    // CSR offsets (8 B), per index hit one posting (8 B), per query its residues, per output hit 12 B
    R.algo_bytes = R.n_kmers * 12ull + R.n_db_matches * 8ull + c->q_res + nHits * 12ull;
    c->have_pref = true;
    return 0;
}

extern "C" int spb_prefilter_stats(spb_ctx *c, spb_prefilter_result *out) {
    SPB_CHECK(c && out, "spb_prefilter_stats: NULL argument");
    SPB_CHECK(c->have_pref, "spb_prefilter_stats: no resident prefilter result");
    *out = c->pref_stats;
    out->offsets = nullptr;
    out->hits = nullptr;
    return 0;
}

extern "C" void spb_free_prefilter_result(spb_prefilter_result *r) {
    if (!r) return;
    free(r->offsets);
    free(r->hits);
    free(r);
}

extern "C" int spb_prefilter_fetch(spb_ctx *c, spb_prefilter_result **out) {
    SPB_CHECK(c && out, "spb_prefilter_fetch: NULL argument");
    SPB_CHECK(c->have_pref, "spb_prefilter_fetch: no resident prefilter result");
    SPB_CUDA(cudaSetDevice(c->device));
    *out = nullptr;
    const uint32_t Q = c->Q;
    spb_prefilter_result *res = (spb_prefilter_result *)calloc(1, sizeof(spb_prefilter_result));
    *res = c->pref_stats;
    res->offsets = (uint64_t *)calloc((size_t)Q + 1, sizeof(uint64_t));
    res->hits = (spb_hit *)malloc(std::max<uint64_t>(c->n_hits, 1) * sizeof(spb_hit));
    cudaError_t e = cudaMemcpyAsync(res->offsets, c->d_hitoff.p, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && c->n_hits) e = cudaMemcpyAsync(res->hits, c->d_hits.p, c->n_hits * sizeof(spb_hit), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        spb_free_prefilter_result(res);
        spb::set_error("spb_prefilter_fetch: %s", cudaGetErrorString(e));
        return 1;
    }
    *out = res;
    return 0;
}

extern "C" int spb_prefilter(spb_ctx *c, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                             const spb_prefilter_params *p, spb_prefilter_result **out) {
    SPB_CHECK(c && q_offsets && p && out, "spb_prefilter: NULL argument");
    *out = nullptr;
    if (spb_queries_set(c, q_residues, q_offsets, Q)) return 1;
    if (spb_prefilter_resident(c, p)) return 1;
    return spb_prefilter_fetch(c, out);
}


// ------------------------------------------------------------------------------------------------------
// multi-GPU target split: device views of the resident result and the merge of the shard lists
// ------------------------------------------------------------------------------------------------------
extern "C" int spb_prefilter_device_result(spb_ctx *c, spb_hit **d_hits, uint64_t **d_offsets, uint64_t *n_hits) {
    SPB_CHECK(c && d_hits && d_offsets && n_hits, "spb_prefilter_device_result: NULL argument");
    SPB_CHECK(c->have_pref, "spb_prefilter_device_result: no resident prefilter result");
    *d_hits = c->d_hits.as<spb_hit>();
    *d_offsets = c->d_hitoff.as<uint64_t>();
    *n_hits = c->n_hits;
    return 0;
}

extern "C" int spb_prefilter_device_trunc_keys(spb_ctx *c, uint64_t **d_keys) {
    SPB_CHECK(c && d_keys, "spb_prefilter_device_trunc_keys: NULL argument");
    SPB_CHECK(c->have_pref && c->have_tkeys, "spb_prefilter_device_trunc_keys: run the prefilter with emit_trunc_keys = 1 first");
    *d_keys = c->d_hit_tkeys.as<uint64_t>();
    return 0;
}

extern "C" int spb_prefilter_merge_shards(spb_ctx *c, uint32_t q_lo, uint32_t q_hi, const spb_hit *d_recv, const uint64_t *d_recv_keys,
                                          uint64_t n_recv, const uint32_t *d_counts, uint32_t n_shards, int32_t max_seqs) {
    SPB_CHECK(c && (d_counts || q_hi == q_lo), "spb_prefilter_merge_shards: NULL argument");
    SPB_CHECK(c->have_queries, "spb_prefilter_merge_shards: no resident query batch");
    SPB_CHECK(q_lo <= q_hi && q_hi <= c->Q, "spb_prefilter_merge_shards: query range %u..%u outside the batch of %u", q_lo, q_hi, c->Q);
    SPB_CHECK(n_shards >= 1 && max_seqs >= 1, "spb_prefilter_merge_shards: bad shard count / max_seqs");
    SPB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint32_t Q = c->Q, nq = q_hi - q_lo;
    const size_t nSeg = (size_t)n_shards * nq;
    SPB_CUDA(cudaEventRecord(c->ev2, st));
    // s_a: segment counts widened u64[nSeg+1], s_b: segment offsets u64[nSeg+1], s_c: totals u64[Q+1], s_f: error flags
    if (c->s_a.ensure((nSeg + 1) * 8) || c->s_b.ensure((nSeg + 1) * 8) || c->s_c.ensure(((size_t)Q + 1) * 8) || c->s_f.ensure(512) ||
        c->d_hitoff.ensure(((size_t)Q + 1) * 8)) return 1;
    SPB_CUDA(cudaMemsetAsync(c->s_f.p, 0, 16, st));
    size_t tmpBytes = 0;
    uint64_t nIn = 0;
    if (nSeg) {
        widen_u32_kernel<<<(unsigned)((nSeg + 1 + 255) / 256), 256, 0, st>>>(d_counts, nSeg, c->s_a.as<uint64_t>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, c->s_a.as<uint64_t>(), c->s_b.as<uint64_t>(), (int)(nSeg + 1), st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, c->s_a.as<uint64_t>(), c->s_b.as<uint64_t>(), (int)(nSeg + 1), st));
        SPB_CUDA(cudaMemcpyAsync(&nIn, c->s_b.as<uint64_t>() + nSeg, 8, cudaMemcpyDeviceToHost, st));
    }
    merge_totals_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(d_counts, nq, n_shards, q_lo, Q, (uint32_t)max_seqs, c->s_c.as<uint64_t>());
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, c->s_c.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), (int)(Q + 1), st));
    if (c->s_cub.ensure(tmpBytes)) return 1;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, c->s_c.as<uint64_t>(), c->d_hitoff.as<uint64_t>(), (int)(Q + 1), st));
    uint64_t nHits = 0;
    SPB_CUDA(cudaMemcpyAsync(&nHits, c->d_hitoff.as<uint64_t>() + Q, 8, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    SPB_CHECK(nIn == n_recv, "spb_prefilter_merge_shards: the counts add up to %llu hits but %llu were received",
              (unsigned long long)nIn, (unsigned long long)n_recv);
    // the receive buffer may alias nothing of ours; the merged lists replace the resident result
    // (grow-only second hit buffer of the context, swapped in below: no allocation on the per-batch path)
    DevBuf &merged = c->d_hits_alt;
    if (merged.ensure(std::max<uint64_t>(nHits, 1) * sizeof(spb_hit))) return 1;
    unsigned int h_err[4] = {0, 0, 0, 0};
    if (nq && nIn) {
        // per device and cheap: set before every launch (a process may hold contexts on several GPUs)
        SPB_CUDA(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
        merge_shards_kernel<<<nq, 256, MG_SMEM, st>>>(d_recv, d_recv_keys, d_counts, c->s_b.as<uint64_t>(), nq, n_shards, q_lo,
                                                      c->d_hitoff.as<uint64_t>(), merged.as<spb_hit>(), (uint32_t)max_seqs,
                                                      reinterpret_cast<unsigned int *>(c->s_f.p));
        cudaError_t e = cudaGetLastError(); c->n_launches++;
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_err, c->s_f.p, 16, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { spb::set_error("spb_prefilter_merge_shards: %s", cudaGetErrorString(e)); return 1; }
    }
    if (h_err[0]) {
        if (d_recv_keys)
            spb::set_error("spb_prefilter_merge_shards: query %u has more than %d candidates across the target shards (merge capacity)",
                           h_err[0] - 1, MG_CAP);
        else
            spb::set_error("spb_prefilter_merge_shards: query %u reaches max-seqs (%d) across the target shards; its truncation order needs "
                           "the shards' truncation keys (emit_trunc_keys) -- or run this batch with query split", h_err[0] - 1, max_seqs);
        return 2;
    }
    if (h_err[1]) {
        spb::set_error("spb_prefilter_merge_shards: the max-seqs (%d) cut of query %u falls among saturated diagonal counters (rescoreHits, "
                       "QueryMatcher.cpp:486-505), which needs the unsplit candidate set -- run this batch with query split",
                       max_seqs, h_err[1] - 1);
        return 2;
    }
    SPB_CHECK(h_err[2] == 0, "spb_prefilter_merge_shards: internal row count mismatch at query %u", h_err[2] - 1);
    std::swap(c->d_hits, c->d_hits_alt);   // the merged lists become the resident result
    c->n_hits = nHits;
    SPB_CUDA(cudaEventRecord(c->ev3, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
    c->pref_stats.n_hits = nHits;
    c->pref_stats.total_ms += ms;
    c->have_pref = true;
    c->have_tkeys = false;
    c->have_pairs = false;
    c->have_aln = false;
    return 0;
}
