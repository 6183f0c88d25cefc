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
    const uint32_t *koff;
    const uint64_t *post;
    const uint32_t *row3;
    const uint16_t *cnt3;
    const Matrices *mat;
    Pattern6 pat;
    int kmerThr, maxHits, minDiag, compBias, minSubUngapped;
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
    // warp-per-query kernel
    const ulonglong2 *direct;  // [20^6] packed k-mer table, 16 bytes per k-mer (index.cu)
    uint64_t *hitScratch;    // [warps in the grid][hitCap] hit records, L2-resident
    uint32_t hitCap;
    unsigned int *qCounter;  // dynamic query scheduler
    const unsigned int *nListPtr;  // device: entries of qlist (team launch reads it, no host round trip); NULL: nList
    uint32_t nList;
    uint32_t kCapSolo;       // TEAM = 1: queries with more similar k-mers go to the TEAM = 8 launch
};

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
        if (!hasX && w < P - 1) {   // quirk A: the last window is never counted
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
            const uint32_t nKmers = warp_unit_kmers(A.row3, A.cnt3, x1, x2, (int)s_thr[w], (int)s_n1[w], a0, s_excl[wid], s_i1[wid],
                [&](uint32_t kmer, uint32_t a, uint32_t b, bool valid) {
                    uint32_t o0 = 0, len = 0;
                    if (valid) {
                        o0 = __ldg(A.koff + kmer);
                        len = __ldg(A.koff + kmer + 1) - o0;
                    }
                    if (pass == 0) {
                        nm += len;
                        for (uint32_t z = 0; z < len; z++) {
                            const uint32_t h = id_hash((uint32_t)(__ldg(A.post + o0 + z) >> 16));
                            const uint32_t m = 1u << (h & 31);
                            const uint32_t old = atomicOr(&s_seen1[h >> 5], m);
                            if (old & m) atomicOr(&s_seen2[h >> 5], m);
                        }
                    } else {
                        for (uint32_t z = 0; z < len; z++) {
                            const uint64_t pe = __ldg(A.post + o0 + z);   // id << 16 | pos
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
                });
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
// Team prefilter: the production path.  A query is owned by a TEAM of warps -- one warp for the bulk of the spacer
// frames (TEAM = 1, four independent warps per CTA, no CTA barriers), eight warps (one CTA) for the queries whose
// similar-k-mer list is too long for one warp to walk without becoming the tail of the launch (TEAM = 8).
//
// Same rules as prefilter_kernel above, re-laid for memory-level parallelism:
//  * the index is read through the packed direct table (one 8-byte read per similar k-mer; postings are touched only
//    for k-mers occurring in >= 2 targets), four independent lookups per lane in flight;
//  * the hit stream is enumerated ONCE.  All similar k-mers of the query are numbered 0..K-1 in reference order
//    (window, outer neighbour a, inner neighbour b).  Inner counts are constant over outer neighbours of equal score,
//    so the numbering is a run-length table in shared memory (one run per distinct outer score of a window: first a,
//    inner count, first k); lane k finds its run by binary search and derives (a, b) with one division;
//  * a hit record is one u64 (id << 32 | k << 16 | 16-bit diagonal) whose integer order is (target, stream order).
//    Records go to a per-team scratch slice in global memory (written and re-read out of L2) while two shared-memory
//    bitmaps mark targets seen once / twice; only records of targets seen twice (or with 8-bit diagonal 0, quirk B)
//    are pulled back -- in groups of targets (hash partition) small enough for a warp to sort in shared memory --
//    and run through the double-hit, de-duplication, ungapped-scoring and best-diagonal rules; the per-target winners
//    are collected as final sort keys behind the hit records and sorted once at the end.
// Anything that does not fit (query longer than 64, > 256 runs, >= 65536 k-mers, hit / group / candidate overflow,
// truncation at max-seqs) is left to prefilter_kernel via status = ST_CAPACITY; queries with more than kCapSolo
// k-mers are passed from the TEAM = 1 launch to the TEAM = 8 launch via ST_TEAM.  Results are identical on every path.
// =========================================================================================================
#define PW_WARPS 4           // independent warps per CTA in the TEAM = 1 launch
#define PW_TEAM 8            // warps per query in the team launch
#define PW_MAXL 64
#define PW_RUNCAP 256
#define PW_SCAP 256          // records a warp sorts at once
#define PW_FCAP_SOLO 512     // per-target winners of one query (slots behind the hit records)
#define PW_FCAP_TEAM 4096
#define PW_BIG 1024          // TEAM = 8: records the CTA sorts at once in the final stage
#define PW_U 2               // consecutive similar k-mers per lane (16-byte table entries: as many bytes in flight as 4 x 8)
#define ST_TEAM 4u

struct Run { uint16_t a0, n2, cum, w; };   // cum holds the first concatenated row index until the prefix pass

template <int BMW, int NRUNS>
struct QuerySmem {
    uint32_t seen1[BMW], seen2[BMW];
    Run runs_[NRUNS];                    // TEAM = 1 keeps its runs in the (then idle) sort buffer of the warp instead
    uint16_t x1[PW_MAXL], x2[PW_MAXL];   // first / second 3-mer of every window
    uint16_t rowbase[PW_MAXL + 2];       // exclusive prefix of n1
    short thr[PW_MAXL];
    uint8_t q[PW_MAXL];
    signed char corr[PW_MAXL];
    uint32_t v_qi, v_K, v_nRuns, v_slow, v_why, v_nHits, v_nSurv, v_nFin;
    int v_minCorr;
};
// Shared memory is budgeted per warp: 8 CTAs x 4 warps must fit the 164 KB carve-out, which leaves 64 KB of L1 for the row3
// reads (with the 196 KB carve-out the kernel is 12 % slower).  Set-up temporaries therefore live in buffers that are idle at
// that time: the composition bias (float[64]) and the outer-neighbour counts (u16[64]) in the head of the sort buffer (dead
// before the runs are written there), the capped per-record scores (u8[256]) behind the bucket counters in seen1.
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
        for (int k = 2; k <= 32; k <<= 1) {
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

__device__ __noinline__ int warp_incl_scan_ool(int v) { return warp_incl_scan(v); }

// out-of-line copies for the team kernels: their instruction footprint must stay small (32 warps per SM run different
// phases of different queries; an inlined kernel of 128 KB thrashes the instruction cache)
__device__ __noinline__ int ungapped_u16_ool(const uint8_t *__restrict__ tres, const uint32_t *__restrict__ toff, uint32_t id,
                                             const uint8_t *q, int L, const signed char *corr, const signed char *m40u, uint32_t diag16) {
    return ungapped_u16(tres, toff, id, q, L, corr, m40u, diag16);
}

// hit records: warp-aggregated append to the team's scratch slice + seen-once / seen-twice bitmaps
__device__ __noinline__ void pw_emit(bool have, uint64_t pe, uint32_t k, uint32_t w, uint32_t *nHitsCtr, uint64_t *scratch, uint32_t HCAP,
                                     uint32_t *seen1, uint32_t *seen2, int hashShift) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, have);
    if (m == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(nHitsCtr, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (have) {
        const uint32_t id = (uint32_t)(pe >> 16);
        const uint32_t dg = (w - (uint32_t)(pe & 0xFFFFull)) & 0xFFFFu;
        const uint32_t dst = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (dst < HCAP) __stcg(scratch + dst, ((uint64_t)id << 32) | ((uint64_t)k << 16) | dg);
        const uint32_t h = (id * 2654435761u) >> hashShift;
        const uint32_t bm = 1u << (h & 31);
        const uint32_t old = atomicOr(&seen1[h >> 5], bm);
        if ((old & bm) || (dg & 0xFFu) == 0u) atomicOr(&seen2[h >> 5], bm);
    }
}

// posting lists of the k-mers held by the lanes of a warp (len == 0: none), flattened over the warp: element p of the
// concatenated lists belongs to the last lane whose exclusive prefix is <= p (coalesced reads, no per-lane walks);
// the PW_U = 2 consecutive k-mers a lane owns in one lookup iteration are served together: both lists of a lane are concatenated
// (k-mer 2 * lane first), so one scan and one pass serve the 64 k-mers of the iteration instead of two of each
__device__ __noinline__ void pw_walk2(uint32_t len0, uint32_t off0, uint32_t len1, uint32_t off1, uint32_t wPack, uint32_t k0,
                                      const uint64_t *__restrict__ post, uint32_t *nHitsCtr, uint64_t *scratch, uint32_t HCAP,
                                      uint32_t *seen1, uint32_t *seen2, int hashShift) {
    const int lane = threadIdx.x & 31;
    const uint32_t lenS = len0 + len1;
    const uint32_t inc = (uint32_t)warp_incl_scan_ool((int)lenS);
    const uint32_t excl = inc - lenS;
    const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll 1
    for (uint32_t p0 = 0; p0 < T; p0 += 32) {
        const uint32_t pp = p0 + lane;
        int j = 0;
#pragma unroll 1
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = j + step;
            const uint32_t e = __shfl_sync(0xffffffffu, excl, cand & 31);
            if (cand < 32 && e <= pp) j = cand;
        }
        const uint32_t rel = pp - __shfl_sync(0xffffffffu, excl, j);
        const uint32_t sL0 = __shfl_sync(0xffffffffu, len0, j);
        const uint32_t sO0 = __shfl_sync(0xffffffffu, off0, j);
        const uint32_t sO1 = __shfl_sync(0xffffffffu, off1, j);
        const uint32_t sW = __shfl_sync(0xffffffffu, wPack, j);
        const bool second = rel >= sL0;
        const bool have = pp < T;
        uint64_t pe = 0;
        if (have) pe = __ldg(post + (second ? sO1 + (rel - sL0) : sO0 + rel));
        pw_emit(have, pe, k0 + 2u * (uint32_t)j + (second ? 1u : 0u), (sW >> (second ? 8 : 0)) & 0xFFu, nHitsCtr, scratch, HCAP, seen1, seen2,
                hashShift);
    }
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
__global__ void __launch_bounds__((TEAM == 1 ? PW_WARPS : TEAM) * 32, TEAM == 1 ? 8 : 4) prefilter_team_kernel(PrefArgs A) {
    constexpr bool SOLO = (TEAM == 1);
    constexpr int BMW = SOLO ? 256 : 1024;          // bitmap words: 8192 / 32768 bits
    constexpr int NQS = SOLO ? PW_WARPS : 1;
    constexpr int NWS = SOLO ? PW_WARPS : TEAM;
    __shared__ QuerySmem<BMW, SOLO ? 1 : PW_RUNCAP> s_qs[NQS];
    __shared__ WorkSmem s_ws[NWS];
    __shared__ signed char s_m40u[SPB_ALPHABET * SPB_ALPHABET];
    __shared__ uint64_t s_big[SOLO ? 1 : 2 * PW_BIG];     // TEAM = 8: CTA-wide final sort (keys | payload)
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tw = SOLO ? 0 : wid;                  // warp index inside the team
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < SPB_ALPHABET * SPB_ALPHABET; i += blockDim.x) s_m40u[i] = (signed char)A.mat->ungapped[i];
    __syncthreads();
    QuerySmem<BMW, SOLO ? 1 : PW_RUNCAP> &S = s_qs[SOLO ? wid : 0];
    WorkSmem &W = s_ws[wid];
    Run *const runs = SOLO ? reinterpret_cast<Run *>(W.surv) : S.runs_;
    float *const cbuf = reinterpret_cast<float *>(W.surv);                       // [PW_MAXL], set-up only
    uint16_t *const n1buf = reinterpret_cast<uint16_t *>(W.surv) + 2 * PW_MAXL;   // [PW_MAXL] behind it, set-up only
    uint8_t *const wcnt = reinterpret_cast<uint8_t *>(S.seen1) + 768 + (SOLO ? 0 : tw * PW_SCAP);   // [PW_SCAP], group phase
    static_assert(PW_MAXL * 4 + PW_MAXL * 2 <= PW_SCAP * 8 && 768 + (SOLO ? 1 : TEAM) * PW_SCAP <= BMW * 4, "aliased buffers fit");
    auto team_sync = [&]() { if (SOLO) __syncwarp(); else __syncthreads(); };
    constexpr uint32_t FCAP = SOLO ? PW_FCAP_SOLO : PW_FCAP_TEAM;
    // hit records | records grouped by target bucket | per-target winners (final key, truncation key)
    const size_t slice = 2 * (size_t)A.hitCap + 2 * (size_t)FCAP;
    uint64_t *scratch = A.hitScratch + (size_t)(SOLO ? blockIdx.x * PW_WARPS + wid : blockIdx.x) * slice;
    uint64_t *finList = scratch + 2 * (size_t)A.hitCap;
    const uint32_t HCAP = A.hitCap;
    unsigned long long statK = 0, statM = 0, statE = 0, statS = 0;
#ifdef SPB_PHASE_TIMING
    long long phase[6] = {0, 0, 0, 0, 0, 0}, phT = clock64();
#endif

    for (;;) {
        PH_MARK(5);
        // ---------------- next query ----------------
        if (tw == 0 && lane == 0) S.v_qi = atomicAdd(A.qCounter, 1u);
        team_sync();
        const uint32_t slot = S.v_qi;
        if (slot >= (A.qlist ? (A.nListPtr ? *A.nListPtr : A.nList) : A.Q)) break;
        const uint32_t qi = A.qlist ? A.qlist[slot] : slot;
        const uint32_t qo = A.qoff[qi];
        const int L = (int)(A.qoff[qi + 1] - qo);
        const int P = L - A.pat.span + 1;
        if (P < 2) {
            if (tw == 0 && lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_OK; }
            team_sync();
            continue;
        }
        for (int i = tw * 32 + lane; i < BMW; i += TEAM * 32) { S.seen1[i] = 0; S.seen2[i] = 0; }
        if (tw == 0) {
            // ======== set-up by the first warp of the team ========
            uint32_t why = (L > PW_MAXL) ? ST_CAPACITY : ST_OK;   // status the query leaves this kernel with if not ST_OK
            uint32_t reason = (L > PW_MAXL) ? 1u : 0u;            // why it was handed over (statistics)
            int nRuns = 0;
            uint32_t K = 0;
            int minCorr = 0;
            if (lane == 0) { S.v_nHits = 0; S.v_nSurv = 0; S.v_nFin = 0; }
            if (why == ST_OK) {
                // query, composition bias (as in prefilter_kernel), window thresholds
                for (int i = lane; i < L; i += 32) S.q[i] = A.q[qo + i];
                __syncwarp();
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
                    cbuf[i] = d;
                    const float t4 = __fdiv_rn(d, 4.0f);
                    const double dd = (d < 0.0f) ? __dsub_rn((double)t4, 0.5) : __dadd_rn((double)t4, 0.5);
                    const signed char cc = (signed char)__double2float_rn(dd);
                    S.corr[i] = cc;
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
                        b = __fadd_rn(b, cbuf[w + A.pat.pos[k]]);
                    }
                    const short bias = (short)((b < 0.0f) ? __dsub_rn((double)b, 0.5) : __dadd_rn((double)b, 0.5));
                    const short thr = (short)max(A.kmerThr - (int)bias, 0);
                    int n1 = 0, x1 = 0, x2 = 0;
                    if (!hasX && w < P - 1) {   // quirk A: the last window is never counted
                        x1 = S.q[w + A.pat.pos[0]] + 20 * S.q[w + A.pat.pos[1]] + 400 * S.q[w + A.pat.pos[2]];
                        x2 = S.q[w + A.pat.pos[3]] + 20 * S.q[w + A.pat.pos[4]] + 400 * S.q[w + A.pat.pos[5]];
                        const int best2 = (int)(__ldg(A.row3 + (size_t)x2 * SPB_N3) >> 16) - SPB_SCORE_OFF;
                        const int cut1 = (int)thr - best2;
                        if (cut1 <= SPB_CNT3_W - SPB_SCORE_OFF - 1)
                            n1 = (int)__ldg(A.cnt3 + (size_t)x1 * SPB_CNT3_W + max(cut1, -SPB_SCORE_OFF) + SPB_SCORE_OFF);
                    }
                    S.thr[w] = thr;
                    n1buf[w] = (unsigned short)n1;
                    S.x1[w] = (unsigned short)x1;
                    S.x2[w] = (unsigned short)x2;
                }
                __syncwarp();
                uint32_t Rtot = 0;
                if (lane == 0) {
                    uint32_t acc = 0;
                    for (int w = 0; w < P; w++) { S.rowbase[w] = (uint16_t)min(acc, 65535u); acc += n1buf[w]; }
                    S.rowbase[P] = (uint16_t)min(acc, 65535u);
                    Rtot = acc;
                }
                Rtot = __shfl_sync(0xffffffffu, Rtot, 0);
                if (Rtot > 65535u) { why = ST_CAPACITY; reason = 2; }
                __syncwarp();
                // runs: maximal stretches of outer neighbours with equal score (hence equal inner count)
                // (two chunks of 32 outer neighbours per trip: their row3 / cnt3 reads are independent and go out together)
                for (uint32_t g0 = 0; g0 < Rtot && why == ST_OK; g0 += 64) {
                    uint32_t av[2], wv[2], n2v[2], cur[2], prv[2];
                    bool bnd[2];
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const uint32_t g = g0 + (uint32_t)(32 * v + lane);
                        av[v] = 0; wv[v] = 0; n2v[v] = 0; cur[v] = 0; prv[v] = 0;
                        if (g < Rtot) {
                            int lo = 0, hi = P - 1;   // largest w with rowbase[w] <= g
                            while (lo < hi) {
                                const int mid = (lo + hi + 1) >> 1;
                                if ((uint32_t)S.rowbase[mid] <= g) lo = mid; else hi = mid - 1;
                            }
                            wv[v] = (uint32_t)lo;
                            av[v] = g - (uint32_t)S.rowbase[lo];
                            const uint32_t *r1 = A.row3 + (size_t)S.x1[lo] * SPB_N3;
                            cur[v] = __ldg(r1 + av[v]);
                            prv[v] = av[v] ? __ldg(r1 + av[v] - 1) : 0u;
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const uint32_t g = g0 + (uint32_t)(32 * v + lane);
                        bnd[v] = false;
                        if (g < Rtot) {
                            const int s1 = (int)(cur[v] >> 16) - SPB_SCORE_OFF;
                            bnd[v] = (av[v] == 0) || ((int)(prv[v] >> 16) - SPB_SCORE_OFF != s1);
                            if (bnd[v]) {
                                const int cut2 = (int)S.thr[wv[v]] - s1;
                                if (cut2 <= SPB_CNT3_W - SPB_SCORE_OFF - 1)
                                    n2v[v] = (uint32_t)__ldg(A.cnt3 + (size_t)S.x2[wv[v]] * SPB_CNT3_W + max(cut2, -SPB_SCORE_OFF) + SPB_SCORE_OFF);
                            }
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const unsigned m = __ballot_sync(0xffffffffu, bnd[v]);
                        const int dst = nRuns + __popc(m & lt);
                        if (bnd[v] && dst < PW_RUNCAP) {
                            Run r;
                            r.a0 = (uint16_t)av[v]; r.n2 = (uint16_t)n2v[v]; r.cum = (uint16_t)(g0 + (uint32_t)(32 * v + lane)); r.w = (uint16_t)wv[v];
                            runs[dst] = r;
                        }
                        nRuns += __popc(m);
                    }
                    if (nRuns > PW_RUNCAP) { why = ST_CAPACITY; reason = 3; }
                }
                __syncwarp();
                // first k of every run: exclusive prefix of (rows in the run) x (inner count)
                for (int c0 = 0; c0 < nRuns && why == ST_OK; c0 += 32) {
                    const int i = c0 + lane;
                    uint32_t cntK = 0;
                    if (i < nRuns) {
                        const uint32_t gThis = runs[i].cum;
                        const uint32_t gNext = (i + 1 < nRuns) ? (uint32_t)runs[i + 1].cum : Rtot;
                        cntK = (gNext - gThis) * (uint32_t)runs[i].n2;
                    }
                    const uint32_t inc = (uint32_t)warp_incl_scan_ool((int)cntK);
                    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
                    __syncwarp();
                    if (K + tot > 65535u) { why = ST_CAPACITY; reason = 4; break; }
                    if (i < nRuns) runs[i].cum = (uint16_t)(K + inc - cntK);
                    K += tot;
                    __syncwarp();
                }
                if (SOLO && why == ST_OK && K > A.kCapSolo) why = ST_TEAM;
            }
            if (lane == 0) { S.v_K = K; S.v_nRuns = (uint32_t)nRuns; S.v_slow = why; S.v_why = reason; S.v_minCorr = minCorr; }
        }
        team_sync();
        PH_MARK(0);
        const uint32_t K = S.v_K;
        const int nRuns = (int)S.v_nRuns;
        // ---------------- one pass over the hit stream ----------------
        // Lane l of a pass owns the PW_U consecutive similar k-mers k0 + l * PW_U ...: one run search and one division give
        // (outer neighbour, inner neighbour) of the first, the others follow by stepping.  A table entry carries up to two
        // postings inline (95 % of the k-mers of a GenBank-phage-sized index occur in at most two ORFs), so most hits are
        // written straight from the entry: one scan, one reservation, no second memory access.  Longer lists are walked
        // warp-wide (pw_walk2).
        if (S.v_slow == ST_OK) {
            constexpr int HSH = SOLO ? 19 : 17;
            for (uint32_t k0 = (uint32_t)tw * (32 * PW_U); k0 < K; k0 += TEAM * 32 * PW_U) {
                const uint32_t kBeg = k0 + (uint32_t)lane * PW_U;
                ulonglong2 d[PW_U];
                uint32_t wPack = 0;
#pragma unroll
                for (int u = 0; u < PW_U; u++) d[u] = make_ulonglong2(0ull, 0ull);
                if (kBeg < K) {
                    int lo = 0, hi = nRuns - 1;   // largest r with runs[r].cum <= kBeg
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if ((uint32_t)runs[mid].cum <= kBeg) lo = mid; else hi = mid - 1;
                    }
                    int ri = lo;
                    Run r = runs[ri];
                    uint32_t nextCum = (ri + 1 < nRuns) ? (uint32_t)runs[ri + 1].cum : 0xFFFFFFFFu;
                    const uint32_t rel = kBeg - (uint32_t)r.cum;
                    uint32_t da = rel / (uint32_t)r.n2;
                    uint32_t b = rel - da * (uint32_t)r.n2;
                    const uint32_t *row1 = A.row3 + (size_t)S.x1[r.w] * SPB_N3 + (uint32_t)r.a0;
                    const uint32_t *row2 = A.row3 + (size_t)S.x2[r.w] * SPB_N3;
                    uint32_t i1 = __ldg(row1 + da) & 0xFFFFu;
#pragma unroll
                    for (int u = 0; u < PW_U; u++) {
                        const uint32_t k = kBeg + (uint32_t)u;
                        if (k < K) {
                            const uint32_t i2 = __ldg(row2 + b) & 0xFFFFu;
                            d[u] = __ldg(A.direct + (i1 + 8000u * i2));
                            wPack |= (uint32_t)r.w << (8 * u);
                            if (u + 1 < PW_U && k + 1 < K) {
                                if (++b == (uint32_t)r.n2) {
                                    b = 0;
                                    da++;
                                    if (k + 1 >= nextCum) {   // next run holding a k-mer (runs without inner neighbours share their cum)
                                        ri++;
                                        while (ri + 1 < nRuns && (uint32_t)runs[ri + 1].cum <= k + 1) ri++;
                                        r = runs[ri];
                                        nextCum = (ri + 1 < nRuns) ? (uint32_t)runs[ri + 1].cum : 0xFFFFFFFFu;
                                        da = 0;
                                        row1 = A.row3 + (size_t)S.x1[r.w] * SPB_N3 + (uint32_t)r.a0;
                                        row2 = A.row3 + (size_t)S.x2[r.w] * SPB_N3;
                                    }
                                    i1 = __ldg(row1 + da) & 0xFFFFu;
                                }
                            }
                        }
                    }
                }
                // inline postings
                uint32_t cnt = 0;
#pragma unroll
                for (int u = 0; u < PW_U; u++) cnt += (uint32_t)(d[u].x >> 63) + (uint32_t)((d[u].x >> 63) & (d[u].y >> 63));
                const uint32_t inc = (uint32_t)warp_incl_scan_ool((int)cnt);
                const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
                if (T) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&S.v_nHits, T);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    uint32_t dst = base + inc - cnt;
#pragma unroll
                    for (int s2 = 0; s2 < 2 * PW_U; s2++) {
                        const int u = s2 >> 1;
                        const uint64_t pe = (s2 & 1) ? d[u].y : d[u].x;
                        if ((pe >> 63) && (d[u].x >> 63)) {
                            const uint32_t id = (uint32_t)(pe >> 16);
                            const uint32_t w = (wPack >> (8 * u)) & 0xFFu;
                            const uint32_t dg = (w - (uint32_t)(pe & 0xFFFFull)) & 0xFFFFu;
                            if (dst < HCAP) __stcg(scratch + dst, ((uint64_t)id << 32) | ((uint64_t)(kBeg + (uint32_t)u) << 16) | dg);
                            dst++;
                            const uint32_t h = (id * 2654435761u) >> HSH;
                            const uint32_t bm = 1u << (h & 31);
                            const uint32_t old = atomicOr(&S.seen1[h >> 5], bm);
                            if ((old & bm) || (dg & 0xFFu) == 0u) atomicOr(&S.seen2[h >> 5], bm);
                        }
                    }
                }
                // posting lists of the k-mers in three or more targets, flattened over the warp.  (Deferring them to a per-warp
                // queue drained in bulk was measured in round 1h: no gain in the lookup phase, and the extra shared memory
                // of the TEAM = 1 launch cost more than it saved.)
                static_assert(PW_U == 2, "pw_walk2 serves exactly two k-mers per lane");
                {
                    const uint32_t len0 = (d[0].x >> 63) ? 0u : (uint32_t)(d[0].x >> 32);
                    const uint32_t len1 = (d[1].x >> 63) ? 0u : (uint32_t)(d[1].x >> 32);
                    if (__any_sync(0xffffffffu, (len0 | len1) != 0u))
                        pw_walk2(len0, (uint32_t)d[0].x, len1, (uint32_t)d[1].x, wPack, k0, A.post, &S.v_nHits, scratch, HCAP, S.seen1,
                                 S.seen2, HSH);
                }
            }
        }
        team_sync();
        PH_MARK(1);
        const uint32_t nHits = S.v_nHits;
        const bool live = (S.v_slow == ST_OK) && nHits <= HCAP;
        team_sync();
        if (S.v_slow == ST_OK && nHits > HCAP && tw == 0 && lane == 0) { S.v_slow = ST_CAPACITY; S.v_why = 5; }
        // ---------------- survivors: records of targets seen twice (or with 8-bit diagonal 0) ----------------
        // Pass A counts them per hash bucket (64 buckets of targets) and -- TEAM = 1 -- collects them optimistically into the
        // warp's sort buffer.  If they do not fit one sort, pass B scatters them bucket by bucket behind the hit records and the
        // buckets are merged into G contiguous groups of at most PW_SCAP records, each processed independently by one warp.
        uint32_t *const cnt64 = S.seen1;            // dead after the gather: [0,64) counts, [64,128) starts, [128,192) fill cursors
        if (live) for (int i = tw * 32 + lane; i < 192; i += TEAM * 32) cnt64[i] = 0;
        team_sync();
        int nS0 = 0;
        if (live) {
            for (uint32_t i0 = (uint32_t)tw * 128; i0 < nHits; i0 += TEAM * 128) {
                uint64_t r[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)(u * 32 + lane);
                    r[u] = (i < nHits) ? __ldcg(scratch + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)(u * 32 + lane);
                    const uint32_t id = (uint32_t)(r[u] >> 32);
                    const uint32_t h = (id * 2654435761u) >> (SOLO ? 19 : 17);
                    const bool keep = (i < nHits) && ((S.seen2[h >> 5] >> (h & 31)) & 1u);
                    if (keep) atomicAdd(&cnt64[(id * 0x85EBCA6Bu) >> 26], 1u);
                    if (SOLO) {
                        const unsigned m = __ballot_sync(0xffffffffu, keep);
                        const int dst = nS0 + __popc(m & lt);
                        if (keep && dst < PW_SCAP) W.surv[dst] = r[u];
                        nS0 += __popc(m);
                    }
                }
            }
        }
        team_sync();
        // group count: the smallest power of two whose merged buckets all fit one sort (every warp computes the same)
        uint32_t nSurv = 0, G = 1;
        if (live) {
            const uint32_t c0 = cnt64[2 * lane], c1 = cnt64[2 * lane + 1];
            // level l uses 64 >> l groups; the largest merged bucket grows with l, so the last level that fits wins
            uint32_t sum = c0 + c1, mxFit = 0;
            int lvl = -1;
#pragma unroll 1
            for (int l = 0; l < 7; l++) {
                if (l >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1 << (l - 2));
                uint32_t m = (l == 0) ? max(c0, c1) : sum;
#pragma unroll 1
                for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
                if (m <= PW_SCAP) { lvl = l; mxFit = m; }
                if (l == 6) nSurv = m;
            }
            (void)mxFit;
            G = 64u >> max(lvl, 0);
            if (lvl < 0 && tw == 0 && lane == 0) { S.v_slow = ST_CAPACITY; S.v_why = 6; }   // one bucket alone overflows a sort
            if (tw == 0) {   // bucket starts
                uint32_t e0 = (uint32_t)warp_incl_scan_ool((int)(c0 + c1)) - (c0 + c1);
                cnt64[64 + 2 * lane] = e0;
                cnt64[64 + 2 * lane + 1] = e0 + c0;
            }
        }
        team_sync();
        const bool scatter = live && (S.v_slow == ST_OK) && (!SOLO || G > 1);
        uint64_t *const grouped = scratch + HCAP;
        if (scatter) {
            for (uint32_t i0 = (uint32_t)tw * 128; i0 < nHits; i0 += TEAM * 128) {
                uint64_t r[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)(u * 32 + lane);
                    r[u] = (i < nHits) ? __ldcg(scratch + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + (uint32_t)(u * 32 + lane);
                    const uint32_t id = (uint32_t)(r[u] >> 32);
                    const uint32_t h = (id * 2654435761u) >> (SOLO ? 19 : 17);
                    if ((i < nHits) && ((S.seen2[h >> 5] >> (h & 31)) & 1u)) {
                        const uint32_t bk = (id * 0x85EBCA6Bu) >> 26;
                        __stcg(grouped + cnt64[64 + bk] + atomicAdd(&cnt64[128 + bk], 1u), r[u]);
                    }
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
                    for (int i = lane; i < nS; i += 32) W.surv[i] = __ldcg(grouped + lo + i);
                }
                // ---- sort by (target, stream order) ----
                {
                    const int np2 = next_pow2(max(nS, 2));
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
                if (lane == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = why; A.need[qi] = reason; }
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
                warp_bitonic_u64(W.surv, np2);
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
            uint32_t *const hist = S.seen2;       // dead after the survivor passes
            const uint32_t nc = S.v_nFin;
            const uint32_t maxHits = (uint32_t)A.maxHits;
            const int tid = threadIdx.x;
            if (tid == 0) {
                S.v_nSurv = 0;                    // reused: selected candidates
                if (S.v_slow == ST_OK && nc > FCAP) { S.v_slow = ST_CAPACITY; S.v_why = 7; }
            }
            for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const bool ok0 = (S.v_slow == ST_OK);
            const bool trunc = ok0 && nc >= maxHits;
            if (trunc) {
                for (uint32_t k = tid; k < nc; k += blockDim.x) atomicAdd(&hist[255u - (uint32_t)(__ldcg(finList + 2 * k + 1) >> 55)], 1u);
                __syncthreads();
                if (tid == 0) {     // computeScoreThreshold (QueryMatcher.h:199-209)
                    uint32_t found = 0, thr = 0;
                    for (int sc = 255; sc >= 1; sc--) {
                        found += hist[sc];
                        if (found >= maxHits) { thr = (uint32_t)sc; break; }
                    }
                    const uint32_t thrD = max(thr, (uint32_t)A.minDiag);
                    S.v_nRuns = thrD;             // reused: cut score
                    // all remaining candidates saturated -> rescoreHits (QueryMatcher.cpp:486-505) lives in prefilter_kernel
                    if ((int)thrD >= capScore) { S.v_slow = ST_CAPACITY; S.v_why = 8; }
                }
                __syncthreads();
            }
            const bool ok1 = (S.v_slow == ST_OK);
            uint32_t ns = nc;                     // rows to output
            if (ok1) {
                if (nc > maxHits) {
                    const uint32_t thrD = S.v_nRuns;
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
                                   uint32_t *__restrict__ list, unsigned int *__restrict__ ctr, uint32_t listStatus) {
    // ctr: [0] entries of `list` (queries whose status is listStatus), [1] maxNeed, [2] arena overflow,
    //      [3] layout error (query index + 1), [4] unprocessed
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
    } else if (s != ST_OK && !(listStatus == ST_TEAM && s == ST_CAPACITY)) {
        atomicMax(&ctr[4], i + 1);
    }
}

__global__ void similar_kmers_test_kernel(const uint32_t *row3, const uint16_t *cnt3, const uint8_t *k6, int thr, uint32_t *out,
                                          uint32_t cap, uint32_t *n) {
    __shared__ int s_excl[32];
    __shared__ uint32_t s_i1[32];
    const int x1 = k6[0] + 20 * k6[1] + 400 * k6[2];
    const int x2 = k6[3] + 20 * k6[4] + 400 * k6[5];
    const int best2 = (int)(row3[(size_t)x2 * SPB_N3] >> 16) - SPB_SCORE_OFF;
    const int cut1 = thr - best2;
    int n1 = 0;
    if (cut1 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) n1 = cnt3[(size_t)x1 * SPB_CNT3_W + max(cut1, -SPB_SCORE_OFF) + SPB_SCORE_OFF];
    uint32_t base = 0;
    for (int a0 = 0; a0 < n1; a0 += 32) {
        uint32_t tot = warp_unit_kmers(row3, cnt3, x1, x2, thr, n1, a0, s_excl, s_i1, [&](uint32_t kmer, uint32_t a, uint32_t b, bool valid) {
            // rank within the unit: position in (a, b) order == base + (exclusive count of a) + b
            if (valid) {
                const uint32_t r = base + (uint32_t)s_excl[a - a0] + b;
                if (r < cap) out[r] = kmer;
            }
        });
        base += tot;
    }
    if (threadIdx.x == 0) *n = base;
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
    if (e == cudaSuccess && *n) e = cudaMemcpy(out, dout.p, (size_t)(*n < cap ? *n : cap) * 4, cudaMemcpyDeviceToHost);
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
    A.koff = c->d_koff.as<uint32_t>();
    A.post = c->d_post.as<uint64_t>();
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

    // scratch: s_a start u64[Q], s_b count u32[Q], s_c status u32[Q], s_d need u32[Q], s_e slow list u32[Q], s_f counters, s_g arena
    if (c->s_a.ensure((size_t)Q * 8) || c->s_b.ensure((size_t)Q * 4) || c->s_c.ensure((size_t)Q * 4) || c->s_d.ensure((size_t)Q * 4) ||
        c->s_e.ensure((size_t)Q * 4) || c->s_f.ensure(512)) return 1;
    uint64_t arenaCap = std::min<uint64_t>((uint64_t)Q * (uint64_t)A.maxHits, std::max<uint64_t>((uint64_t)Q * 24, 1u << 20));
    float kernelMs = 0;
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
        // stage 1: team kernels.  TEAM = 1 over the whole batch (persistent grid, dynamic query scheduler); the queries it
        // flags ST_TEAM (long similar-k-mer lists) are listed on the device and taken by the TEAM = 8 launch, which reads
        // the list length from device memory -- no host round trip in between.
        int soloPerSm = 8;
        if (const char *e = getenv("SPB_SOLO_CTAS")) soloPerSm = std::max(1, std::min(8, atoi(e)));   // tuning aid (L1 carve-out experiments)
        if (const char *e = getenv("SPB_SOLO_CARVEOUT"))    // percent of the 228 KB given to shared memory (the rest is L1)
            cudaFuncSetAttribute(prefilter_team_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
        int teamPerSm = 4;
        if (const char *e = getenv("SPB_TEAM_CTAS")) teamPerSm = std::max(1, std::min(4, atoi(e)));
        if (const char *e = getenv("SPB_TEAM_CARVEOUT"))
            cudaFuncSetAttribute(prefilter_team_kernel<PW_TEAM>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
        const int soloBlocks = c->n_sm * soloPerSm, teamBlocks = c->n_sm * teamPerSm;
        A.direct = c->d_direct.as<ulonglong2>();
        uint32_t hitCapSolo = 8192, hitCapTeam = 32768;
        A.kCapSolo = 4096;
        if (const char *e = getenv("SPB_KCAP_SOLO")) A.kCapSolo = (uint32_t)atoi(e);        // tuning aids (tools/perf_probe.py)
        if (const char *e = getenv("SPB_HITCAP_SOLO")) hitCapSolo = (uint32_t)atoi(e);
        const size_t sliceSolo = 2 * (size_t)hitCapSolo + 2 * PW_FCAP_SOLO, sliceTeam = 2 * (size_t)hitCapTeam + 2 * PW_FCAP_TEAM;
        if (c->s_j.ensure(std::max((size_t)soloBlocks * PW_WARPS * sliceSolo, (size_t)teamBlocks * sliceTeam) * sizeof(uint64_t)) ||
            c->s_k.ensure((size_t)Q * 4 + 64)) return 1;
        A.hitCap = hitCapSolo;
        A.hitScratch = c->s_j.as<uint64_t>();
        unsigned int *d_qc = reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 14);   // [0] solo, [1] team scheduler
        unsigned int *d_ctrT = reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 24); // counters of the ST_TEAM scan
        A.qCounter = d_qc;
        A.qlist = c->d_qorder.as<uint32_t>();     // longest queries first: their k-mer lists are the long ones (tail of the launch)
        A.nListPtr = nullptr;
        A.nList = Q;
        SPB_CUDA(cudaEventRecord(c->ev0, st));
        prefilter_team_kernel<1><<<soloBlocks, PW_WARPS * 32, 0, st>>>(A);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_k.as<uint32_t>(), d_ctrT, ST_TEAM);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        A.qCounter = d_qc + 1;
        A.qlist = c->s_k.as<uint32_t>();
        A.nListPtr = d_ctrT;
        A.hitCap = hitCapTeam;
        prefilter_team_kernel<PW_TEAM><<<teamBlocks, PW_TEAM * 32, 0, st>>>(A);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        A.qlist = nullptr;
        A.nListPtr = nullptr;
        SPB_CUDA(cudaEventRecord(c->ev1, st));
        status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_e.as<uint32_t>(), d_ctr, ST_CAPACITY);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, 32, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        kernelMs += ms;
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
        SPB_CHECK(h_ctr[3] == 0, "spb_prefilter: query %u is longer than the supported %d residues", h_ctr[3] - 1, PF_MAXL);
        SPB_CHECK(h_ctr[4] == 0, "spb_prefilter: query %u was not processed", h_ctr[4] - 1);
        bool arenaOverflow = h_ctr[2] != 0;
        uint32_t nList = h_ctr[0], maxNeed = 0, capG = 2048;
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
            status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_i.as<uint32_t>(), d_ctr, ST_CAPACITY);
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
            status_scan_kernel<<<(Q + 255) / 256, 256, 0, st>>>(A.status, A.need, Q, c->s_i.as<uint32_t>(), d_ctr, ST_CAPACITY);
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
    if (getenv("SPB_DEBUG")) {
        unsigned int nTeam = 0;
        cudaMemcpy(&nTeam, reinterpret_cast<unsigned int *>(c->s_f.as<unsigned long long>() + 24), 4, cudaMemcpyDeviceToHost);
        fprintf(stderr, "spb_prefilter: %u queries, %u taken by the 8-warp team launch, %llu handed to the CTA kernel "
                "(long query %u, rows %u, runs %u, k-mers %u, hits %u, bucket %u, max-seqs %u)\n", Q, nTeam,
                (unsigned long long)nSlow, dbgWhy[1], dbgWhy[2], dbgWhy[3], dbgWhy[4], dbgWhy[5], dbgWhy[6], dbgWhy[7]);
    }
    spb_prefilter_result &R = c->pref_stats;
    R.n_hits = nHits;
    R.n_kmers = h_cursor[1];
    R.n_db_matches = h_cursor[2];
    R.n_events = h_cursor[3];
    R.n_slow_path = nSlow;
    R.kernel_ms = kernelMs;
    R.total_ms = totalMs;
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
        static bool attr_done = false;
        if (!attr_done) {
            SPB_CUDA(cudaFuncSetAttribute(merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
            attr_done = true;
        }
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
