// prefilter.cu -- k-mer prefilter for a batch of query frames, one CTA per query.
//
// Replaces QueryMatcher::matchQuery and callees for AA/AA, --diag-score 1
// (M/prefiltering/QueryMatcher.cpp:84-307,324-381,459-505; KmerGenerator.cpp:107-216;
//  CacheFriendlyOperations.cpp:146-235,311-341; UngappedAlignment.cpp:27-51,216-233,295-403).
//
// The reference consumes a sequential hit stream ordered by (window, similar-k-mer rank, posting order) and runs
// order-dependent per-target byte state machines over it.  Here every hit is tagged with (id, window, rank),
// the CTA sorts its hits by that key in shared memory, and every reference rule becomes a comparison between
// neighbours of the sorted list:
//   double hit      : hit i is an event iff its 8-bit diagonal equals that of the previous hit of the same target,
//                     or -- first hit of a target -- equals 0 (quirk B: the byte array starts zeroed)
//   event de-dup    : an event is kept iff the previous EVENT of the same target has a different 8-bit diagonal
//   keepMaxElement  : per target the first kept event (stream order) reaching the target's best capped score
//   truncation      : order key (score desc, id & (BIN-1), window, rank, id) == order of the reference's two stable
//                     counting sorts over the bin-major event list
// Windows w >= P-1 (the last window) never contribute (quirk A, CacheFriendlyOperations.cpp:43).
#include <algorithm>

#include "common.cuh"

namespace spb {

struct Pattern6 { int span; int pos[6]; };

struct PrefArgs {
    const uint8_t *q;
    const uint32_t *qoff;
    uint32_t Q;
    const uint32_t *qlist;   // optional list of query indices (slow path), else NULL
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
    int rankBits, wBits;     // record layout: id | w | rank | pos(16)
    // storage for the global-memory variant
    uint64_t *gScratch;      // [gridDim.x][cap]
    uint32_t cap;            // record capacity per CTA
    // outputs
    spb_hit *arena;
    unsigned long long *cursor;  // [0] arena cursor, [1] kmers, [2] dbMatches, [3] events, [4] gather bytes
    uint64_t arenaCap;
    uint64_t *outStart;      // per query
    uint32_t *outCount;      // per query
    uint32_t *status;        // per query: 0 ok, 1 exceeds this variant's capacity, 2 arena overflow, 3 record layout overflow
    uint32_t *need;          // per query: hits needed when status == 1
};

#define ST_OK 0u
#define ST_CAPACITY 1u
#define ST_ARENA 2u
#define ST_LAYOUT 3u

// Enumerate the similar k-mers of one window with one warp, in the order of KmerGenerator::calculateArrayProduct:
// outer loop over the neighbours of the first 3-mer (row order) while score >= thr - best2, inner loop over the
// neighbours of the second 3-mer while score >= thr - s1.  f(kmer, rank, valid) is called by all 32 lanes.
template <class F>
__device__ __forceinline__ uint32_t warp_similar_kmers(const uint32_t *__restrict__ row3, const uint16_t *__restrict__ cnt3,
                                                       int x1, int x2, int thr, int *s_excl, uint32_t *s_i1, F f) {
    const int lane = threadIdx.x & 31;
    const uint32_t *r1 = row3 + (size_t)x1 * SPB_N3;
    const uint32_t *r2 = row3 + (size_t)x2 * SPB_N3;
    const uint16_t *c1 = cnt3 + (size_t)x1 * SPB_CNT3_W;
    const uint16_t *c2 = cnt3 + (size_t)x2 * SPB_CNT3_W;
    const int best2 = (int)(__ldg(r2) >> 16) - SPB_SCORE_OFF;
    int cut1 = thr - best2;
    cut1 = max(-SPB_SCORE_OFF, min(cut1, SPB_CNT3_W - SPB_SCORE_OFF - 1));
    const int n1 = (thr - best2 > SPB_CNT3_W - SPB_SCORE_OFF - 1) ? 0 : (int)__ldg(c1 + cut1 + SPB_SCORE_OFF);
    uint32_t rankBase = 0;
    for (int a0 = 0; a0 < n1; a0 += 32) {
        const int a = a0 + lane;
        uint32_t e1 = 0;
        int n2 = 0;
        if (a < n1) {
            e1 = __ldg(r1 + a);
            int s1 = (int)(e1 >> 16) - SPB_SCORE_OFF;
            int cut2 = thr - s1;
            if (cut2 <= SPB_CNT3_W - SPB_SCORE_OFF - 1) {
                cut2 = max(cut2, -SPB_SCORE_OFF);
                n2 = (int)__ldg(c2 + cut2 + SPB_SCORE_OFF);
            }
        }
        int inc = warp_incl_scan(n2);
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        __syncwarp();
        s_excl[lane] = inc - n2;
        s_i1[lane] = e1 & 0xFFFFu;
        __syncwarp();
        for (int p0 = 0; p0 < total; p0 += 32) {
            const int p = p0 + lane;
            const bool valid = p < total;
            uint32_t kmer = 0;
            if (valid) {
                // largest j with s_excl[j] <= p
                int lo = 0, hi = 31;
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    int mid = (lo + hi + 1) >> 1;
                    if (s_excl[mid] <= p) lo = mid; else hi = mid - 1;
                }
                const int b = p - s_excl[lo];
                kmer = s_i1[lo] + 8000u * (__ldg(r2 + b) & 0xFFFFu);
            }
            f(kmer, rankBase + (uint32_t)p, valid);
        }
        rankBase += (uint32_t)total;
        __syncwarp();
    }
    return rankBase;
}

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
    if (L >= 32768 || tLen >= 32768) {
        int totalMax = 0;
        for (uint32_t d = 1; d <= 1 + tLen / 32768; d++) {
            int real = (int)(-(long long)d * 65536 + (long long)diag16);
            totalMax = max(totalMax, ungapped_diag(t, tLen, q, L, corr, m40u, real));
        }
        for (uint32_t d = 0; d <= (uint32_t)L / 65536; d++) {
            int real = (int)(d * 65536 + diag16);
            totalMax = max(totalMax, ungapped_diag(t, tLen, q, L, corr, m40u, real));
        }
        return totalMax;
    }
    return ungapped_diag(t, tLen, q, L, corr, m40u, (int)(short)diag16);
}

#define PF_THREADS 128
#define PF_WARPS (PF_THREADS / 32)
#define PF_MAXL_SMEM 512     // queries up to this length keep their per-position arrays in shared memory

// In-place stable compaction of a[0..n) (and optionally b[0..n)) keeping elements with flag; the flag functor sees
// (index, element, previous element or ~0 for index 0).  Returns the new length (uniform across the block).
template <class Flag>
__device__ __forceinline__ int block_compact_prev(uint64_t *a, int n, int *wsum, Flag flag) {
    __shared__ uint64_t s_carry;
    int outN = 0;
    if (threadIdx.x == 0) s_carry = ~0ull;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        uint64_t r = 0, rp = ~0ull;
        int keep = 0;
        if (i < n) {
            r = a[i];
            rp = (threadIdx.x == 0) ? s_carry : a[i - 1];
            keep = flag(i, r, rp) ? 1 : 0;
        }
        int total;
        int ex = block_excl_scan(keep, wsum, &total);  // contains __syncthreads before any write below
        const int last = min(n, c0 + (int)blockDim.x) - 1;
        if (i == last) s_carry = r;
        if (keep) a[outN + ex] = r;
        outN += total;
        __syncthreads();
    }
    return outN;
}

template <bool GLOBAL>
__global__ void __launch_bounds__(PF_THREADS) prefilter_kernel(PrefArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_q[PF_MAXL_SMEM];
    __shared__ float s_cb[PF_MAXL_SMEM];
    __shared__ signed char s_corr[PF_MAXL_SMEM];
    __shared__ short s_thr[PF_MAXL_SMEM];       // per window: similar-k-mer threshold, -1 = window contains X
    __shared__ signed char s_m40u[SPB_ALPHABET * SPB_ALPHABET];
    __shared__ int s_wsum[33];
    __shared__ int s_excl[PF_WARPS][32];
    __shared__ uint32_t s_i1[PF_WARPS][32];
    __shared__ int s_hist[256];
    __shared__ unsigned int s_nHits, s_flags;
    __shared__ int s_minCorr, s_thrD;
    __shared__ unsigned long long s_base, s_stat[3];

    const uint32_t qi = A.qlist ? A.qlist[blockIdx.x] : blockIdx.x;
    if (qi >= A.Q) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t qo = A.qoff[qi];
    const int L = (int)(A.qoff[qi + 1] - qo);
    const int P = L - A.pat.span + 1;
    if (P < 1) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_OK; }
        return;
    }
    if (L > PF_MAXL_SMEM || P > (1 << A.wBits)) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_LAYOUT; }
        return;
    }
    uint64_t *rec = GLOBAL ? (A.gScratch + (size_t)blockIdx.x * A.cap) : reinterpret_cast<uint64_t *>(smem_raw);
    const uint32_t CAP = A.cap;
    const int rankBits = A.rankBits, wShift = 16 + A.rankBits, idShift = 16 + A.rankBits + A.wBits;
    const uint64_t rankMask = (1ull << rankBits) - 1, wMask = (1ull << A.wBits) - 1;

    // ---------------- phase 0/1: query, composition bias, thresholds ----------------
    for (int i = tid; i < SPB_ALPHABET * SPB_ALPHABET; i += PF_THREADS) s_m40u[i] = (signed char)A.mat->ungapped[i];
    for (int i = tid; i < L; i += PF_THREADS) s_q[i] = A.q[qo + i];
    if (tid == 0) { s_nHits = 0; s_flags = 0; s_minCorr = 0; s_stat[0] = s_stat[1] = s_stat[2] = 0; }
    for (int i = tid; i < 256; i += PF_THREADS) s_hist[i] = 0;
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
        float t4 = __fdiv_rn(d, 4.0f);
        double dd = (d < 0.0f) ? __dsub_rn((double)t4, 0.5) : __dadd_rn((double)t4, 0.5);
        signed char cc = (signed char)__double2float_rn(dd);
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
        short bias = (short)((b < 0.0f) ? __dsub_rn((double)b, 0.5) : __dadd_rn((double)b, 0.5));
        short thr = (short)max(A.kmerThr - (int)bias, 0);
        s_thr[w] = hasX ? (short)-1 : thr;
    }
    __syncthreads();
    const int ungBias = abs(A.minSubUngapped) + abs(min(s_minCorr, 0));
    const int capScore = 255 - ungBias;

    // ---------------- phase 2: similar k-mers -> index postings -> tagged hits ----------------
    unsigned long long nk = 0, nm = 0;
    for (int w = wid; w < P - 1; w += PF_WARPS) {   // quirk A: the last window is never counted
        const int thr = s_thr[w];
        if (thr < 0) continue;
        const int x1 = s_q[w + A.pat.pos[0]] + 20 * s_q[w + A.pat.pos[1]] + 400 * s_q[w + A.pat.pos[2]];
        const int x2 = s_q[w + A.pat.pos[3]] + 20 * s_q[w + A.pat.pos[4]] + 400 * s_q[w + A.pat.pos[5]];
        uint32_t nKmers = warp_similar_kmers(A.row3, A.cnt3, x1, x2, thr, s_excl[wid], s_i1[wid],
            [&](uint32_t kmer, uint32_t rank, bool valid) {
                uint32_t o0 = 0, len = 0;
                if (valid) {
                    o0 = __ldg(A.koff + kmer);
                    len = __ldg(A.koff + kmer + 1) - o0;
                }
                int inc = warp_incl_scan((int)len);
                const int tot = __shfl_sync(0xffffffffu, inc, 31);
                if (tot == 0) return;
                unsigned int base = 0;
                if (lane == 31) base = atomicAdd(&s_nHits, (unsigned int)tot);
                base = __shfl_sync(0xffffffffu, base, 31);
                nm += len;
                if (len) {
                    if ((uint64_t)rank > rankMask) atomicOr(&s_flags, 1u);
                    unsigned int dst = base + (unsigned int)(inc - (int)len);
                    for (uint32_t z = 0; z < len; z++, dst++) {
                        if (dst < CAP) {
                            uint64_t pe = __ldg(A.post + o0 + z);   // id << 16 | pos
                            rec[dst] = ((pe >> 16) << idShift) | ((uint64_t)w << wShift) |
                                       (((uint64_t)rank & rankMask) << 16) | (pe & 0xFFFFull);
                        }
                    }
                }
            });
        if (lane == 0) nk += nKmers;
    }
    for (int d = 16; d > 0; d >>= 1) nm += __shfl_down_sync(0xffffffffu, nm, d);
    if (lane == 0) { atomicAdd(&s_stat[0], nk); atomicAdd(&s_stat[1], nm); }
    __syncthreads();
    int n = (int)s_nHits;
    if (s_flags & 1u) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_LAYOUT; }
        return;
    }
    if ((uint32_t)n > CAP) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY; A.need[qi] = (uint32_t)n; }
        return;
    }

    // ---------------- phase 3: sort hits by (id, window, rank) ----------------
    {
        const int np2 = next_pow2(max(n, 2));
        for (int i = n + tid; i < np2; i += PF_THREADS) rec[i] = ~0ull;
        __syncthreads();
        block_bitonic_sort(rec, np2, [](uint64_t a, uint64_t b) { return a < b; });
    }

    // ---------------- phase 4: double-hit events, then de-duplication of consecutive equal events ----------------
    auto diag16 = [&](uint64_t r) -> uint32_t { return (uint32_t)(((r >> wShift) & wMask) - (r & 0xFFFFull)) & 0xFFFFu; };
    int ne = block_compact_prev(rec, n, s_wsum, [&](int i, uint64_t r, uint64_t rp) {
        const uint32_t d8 = diag16(r) & 0xFFu;
        const bool same = (i > 0) && ((rp >> idShift) == (r >> idShift));
        return same ? (d8 == (diag16(rp) & 0xFFu)) : (d8 == 0u);
    });
    int nkept = block_compact_prev(rec, ne, s_wsum, [&](int i, uint64_t r, uint64_t rp) {
        const bool same = (i > 0) && ((rp >> idShift) == (r >> idShift));
        return !same || ((diag16(r) & 0xFFu) != (diag16(rp) & 0xFFu));
    });
    if (tid == 0) s_stat[2] = (unsigned long long)nkept;

    // later stages need: events A[0..EV), final keys B[0..EV), cnt u8[EV], ord u16[EV] inside the record buffer
    const uint32_t EV = CAP / 4;
    if ((uint32_t)nkept > EV || (uint32_t)next_pow2(max(A.maxHits, 2)) > EV) {
        if (tid == 0) { A.outStart[qi] = 0; A.outCount[qi] = 0; A.status[qi] = ST_CAPACITY; A.need[qi] = max((uint32_t)n, 4u * (uint32_t)max(nkept, 2 * A.maxHits)); }
        return;
    }
    uint64_t *evA = rec;
    uint64_t *finB = rec + EV;
    uint8_t *cnt = reinterpret_cast<uint8_t *>(rec + 2 * (size_t)EV);
    uint16_t *ord16 = reinterpret_cast<uint16_t *>(rec + 2 * (size_t)EV + (EV + 7) / 8);
    uint32_t *ord32 = reinterpret_cast<uint32_t *>(rec + 2 * (size_t)EV + (EV + 7) / 8);
    const bool wideOrd = EV > 65536u;

    // ---------------- phase 5: ungapped diagonal scores (capped like the uint8 counter) ----------------
    for (int i = tid; i < nkept; i += PF_THREADS) {
        const uint64_t r = evA[i];
        int s = ungapped_u16(A.tres, A.toff, (uint32_t)(r >> idShift), s_q, L, s_corr, s_m40u, diag16(r));
        cnt[i] = (uint8_t)min(s, capScore);
    }
    __syncthreads();

    // ---------------- phase 6: score >= minDiag and, per target, the first event reaching the target's maximum ----------------
    int nc = 0;
    for (int c0 = 0; c0 < nkept; c0 += PF_THREADS) {
        const int i = c0 + tid;
        int keep = 0;
        uint64_t r = 0;
        uint8_t best = 0;
        if (i < nkept) {
            r = evA[i];
            const uint64_t id = r >> idShift;
            const bool head = (i == 0) || ((evA[i - 1] >> idShift) != id);
            if (head) {
                int bi = i;
                best = cnt[i];
                for (int j = i + 1; j < nkept && (evA[j] >> idShift) == id; j++)
                    if (cnt[j] > best) { best = cnt[j]; bi = j; }
                if ((int)best >= A.minDiag) { keep = 1; r = evA[bi]; }
            }
        }
        int total;
        int ex = block_excl_scan(keep, s_wsum, &total);
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
        // rescoreHits (QueryMatcher.cpp:486-505): all remaining candidates are saturated; rescale exact scores to 0..255
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
            uint64_t r = 0;
            uint8_t nc8 = 0;
            if (i < nc && (int)cnt[i] >= capScore) {
                r = evA[i];
                unsigned int nsx = (unsigned int)ungapped_u16(A.tres, A.toff, (uint32_t)(r >> idShift), s_q, L, s_corr, s_m40u, diag16(r));
                nsx -= (unsigned int)capScore;
                float sc = (float)min(nsx, 65535u);
                nc8 = (uint8_t)__dadd_rn((double)__fmul_rn(__fdiv_rn(sc, fltMax), 255.0f), 0.5);
                keep = 1;
            }
            int total;
            int ex = block_excl_scan(keep, s_wsum, &total);
            if (keep) { evA[m + ex] = r; cnt[m + ex] = nc8; }
            m += total;
            __syncthreads();
        }
        nc = m;
    }
    if (nc > maxHits) {
        // order of the reference's stable counting sorts: score desc, bin, (window, rank), id
        const uint32_t binMask = A.binMask, idBase = A.idBase;
        auto less = [&](uint32_t a, uint32_t b) {
            if (a == 0xFFFFFFFFu || b == 0xFFFFFFFFu) return a != 0xFFFFFFFFu && b == 0xFFFFFFFFu;
            const uint8_t ca = cnt[a], cb = cnt[b];
            if (ca != cb) return ca > cb;
            const uint64_t ra = evA[a], rb = evA[b];
            const uint32_t ba = ((uint32_t)(ra >> idShift) + idBase) & binMask, bb = ((uint32_t)(rb >> idShift) + idBase) & binMask;
            if (ba != bb) return ba < bb;
            const uint64_t oa = (ra >> 16) & ((1ull << (idShift - 16)) - 1), ob = (rb >> 16) & ((1ull << (idShift - 16)) - 1);
            if (oa != ob) return oa < ob;
            return (ra >> idShift) < (rb >> idShift);
        };
        const int np2 = next_pow2(nc);
        if (wideOrd) {
            for (int i = tid; i < np2; i += PF_THREADS) ord32[i] = i < nc ? (uint32_t)i : 0xFFFFFFFFu;
            __syncthreads();
            block_bitonic_sort(ord32, np2, less);
        } else {
            for (int i = tid; i < np2; i += PF_THREADS) ord16[i] = i < nc ? (uint16_t)i : (uint16_t)0xFFFF;
            __syncthreads();
            block_bitonic_sort(ord16, np2, [&](uint16_t a, uint16_t b) {
                return less(a == 0xFFFF ? 0xFFFFFFFFu : (uint32_t)a, b == 0xFFFF ? 0xFFFFFFFFu : (uint32_t)b);
            });
        }
        ns = maxHits;
    } else {
        ns = nc;
    }

    // ---------------- phase 8: exact scores, final order (score desc, id asc), output ----------------
    for (int k = tid; k < ns; k += PF_THREADS) {
        const int i = (nc > maxHits) ? (wideOrd ? (int)ord32[k] : (int)ord16[k]) : k;
        const uint64_t r = evA[i];
        const uint32_t id = (uint32_t)(r >> idShift);
        int sc = cnt[i];
        if (rescale) {
            sc = capScore + (int)((unsigned int)cnt[i] * (unsigned int)maxSelf / 255u);
        } else if (sc >= capScore) {
            sc = ungapped_u16(A.tres, A.toff, id, s_q, L, s_corr, s_m40u, diag16(r));   // getResult, QueryMatcher.cpp:371-374
        }
        finB[k] = ((uint64_t)(0xFFFFFu - (uint32_t)min(sc, 0xFFFFF)) << 44) | ((uint64_t)id << 16) | diag16(r);
    }
    {
        const int np2 = next_pow2(max(ns, 2));
        for (int i = ns + tid; i < np2; i += PF_THREADS) finB[i] = ~0ull;
        __syncthreads();
        block_bitonic_sort(finB, np2, [](uint64_t a, uint64_t b) { return a < b; });
    }
    if (tid == 0) {
        unsigned long long base = atomicAdd(A.cursor, (unsigned long long)ns);
        s_base = base;
        atomicAdd(A.cursor + 1, s_stat[0]);
        atomicAdd(A.cursor + 2, s_stat[1]);
        atomicAdd(A.cursor + 3, s_stat[2]);
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
        h.id = (uint32_t)((f >> 16) & 0xFFFFFFFull);
        h.score = (int32_t)(0xFFFFFu - (uint32_t)(f >> 44));
        h.diag = (int32_t)(short)(uint16_t)(f & 0xFFFFull);
        A.arena[base + k] = h;
    }
    if (tid == 0) { A.outStart[qi] = base; A.outCount[qi] = (uint32_t)ns; A.status[qi] = ST_OK; }
}

// gather per-query slices of the arena into query order
__global__ void gather_hits_kernel(const spb_hit *__restrict__ arena, const uint64_t *__restrict__ start, const uint64_t *__restrict__ outOff,
                                   uint32_t Q, spb_hit *__restrict__ out) {
    const uint32_t qi = blockIdx.x;
    if (qi >= Q) return;
    const uint64_t s = start[qi], o = outOff[qi];
    const uint32_t n = (uint32_t)(outOff[qi + 1] - o);
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) out[o + k] = arena[s + k];
}

__global__ void similar_kmers_test_kernel(const uint32_t *row3, const uint16_t *cnt3, const uint8_t *k6, int thr, uint32_t *out,
                                          uint32_t cap, uint32_t *n) {
    __shared__ int s_excl[32];
    __shared__ uint32_t s_i1[32];
    const int x1 = k6[0] + 20 * k6[1] + 400 * k6[2];
    const int x2 = k6[3] + 20 * k6[4] + 400 * k6[5];
    uint32_t tot = warp_similar_kmers(row3, cnt3, x1, x2, thr, s_excl, s_i1, [&](uint32_t kmer, uint32_t rank, bool valid) {
        if (valid && rank < cap) out[rank] = kmer;
    });
    if (threadIdx.x == 0) *n = tot;
}

static int bits_for(uint64_t n) {
    int b = 1;
    while ((1ull << b) < n) b++;
    return b;
}

}  // namespace spb

using namespace spb;

#define PF_SMEM_CAP 4096u   // records per CTA in shared memory (32 KB)

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

extern "C" void spb_free_prefilter_result(spb_prefilter_result *r) {
    if (!r) return;
    free(r->offsets);
    free(r->hits);
    free(r);
}

extern "C" int spb_prefilter(spb_ctx *c, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                             const spb_prefilter_params *p, spb_prefilter_result **out) {
    SPB_CHECK(c && q_offsets && p && out, "spb_prefilter: NULL argument");
    SPB_CHECK(c->have_index, "spb_prefilter: build the index first (spb_index_build)");
    SPB_CHECK(p->bin_size >= 2 && p->bin_size <= 2048 && (p->bin_size & (p->bin_size - 1)) == 0,
              "spb_prefilter: bin_size %d must be a power of two in 2..2048", p->bin_size);
    SPB_CHECK(p->max_seqs >= 1, "spb_prefilter: max_seqs must be >= 1");
    SPB_CHECK(p->min_diag_score >= 0 && p->min_diag_score <= 255, "spb_prefilter: min_diag_score out of range");
    SPB_CUDA(cudaSetDevice(c->device));
    *out = nullptr;
    cudaStream_t st = c->stream;
    const uint64_t NQ = q_offsets[Q];
    SPB_CHECK(NQ < 0xFFFFFFF0ull, "spb_prefilter: query batch too large for 32-bit offsets; split the batch");

    spb_prefilter_result *res = (spb_prefilter_result *)calloc(1, sizeof(spb_prefilter_result));
    res->n_queries = Q;
    res->offsets = (uint64_t *)calloc((size_t)Q + 1, sizeof(uint64_t));
    if (Q == 0) { res->hits = (spb_hit *)calloc(1, sizeof(spb_hit)); *out = res; return 0; }

    // ---- upload queries
    std::vector<uint32_t> off32((size_t)Q + 1);
    uint32_t maxL = 0;
    for (uint32_t i = 0; i <= Q; i++) {
        off32[i] = (uint32_t)q_offsets[i];
        if (i) maxL = std::max(maxL, (uint32_t)(q_offsets[i] - q_offsets[i - 1]));
    }
    int rcode = 1;
    DevBuf d_start, d_count, d_status, d_need, d_cursor, d_arena;
    auto fail = [&]() { d_start.release(); d_count.release(); d_status.release(); d_need.release(); d_cursor.release(); d_arena.release(); spb_free_prefilter_result(res); return rcode; };
#define PF_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            spb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return fail();                                                                        \
        }                                                                                         \
    } while (0)
    if (c->d_q.ensure(NQ + 64) || c->d_qoff.ensure(((size_t)Q + 1) * 4)) return fail();
    PF_CUDA(cudaMemcpyAsync(c->d_q.p, q_residues, NQ, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(c->d_qoff.p, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice, st));

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
    const int idBits = bits_for(std::max<uint64_t>(c->T, 2));
    const int P = (int)maxL - c->span + 1;
    A.wBits = bits_for((uint64_t)std::max(P, 2));
    A.rankBits = 48 - idBits - A.wBits;
    if (A.rankBits > 23) A.rankBits = 23;   // KmerGenerator::MAX_KMER_RESULT_SIZE = 2^23
    if (A.rankBits < 10) {
        spb::set_error("spb_prefilter: %u targets x query length %u do not fit the 64-bit hit record; shard the targets", c->T, maxL);
        return fail();
    }
    if (maxL > PF_MAXL_SMEM) {
        spb::set_error("spb_prefilter: query of %u residues exceeds the supported query length %d of the spacer path", maxL, PF_MAXL_SMEM);
        return fail();
    }

    if (d_start.ensure((size_t)Q * 8) || d_count.ensure((size_t)Q * 4) || d_status.ensure((size_t)Q * 4) || d_need.ensure((size_t)Q * 4) ||
        d_cursor.ensure(64)) return fail();
    uint64_t arenaCap = std::min<uint64_t>((uint64_t)Q * (uint64_t)A.maxHits, std::max<uint64_t>((uint64_t)Q * 24, 1u << 20));
    std::vector<uint32_t> h_status(Q), h_count(Q), h_need(Q);
    std::vector<uint64_t> h_start(Q);
    float kernelMs = 0;
    uint64_t nSlow = 0;
    unsigned long long h_cursor[8];
    for (int attempt = 0;; attempt++) {
        if (d_arena.ensure(arenaCap * sizeof(spb_hit) + 64)) return fail();
        A.arena = d_arena.as<spb_hit>();
        A.arenaCap = arenaCap;
        A.cursor = d_cursor.as<unsigned long long>();
        A.outStart = d_start.as<uint64_t>();
        A.outCount = d_count.as<uint32_t>();
        A.status = d_status.as<uint32_t>();
        A.need = d_need.as<uint32_t>();
        A.qlist = nullptr;
        A.gScratch = nullptr;
        A.cap = PF_SMEM_CAP;
        PF_CUDA(cudaMemsetAsync(d_cursor.p, 0, 64, st));
        PF_CUDA(cudaMemsetAsync(d_status.p, 0xFF, (size_t)Q * 4, st));
        const size_t smem = (size_t)PF_SMEM_CAP * 8;
        PF_CUDA(cudaFuncSetAttribute(prefilter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PF_CUDA(cudaEventRecord(c->ev0, st));
        prefilter_kernel<false><<<Q, PF_THREADS, smem, st>>>(A);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaEventRecord(c->ev1, st));
        PF_CUDA(cudaMemcpyAsync(h_status.data(), d_status.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaMemcpyAsync(h_need.data(), d_need.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaMemcpyAsync(h_cursor, d_cursor.p, 64, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        PF_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        kernelMs += ms;
        // ---- slow path: queries whose hit list does not fit shared memory run again on global-memory scratch
        std::vector<uint32_t> slow;
        uint32_t maxNeed = 0;
        bool arenaOverflow = false;
        for (uint32_t i = 0; i < Q; i++) {
            if (h_status[i] == ST_CAPACITY) { slow.push_back(i); maxNeed = std::max(maxNeed, h_need[i]); }
            else if (h_status[i] == ST_ARENA) arenaOverflow = true;
            else if (h_status[i] == ST_LAYOUT) {
                spb::set_error("spb_prefilter: query %u exceeds the hit-record layout (rank bits %d, window bits %d)", i, A.rankBits, A.wBits);
                return fail();
            } else if (h_status[i] != ST_OK) {
                spb::set_error("spb_prefilter: query %u was not processed (status %u)", i, h_status[i]);
                return fail();
            }
        }
        if (!slow.empty() && !arenaOverflow) {
            nSlow = slow.size();
            uint32_t capG = 1;
            while (capG < std::max(maxNeed, 4096u)) capG <<= 1;
            DevBuf d_list, d_scr;
            for (size_t done = 0; done < slow.size() && !arenaOverflow;) {
                // a need estimate can grow once later stages are reached: loop until every query fits
                size_t chunk = std::min<size_t>(slow.size() - done, std::max<size_t>(1, ((size_t)2 << 30) / ((size_t)capG * 8)));
                if (d_list.ensure(chunk * 4) || d_scr.ensure(chunk * (size_t)capG * 8)) { d_list.release(); d_scr.release(); return fail(); }
                PF_CUDA(cudaMemcpyAsync(d_list.p, slow.data() + done, chunk * 4, cudaMemcpyHostToDevice, st));
                A.qlist = d_list.as<uint32_t>();
                A.gScratch = d_scr.as<uint64_t>();
                A.cap = capG;
                PF_CUDA(cudaEventRecord(c->ev0, st));
                prefilter_kernel<true><<<(unsigned)chunk, PF_THREADS, 0, st>>>(A);
                PF_CUDA(cudaGetLastError());
                PF_CUDA(cudaEventRecord(c->ev1, st));
                PF_CUDA(cudaMemcpyAsync(h_status.data(), d_status.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
                PF_CUDA(cudaMemcpyAsync(h_need.data(), d_need.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
                PF_CUDA(cudaStreamSynchronize(st));
                PF_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
                kernelMs += ms;
                bool again = false;
                for (size_t k = done; k < done + chunk; k++) {
                    uint32_t s = h_status[slow[k]];
                    if (s == ST_ARENA) arenaOverflow = true;
                    else if (s == ST_CAPACITY) { again = true; while (capG < h_need[slow[k]]) capG <<= 1; }
                    else if (s != ST_OK) {
                        spb::set_error("spb_prefilter: query %u failed on the global-memory path (status %u)", slow[k], s);
                        d_list.release(); d_scr.release();
                        return fail();
                    }
                }
                if (!again) done += chunk;
                else if (capG > (1u << 28)) {
                    spb::set_error("spb_prefilter: hit list of a query exceeds 2^28 records");
                    d_list.release(); d_scr.release();
                    return fail();
                }
            }
            d_list.release();
            d_scr.release();
            PF_CUDA(cudaMemcpyAsync(h_cursor, d_cursor.p, 64, cudaMemcpyDeviceToHost, st));
            PF_CUDA(cudaStreamSynchronize(st));
        }
        if (!arenaOverflow) break;
        if (arenaCap >= (uint64_t)Q * (uint64_t)A.maxHits || attempt > 8) {
            spb::set_error("spb_prefilter: result arena overflow could not be resolved");
            return fail();
        }
        arenaCap = std::min<uint64_t>((uint64_t)Q * (uint64_t)A.maxHits, arenaCap * 4);
    }

    // ---- compact the arena into query order
    PF_CUDA(cudaMemcpyAsync(h_count.data(), d_count.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
    PF_CUDA(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < Q; i++) res->offsets[i + 1] = res->offsets[i] + h_count[i];
    const uint64_t nHits = res->offsets[Q];
    res->n_hits = nHits;
    res->hits = (spb_hit *)malloc(std::max<uint64_t>(nHits, 1) * sizeof(spb_hit));
    if (nHits) {
        DevBuf d_ooff, d_out;
        if (d_ooff.ensure(((size_t)Q + 1) * 8) || d_out.ensure(nHits * sizeof(spb_hit))) { d_ooff.release(); d_out.release(); return fail(); }
        PF_CUDA(cudaMemcpyAsync(d_ooff.p, res->offsets, ((size_t)Q + 1) * 8, cudaMemcpyHostToDevice, st));
        gather_hits_kernel<<<Q, 64, 0, st>>>(d_arena.as<spb_hit>(), d_start.as<uint64_t>(), d_ooff.as<uint64_t>(), Q, d_out.as<spb_hit>());
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpyAsync(res->hits, d_out.p, nHits * sizeof(spb_hit), cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        d_ooff.release();
        d_out.release();
    }
#undef PF_CUDA
    res->n_kmers = h_cursor[1];
    res->n_db_matches = h_cursor[2];
    res->n_events = h_cursor[3];
    res->n_slow_path = nSlow;
    res->kernel_ms = kernelMs;
    // algorithmic bytes: per similar k-mer one 3-mer row entry (4 B) + two CSR offsets (8 B),
    // per index hit one posting (8 B), per query its residues, per output hit 12 B
    res->algo_bytes = res->n_kmers * 12ull + res->n_db_matches * 8ull + NQ + nHits * 12ull;
    d_start.release(); d_count.release(); d_status.release(); d_need.release(); d_cursor.release(); d_arena.release();
    *out = res;
    return 0;
}
