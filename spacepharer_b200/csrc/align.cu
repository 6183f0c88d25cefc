// align.cu -- gapped alignment of (query frame, target ORF) pairs: score, end/start coordinates, CIGAR, E-value.
//
// Replaces Matcher::getSWResult -> SmithWaterman::ssw_init / ssw_align / banded_sw
// (M/alignment/Matcher.cpp:61-187, StripedSmithWaterman.cpp:130-290,301-687,689-984) and
// EvalueComputation::computeEvalue / computeBitScore (M/alignment/EvalueComputation.h:18-40).
//
// The reference computes a Farrar-striped SW whose E update sees H *before* the cross-stripe lazy-F correction,
// so its result depends on the SIMD layout.  One thread per pair reproduces that exactly with two F accumulators
// per column (no striped emulation needed):
//     Fs : F chain restarted at every stripe start (row % segLen == 0)      -> H1 = max(0, Hdiag + s, E, Fs)
//     Ff : the full F chain over all rows                                     -> H  = max(H1, Ff)
//     E' = max(E - ge, H1 - go)      (never opens from a cell that was raised by a cross-stripe F)
// with segLen = ceil(L / lanes) of the emulated build (byte kernel 32 lanes, word kernel 16 lanes for AVX2).
// The byte kernel's saturation only matters through its overflow test (max + bias >= 255 -> word kernel).
#include <math.h>

#include <cub/cub.cuh>

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "common.cuh"

namespace spb {

// ---- per-query composition bias (SmithWaterman::ssw_init, StripedSmithWaterman.cpp:689-706) ----
// Also emits, for queries short enough for the register kernel (L <= fastMaxL <= 16), the 16-row profile
// prof16[q][a][i] = M[a][q_i] + cb_i (int8; rows >= L and the pad symbol a = 21 hold -128) and qfast[q] = 1.
#define SWF_ROWS 16
#define SWF_THREADS 64
#define SWF_PROF_ROWS 22      // 21 residues + the pad symbol used beyond the end of a target
#define SWF_PAD 21
__global__ void __launch_bounds__(128) align_profile_kernel(const uint8_t *__restrict__ q, const uint64_t *__restrict__ qoff, uint32_t Q,
                                     const Matrices *__restrict__ mat, int compBias, int matMin, signed char *__restrict__ cb,
                                     int *__restrict__ qbias, signed char *__restrict__ prof16, uint8_t *__restrict__ qfast,
                                     int fastMaxL) {
    // one warp per query; matMin = min over the alignment matrix (byte-profile bias, StripedSmithWaterman.cpp:723-737)
    __shared__ int16_t s_align[SPB_ALPHABET * SPB_ALPHABET];
    __shared__ double s_pback[SPB_ALPHABET];
    for (int i = threadIdx.x; i < SPB_ALPHABET * SPB_ALPHABET; i += blockDim.x) s_align[i] = mat->align[i];
    if (threadIdx.x < SPB_ALPHABET) s_pback[threadIdx.x] = mat->pback_align[threadIdx.x];
    __syncthreads();
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const int lane = threadIdx.x & 31;
    const uint64_t o = qoff[qi];
    const int L = (int)(qoff[qi + 1] - o);
    const uint8_t *s = q + o;
    int mn = 0;
    for (int i = lane; i < L; i += 32) {
        signed char c = 0;
        if (compBias) {
            const int lo = max(0, i - 20), hi = min(L, i + 20);
            const int16_t *row = s_align + s[i] * SPB_ALPHABET;
            int sum = 0;
            for (int j = lo; j < hi; j++) sum += row[s[j]];
            sum -= row[s[i]];
            float d = (float)sum;
            d = __double2float_rn(__ddiv_rn((double)d, -1.0 * (double)(float)(hi - lo)));
            for (int a = 0; a < SPB_ALPHABET; a++)
                d = __double2float_rn(__dadd_rn((double)d, __dmul_rn(s_pback[a], (double)(float)row[a])));
            double r = (d < 0.0f) ? __dsub_rn((double)d, 0.5) : __dadd_rn((double)d, 0.5);
            c = (signed char)r;
        }
        cb[o + i] = c;
        mn = min(mn, (int)c);
    }
    for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    const bool fast = (L >= 1 && L <= fastMaxL);
    if (lane == 0) {
        qbias[qi] = abs(matMin) + abs(min(mn, 0));
        qfast[qi] = (uint8_t)fast;
    }
    if (fast) {
        __syncwarp();
        signed char *pq = prof16 + (size_t)qi * SWF_PROF_ROWS * SWF_ROWS;
        for (int x = lane; x < SWF_PROF_ROWS * SWF_ROWS; x += 32) {
            const int a = x / SWF_ROWS, i = x % SWF_ROWS;
            int v = -128;
            if (a < SPB_ALPHABET && i < L) v = (int)(signed char)s_align[a * SPB_ALPHABET + s[i]] + (int)cb[o + i];
            pq[x] = (signed char)v;
        }
    }
}

struct PairOut {      // per pair, written by the forward kernel, completed by reverse / traceback
    int32_t score, qEnd, tEnd, word;
    int32_t qStart, tStart;
    int32_t status;   // 0 rejected (E-value / nothing aligned), 1 accepted, <0 error
    int32_t bits;
    double eval;
    int32_t nIdent, alnLen;
    uint32_t cigOff, cigLen;
};

// One DP pass over columns jBegin, jBegin+step, ... (exclusive jEnd) for one pair held by one thread.
// H/E live at he[i * stride] (int16 pairs).  prof(i, a) = substitution score + composition bias of query row i.
template <class Prof>
__device__ __forceinline__ void sw_pass(const uint8_t *__restrict__ t, int jBegin, int jEnd, int step, int L, int segLen,
                                        int go, int ge, int terminate, Prof prof, short2 *he, int stride,
                                        int &bestScore, int &bestCol, int &bestRow) {
    for (int i = 0; i < L; i++) he[i * stride] = make_short2(0, 0);
    bestScore = 0;
    bestCol = -1;
    bestRow = L - 1;
    for (int j = jBegin; j != jEnd; j += step) {
        const int a = __ldg(t + j);
        int hdiag = 0, Fs = 0, Ff = 0, colMax = 0, colRow = 0, seg = 0;
        for (int i = 0; i < L; i++) {
            if (seg == 0) Fs = 0;
            const short2 old = he[i * stride];
            int h1 = hdiag + prof(i, a);
            h1 = min(h1, 32767);
            h1 = max(max(h1, (int)old.y), max(Fs, 0));
            const int open = max(h1 - go, 0);
            const int e = max(max((int)old.y - ge, 0), open);
            Fs = max(max(Fs - ge, 0), open);
            const int h = max(h1, Ff);
            Ff = max(max(Ff - ge, 0), open);
            hdiag = old.x;
            he[i * stride] = make_short2((short)h, (short)e);
            if (h > colMax) { colMax = h; colRow = i; }
            if (++seg == segLen) seg = 0;
        }
        if (colMax > bestScore) {
            bestScore = colMax;
            bestCol = j;
            bestRow = colRow;
        }
        if (terminate > 0 && colMax == terminate) break;
    }
}

struct FwdItem { uint32_t query, pairStart, count; };

struct AlnArgs {
    const uint8_t *q;
    const uint64_t *qoff;
    const signed char *qcb;
    const int *qbias;
    const uint8_t *tres;
    const uint64_t *toff;
    const Matrices *mat;
    const FwdItem *items;
    uint32_t nItems;
    const uint32_t *pairT;     // target id per pair (work order)
    uint32_t *pairQ;           // out: query of every pair
    uint8_t *accFlag;          // out: 1 if the pair passed the E-value test
    PairOut *out;
    int go, ge, byteLanes, wordLanes;
    double evalThr, dbRes;
    spb_gumbel g;
    int Lpad;                  // padded max query length of the batch (smem layout)
    unsigned long long *cells; // forward cells counter
};

// Forward pass: one warp per (query, <=32 targets) item, one lane per pair.  Per warp shared memory:
// profile int8[21][Lpad] followed by H/E short2[Lpad][32].
__global__ void __launch_bounds__(128) sw_forward_kernel(AlnArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t item = blockIdx.x * nw + wid;
    if (item >= A.nItems) return;
    const size_t perWarp = (size_t)SPB_ALPHABET * A.Lpad + (size_t)A.Lpad * 32 * sizeof(short2);
    signed char *sprof = reinterpret_cast<signed char *>(smem + wid * perWarp);
    short2 *she = reinterpret_cast<short2 *>(smem + wid * perWarp + (size_t)SPB_ALPHABET * A.Lpad);
    const FwdItem it = A.items[item];
    const uint64_t qo = A.qoff[it.query];
    const int L = (int)(A.qoff[it.query + 1] - qo);
    const int Lpad = A.Lpad;
    // profile(i, a) = M[a][q_i] + cb_i  (StripedSmithWaterman.cpp:93-127)
    for (int x = lane; x < SPB_ALPHABET * L; x += 32) {
        const int a = x / L, i = x - a * L;
        sprof[a * Lpad + i] = (signed char)((int)(signed char)A.mat->align[a * SPB_ALPHABET + A.q[qo + i]] + (int)A.qcb[qo + i]);
    }
    __syncwarp();
    if (lane >= (int)it.count) return;
    const uint32_t pair = it.pairStart + lane;
    const uint32_t tid = A.pairT[pair];
    const uint64_t to = A.toff[tid];
    const int tLen = (int)(A.toff[tid + 1] - to);
    const uint8_t *t = A.tres + to;
    const int bias = A.qbias[it.query];
    auto prof = [&](int i, int a) -> int { return (int)sprof[a * Lpad + i]; };
    const int segB = (L + A.byteLanes - 1) / A.byteLanes, segW = (L + A.wordLanes - 1) / A.wordLanes;
    int score, col, row, word = 0;
    sw_pass(t, 0, tLen, 1, L, segB, A.go, A.ge, -1, prof, she + lane, 32, score, col, row);
    if (score + bias >= 255) {          // byte kernel overflow -> word kernel (StripedSmithWaterman.cpp:156-163)
        word = 1;
        if (segW != segB) sw_pass(t, 0, tLen, 1, L, segW, A.go, A.ge, -1, prof, she + lane, 32, score, col, row);
    }
    PairOut o;
    memset(&o, 0, sizeof(o));
    o.score = score;
    o.qEnd = row;
    o.tEnd = col;
    o.word = word;
    o.qStart = o.tStart = -1;
    o.status = 0;
    if (col >= 0) {
        o.eval = dev_evalue(A.g, (double)score, (double)L, A.dbRes);
        o.bits = (int)((A.g.lambda * (double)score - log(A.g.K)) / log(2.0) + 0.5);
        if (!(o.eval > A.evalThr)) o.status = 1;
    }
    A.out[pair] = o;
    A.pairQ[pair] = it.query;
    A.accFlag[pair] = (uint8_t)(o.status == 1);
    atomicAdd(A.cells, (unsigned long long)L * (unsigned long long)tLen);
}

// ---------------------------------------------------------------------------------------------------------
// Register-resident forward pass for short queries (L <= 16: every spacer frame).
//
// One thread owns TWO pairs, packed into the two int16 halves of every register, and keeps the whole DP column
// (16 rows of H and E) in registers; the target is streamed column by column.  With segLen == 1 (L <= lanes of
// both emulated SIMD kernels) the reference recurrence is exactly (StripedSmithWaterman.cpp:382-416)
//     h1 = max(0, Hdiag + s, E)        open = max(0, h1 - go)
//     E' = max(E - ge, open)           H = max(h1, F)            F' = max(F - ge, open)
// i.e. 7 packed DPX/ALU instructions per 2 cells: PRMT (two int8 profile scores -> int16x2), VIADD.16x2,
// 3 x VIADDMNMX.S16x2.RELU, 2 x VIMNMX.S16x2.  Pairs are sorted by target length (descending) so the 64 pairs of a warp
// finish together.  The per-query profile (22 x 16 int8) is staged in shared memory as s_prof[pair][a][thread]:
// a thread reads its own 16-byte column, so the LDS.128 of a quarter warp never conflicts.
// Rows >= L and columns beyond a target's end score -128 and can never raise the maximum.
// ---------------------------------------------------------------------------------------------------------
struct FastArgs {
    const uint32_t *sortedKey;   // tLen + 1 (0: pair not eligible), descending
    const uint32_t *sortedPair;
    uint32_t NP;
    const uint32_t *pairQ, *pairT;
    const uint64_t *qoff;
    const int *qbias;
    const uint4 *prof16;
    const uint8_t *tres;
    const uint64_t *toff;
    PairOut *out;
    uint8_t *accFlag;
    int go, ge;
    uint32_t keyMul;             // 16, as a run-time value: `h * keyMul + row key` stays an IMAD (FMA pipe) instead of an ALU-pipe LEA
    double evalThr, dbRes;
    spb_gumbel g;
    unsigned long long *cells;
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// a * b + c on the FMA pipe (IMAD): keeps the ALU pipe, which the DPX recurrence saturates, free of the key arithmetic
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __noinline__ void sw_finish_pair(const FastArgs &A, uint32_t pair, int L, int tLen, int bias, int score, int col, int row) {
    PairOut o;
    memset(&o, 0, sizeof(o));
    if (score == 0) { col = -1; row = L - 1; }
    o.score = score;
    o.qEnd = row;
    o.tEnd = col;
    o.word = (score + bias >= 255) ? 1 : 0;     // byte kernel overflow -> word kernel; same recurrence when segLen == 1
    o.qStart = o.tStart = -1;
    o.status = 0;
    if (col >= 0) {
        o.eval = dev_evalue(A.g, (double)score, (double)L, A.dbRes);
        o.bits = (int)((A.g.lambda * (double)score - log(A.g.K)) / log(2.0) + 0.5);
        if (!(o.eval > A.evalThr)) o.status = 1;
    }
    A.out[pair] = o;
    A.accFlag[pair] = (uint8_t)(o.status == 1);
}

__global__ void __launch_bounds__(SWF_THREADS) sw_fast_kernel(FastArgs A) {
    __shared__ uint4 s_prof[2][SWF_PROF_ROWS][SWF_THREADS];
    const int tid = threadIdx.x;
    const uint32_t e0 = 2u * (blockIdx.x * SWF_THREADS + tid);
    const uint32_t keyA = e0 < A.NP ? A.sortedKey[e0] : 0u;
    const uint32_t keyB = e0 + 1 < A.NP ? A.sortedKey[e0 + 1] : 0u;
    unsigned long long myCells = 0;
    if (keyA != 0u) {      // keys descend: keyA == 0 implies keyB == 0
        const uint32_t pA = A.sortedPair[e0];
        const uint32_t pB = keyB ? A.sortedPair[e0 + 1] : pA;
        const uint32_t qA = A.pairQ[pA], qB = A.pairQ[pB];
        const uint32_t tiA = A.pairT[pA], tiB = A.pairT[pB];
        const int lenA = (int)(keyA - 1u), lenB = keyB ? (int)(keyB - 1u) : 0;
        const int LA = (int)(A.qoff[qA + 1] - A.qoff[qA]), LB = (int)(A.qoff[qB + 1] - A.qoff[qB]);
#pragma unroll
        for (int a = 0; a < SWF_PROF_ROWS; a++) {
            s_prof[0][a][tid] = __ldg(A.prof16 + (size_t)qA * SWF_PROF_ROWS + a);
            s_prof[1][a][tid] = __ldg(A.prof16 + (size_t)qB * SWF_PROF_ROWS + a);
        }
        // aligned 4-byte windows over the two targets (the residue buffer is padded by 64 bytes)
        const uint8_t *tA = A.tres + A.toff[tiA], *tB = A.tres + A.toff[tiB];
        const uint32_t *wA = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(tA) & ~(uintptr_t)3);
        const uint32_t *wB = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(tB) & ~(uintptr_t)3);
        const uint32_t shA = (uint32_t)(reinterpret_cast<uintptr_t>(tA) & 3) * 8u, shB = (uint32_t)(reinterpret_cast<uintptr_t>(tB) & 3) * 8u;
        // target words are fetched two trips (8 columns) ahead: the pairs of a warp stream 64 different targets, mostly from DRAM
        const int lastA = (int)((shA >> 3) + (uint32_t)lenA + 3u) / 4 + 1, lastB = (int)((shB >> 3) + (uint32_t)lenB + 3u) / 4 + 1;
        uint32_t curA = __ldg(wA), curB = __ldg(wB);
        uint32_t nxA = __ldg(wA + 1), nxB = __ldg(wB + 1);
        const uint32_t MGO = ((uint32_t)(-A.go) & 0xFFFFu) * 0x10001u, MGE = ((uint32_t)(-A.ge) & 0xFFFFu) * 0x10001u;
        uint32_t H[SWF_ROWS], E[SWF_ROWS];
#pragma unroll
        for (int i = 0; i < SWF_ROWS; i++) { H[i] = 0; E[i] = 0; }
        // End position without a branch in the column loop: every cell contributes the key `h << 4 | (15 - row)` to the column
        // maximum (h <= 16 rows x 127 < 2^11, so a key fits the positive half of an int16; the shift cannot carry between
        // the packed halves); the largest key of a column is its best score at the smallest row.  `thr` = best score so far
        // << 4 | 15 per half: a column key above it is a strictly better score -- the first such column wins, as in the
        // reference's striped kernel (StripedSmithWaterman.cpp:450-484).  The key adds run on the FMA pipe (IMAD), beside the
        // DPX instructions of the recurrence on the ALU pipe.
        uint32_t thr = 0x000F000Fu, keyA = 0, keyB = 0;
        const uint32_t KM = A.keyMul;
        int colA = -1, colB = -1;
        const int n = max(lenA, lenB);
        for (int j0 = 0; j0 < n; j0 += 4) {
            // (the shorter target of the pair stops fetching at its own end: never more than 8 bytes past a target)
            const uint32_t n2A = __ldg(wA + min((j0 >> 2) + 2, lastA)), n2B = __ldg(wB + min((j0 >> 2) + 2, lastB));
            uint32_t resA = __funnelshift_r(curA, nxA, shA), resB = __funnelshift_r(curB, nxB, shB);
            curA = nxA;
            curB = nxB;
            nxA = n2A;
            nxB = n2B;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = j0 + jj;
                uint32_t a = resA & 0xFFu, b = resB & 0xFFu;
                resA >>= 8;
                resB >>= 8;
                if (j >= lenA || a > SPB_X) a = SWF_PAD;
                if (j >= lenB || b > SPB_X) b = SWF_PAD;
                const uint4 pa = s_prof[0][a][tid], pb = s_prof[1][b][tid];
                const uint32_t paw[4] = {pa.x, pa.y, pa.z, pa.w}, pbw[4] = {pb.x, pb.y, pb.z, pb.w};
                uint32_t Hd = 0, F = 0, colkey = 0;
#pragma unroll
                for (int i = 0; i < SWF_ROWS; i++) {
                    const int r = i & 3;
                    const uint32_t sel = (uint32_t)r | ((uint32_t)(r | 8) << 4) | ((uint32_t)(4 + r) << 8) | ((uint32_t)((4 + r) | 8) << 12);
                    const uint32_t sc = prmt(paw[i >> 2], pbw[i >> 2], sel);       // {s_A(i), s_B(i)} sign-extended
                    const uint32_t h1 = __viaddmax_s16x2_relu(Hd, sc, E[i]);        // max(Hd + s, E, 0)
                    const uint32_t open = __vadd2(h1, MGO);                         // h1 - go (may be negative)
                    E[i] = __viaddmax_s16x2_relu(E[i], MGE, open);                  // max(E - ge, h1 - go, 0)
                    const uint32_t h = __vmaxs2(h1, F);
                    F = __viaddmax_s16x2_relu(F, MGE, open);
                    Hd = H[i];
                    H[i] = h;
                    colkey = __vmaxs2(colkey, mad_u32(h, KM, (uint32_t)(SWF_ROWS - 1 - i) * 0x10001u));
                }
                const uint32_t nb = __vmaxs2(thr, colkey);
                const uint32_t chg = nb ^ thr;         // non-zero halves: a strictly better column for that pair
                if (chg & 0xFFFFu) { keyA = colkey; colA = j; }
                if (chg > 0xFFFFu) { keyB = colkey; colB = j; }
                thr = nb | 0x000F000Fu;
            }
        }
        const int rowA = SWF_ROWS - 1 - (int)(keyA & 15u), rowB = SWF_ROWS - 1 - (int)((keyB >> 16) & 15u);
        const uint32_t best = ((keyA & 0xFFFFu) >> 4) | ((keyB >> 20) << 16);
        sw_finish_pair(A, pA, LA, lenA, A.qbias[qA], (int)(best & 0xFFFFu), colA, rowA);
        myCells = (unsigned long long)LA * (unsigned long long)lenA;
        if (keyB) {
            sw_finish_pair(A, pB, LB, lenB, A.qbias[qB], (int)(best >> 16), colB, rowB);
            myCells += (unsigned long long)LB * (unsigned long long)lenB;
        }
    }
    for (int d = 16; d > 0; d >>= 1) myCells += __shfl_down_sync(0xffffffffu, myCells, d);
    if ((tid & 31) == 0 && myCells) atomicAdd(A.cells, myCells);
}

// pair -> query map and the sort key of the register kernel (tLen + 1 for eligible queries, else 0)
__global__ void pair_key_kernel(const uint64_t *__restrict__ pairOff, const uint32_t *__restrict__ pairT,
                                const uint64_t *__restrict__ toff, const uint8_t *__restrict__ qfast, uint32_t Q,
                                uint32_t *__restrict__ pairQ, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const uint64_t a = pairOff[q], b = pairOff[q + 1];
    const bool f = qfast[q] != 0;
    for (uint64_t k = a + (threadIdx.x & 31); k < b; k += 32) {
        const uint32_t t = pairT[k];
        const uint64_t tl = toff[t + 1] - toff[t];
        pairQ[k] = q;
        keys[k] = (f && tl < 0x7FFFFFFFull) ? (uint32_t)tl + 1u : 0u;
        vals[k] = (uint32_t)k;
    }
}

// Reverse pass + banded traceback for accepted pairs: one thread per pair, scratch in global memory.
struct TbArgs {
    const uint8_t *q;
    const uint64_t *qoff;
    const signed char *qcb;
    const uint8_t *tres;
    const uint64_t *toff;
    const Matrices *mat;
    const uint32_t *acc;       // accepted pair indices
    uint32_t nAcc;
    const uint32_t *pairQ, *pairT;
    PairOut *out;
    int go, ge, byteLanes, wordLanes;
    short2 *heScratch;         // [nAcc] slices, offsets heOff
    const uint64_t *heOff;
    // traceback scratch
    const uint64_t *tbOff;     // [nAcc+1] byte offsets into tbScratch
    unsigned char *tbScratch;
    const uint64_t *cigOff;    // [nAcc+1] offsets into cigar
    char *cigar;
    const uint8_t *qfast;      // per query: 1 = register-kernel query (L <= 16, segLen 1), 16-row int8 profile in prof16
    const uint4 *prof16;
};

// ---------------------------------------------------------------------------------------------------------
// Start coordinates: the reverse pass (StripedSmithWaterman.cpp:202-231).  Queries of the register kernel (qfast: L <= 16
// and segLen == 1 under both emulated SIMD widths) run it the same way as the forward pass: the column of H / E lives in
// registers, rows are walked in query coordinates from qEnd down to 0 (the reversed prefix), the int8 profile row of the
// target residue is one 16-byte load, and the pass stops at the first column whose maximum equals the forward score.
// Longer queries take the general striped-layout emulation (sw_pass) on a global H/E slice.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sw_reverse_kernel(TbArgs A) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.nAcc) return;
    const uint32_t pair = A.acc[k];
    PairOut o = A.out[pair];
    const uint32_t qi = A.pairQ[pair], ti = A.pairT[pair];
    const uint8_t *t = A.tres + A.toff[ti];
    const int qEnd = o.qEnd;
    int score = 0, col = -1, row = qEnd;
    if (A.qfast[qi] && qEnd < SWF_ROWS) {
        const uint4 *prof = A.prof16 + (size_t)qi * SWF_PROF_ROWS;
        int H[SWF_ROWS], E[SWF_ROWS];
#pragma unroll
        for (int p = 0; p < SWF_ROWS; p++) { H[p] = 0; E[p] = 0; }
        const int go = A.go, ge = A.ge, want = o.score;
        for (int j = o.tEnd; j >= 0; j--) {
            uint32_t a = __ldg(t + j);
            if (a > SPB_X) a = SPB_X;
            const uint4 pr = __ldg(prof + a);
            const uint32_t pw[4] = {pr.x, pr.y, pr.z, pr.w};
            int hd = 0, F = 0, colMax = 0, colP = qEnd;
#pragma unroll
            for (int p = SWF_ROWS - 1; p >= 0; p--) {
                if (p <= qEnd) {
                    const int s = (int)(pw[p >> 2] << (24 - 8 * (p & 3))) >> 24;
                    const int h1 = max(max(hd + s, E[p]), 0);
                    const int open = max(h1 - go, 0);
                    E[p] = max(max(E[p] - ge, 0), open);
                    const int h = max(h1, F);
                    F = max(max(F - ge, 0), open);
                    hd = H[p];
                    H[p] = h;
                    if (h > colMax) { colMax = h; colP = p; }      // first (largest p = smallest reversed row) strictly better cell
                }
            }
            if (colMax > score) { score = colMax; col = j; row = qEnd - colP; }
            if (want > 0 && colMax == want) break;
        }
    } else {
        const uint64_t qo = A.qoff[qi];
        const uint8_t *q = A.q + qo;
        const signed char *cb = A.qcb + qo;
        const int rl = qEnd + 1;
        const int lanes = o.word ? A.wordLanes : A.byteLanes;
        const int seg = (rl + lanes - 1) / lanes;
        const int16_t *M = A.mat->align;
        auto prof = [&](int i, int a) -> int { return (int)(signed char)M[a * SPB_ALPHABET + q[qEnd - i]] + (int)cb[qEnd - i]; };
        sw_pass(t, o.tEnd, -1, -1, rl, seg, A.go, A.ge, o.score, prof, A.heScratch + A.heOff[k], 1, score, col, row);
    }
    if (score != o.score || col < 0) {
        o.status = -1;      // "Score of forward/backward SW differ" (StripedSmithWaterman.cpp:225-228)
    } else {
        o.tStart = col;
        o.qStart = qEnd - row;
    }
    A.out[pair] = o;
}

// ---------------------------------------------------------------------------------------------------------
// Backtrace: the banded global-in-a-local-box DP of SmithWaterman::banded_sw (StripedSmithWaterman.cpp:781-984) over the
// aligned box [qStart, qEnd] x [tStart, tEnd], restated in BAND coordinates.
//
// A row i of the band covers columns x .. x + 2 bw with x = max(i - bw, 0); slot u = j - x + 1 (1 .. 2 bw + 1) of three row
// arrays holds H of the previous row (hb), E (eb) and H of the current row (hc); slots 0 and `edge` are the zero borders.
// Because x advances by one per row once i > bw, the cell above (i-1, j) sits at slot u + s and the diagonal cell at slot
// u + s - 1 with s = (i > bw) -- the whole row is a static sweep over the slots, so the arrays live in REGISTERS (fully
// unrolled, no indexed local memory).  A cell's three direction codes (E: 2|3, F: 4|5, H: 1 or a copy of the E / F code)
// pack into 5 bits; a band row of <= 9 cells is one 64-bit word kept in shared memory ([row][thread], conflict free).  The
// band doubles until the forward score is reached; the walk back from the bottom-right cell reads the row words.  The reference's
// tie rules are kept verbatim: E opens only on h - go > e - ge (:852), F likewise (:857), the diagonal wins ties against
// the gap states and E wins against F only when strictly greater (:872-873).  `edge = min(end + 1, width - 1)` is applied
// as a SLOT number exactly as the reference does (it zeroes a live slot when the band is clipped by the end of the target).
// Boxes whose band outgrows 9 cells or whose rows exceed the shared-memory slice run the same sweep with the arrays in a
// global scratch slice (band_pass_mem).
// ---------------------------------------------------------------------------------------------------------
#define TB_BW 4                        // widest band of the register path
#define TB_W (2 * TB_BW + 1)           // cells per row
#define TB_ROWS 24                     // rows of the shared-memory direction words
#define TB_CIG 64                      // ops of the packed op list
#define TB_THREADS 128

struct BandCell { int e, f, h, code; };
// one cell: hAbove / eAbove = H, E of (i-1, j); hLeft = H of (i, j-1); hDiag = H of (i-1, j-1); f = running F of the row
__device__ __forceinline__ BandCell band_cell(bool firstRow, int hAbove, int eAbove, int hLeft, int hDiag, int f, int sub, int go, int ge) {
    BandCell c;
    int t1 = firstRow ? -go : hAbove - go;
    int t2 = firstRow ? -ge : eAbove - ge;
    c.e = t1 > t2 ? t1 : t2;
    const int de = t1 > t2 ? 3 : 2;
    t1 = hLeft - go;
    t2 = f - ge;
    c.f = t1 > t2 ? t1 : t2;
    const int df = t1 > t2 ? 5 : 4;
    const int e1 = max(c.e, 0), f1 = max(c.f, 0);
    t1 = max(e1, f1);
    t2 = hDiag + sub;
    c.h = t1 > t2 ? t1 : t2;
    const int dh = (t1 <= t2) ? 1 : (e1 > f1 ? de : df);
    c.code = dh | ((de & 1) << 3) | ((df & 1) << 4);
    return c;
}
// direction code of (row word, cell, state): 0 when the cell was never written
__device__ __forceinline__ int band_code(unsigned long long w, int cell, int st) {
    const int bits = (int)(w >> (5 * cell)) & 31;
    if ((bits & 7) == 0) return 0;
    return st == 2 ? (bits & 7) : (st == 0 ? 2 + ((bits >> 3) & 1) : 4 + ((bits >> 4) & 1));
}

// one pass over the box with band bw <= BWMAX; the row arrays (slots 0 .. 2 TB_BW + 2) persist between passes like the
// reference's heap arrays.  Returns the largest H.
// qs / cb / db: the thread's column of the staged sequences (element i at [i * TB_THREADS])
template <int BWMAX>
__device__ __forceinline__ int band_pass_reg(int (&hb)[TB_W + 2], int (&eb)[TB_W + 2], int (&hc)[TB_W + 2], unsigned long long *dirs,
                                             const uint8_t *qs, const signed char *cb, const uint8_t *db, const signed char *s_mat,
                                             int qLen, int dbLen, int bw, int go, int ge) {
    constexpr int W = 2 * BWMAX + 1;
    const int width = 2 * bw + 3;
#pragma unroll
    for (int u = 1; u <= W; u++) if (u < width - 1) hb[u] = 0;
    int mx = 0;
    for (int i = 0; i < qLen; i++) {
        const int x = max(i - bw, 0), end = min(dbLen - 1, i + bw);
        const int edge = min(end + 1, width - 1), uEnd = end - x + 1;
        const bool s = i > bw, first = (i == 0);
        hb[0] = eb[0] = hc[0] = 0;
#pragma unroll
        for (int u = 1; u <= W + 1; u++) if (u == edge) { hb[u] = 0; eb[u] = 0; }
        const signed char *mrow = s_mat + (int)qs[i * TB_THREADS] * 32;
        const int cbi = cb[i * TB_THREADS];
        int f = 0;
        unsigned long long word = 0;
#pragma unroll
        for (int u = 1; u <= W; u++) {
            if (u <= uEnd) {
                const int sub = (int)mrow[db[(u + x - 1) * TB_THREADS]] + cbi;
                const BandCell c = band_cell(first, s ? hb[u + 1] : hb[u], s ? eb[u + 1] : eb[u], hc[u - 1], s ? hb[u] : hb[u - 1], f, sub, go, ge);
                eb[u] = c.e;
                f = c.f;
                hc[u] = c.h;
                mx = max(mx, c.h);
                word |= (unsigned long long)c.code << (5 * (u - 1));
            }
        }
#pragma unroll
        for (int u = 1; u <= W; u++) if (u <= uEnd) hb[u] = hc[u];
        dirs[(size_t)i * TB_THREADS] = word;
    }
    return mx;
}

// the same sweep with the row arrays and one direction byte per cell in memory (any band width)
__device__ __noinline__ int band_pass_mem(int *hb, int *eb, int *hc, uint8_t *dirs, const uint8_t *qs, const signed char *cb, const uint8_t *db,
                                          const signed char *s_mat, int qLen, int dbLen, int bw, int go, int ge) {
    const int width = 2 * bw + 3, cells = 2 * bw + 1;
    for (int u = 1; u < width - 1; u++) hb[u] = 0;
    int mx = 0;
    for (int i = 0; i < qLen; i++) {
        const int x = max(i - bw, 0), end = min(dbLen - 1, i + bw);
        const int edge = min(end + 1, width - 1), uEnd = end - x + 1;
        const int s = i > bw ? 1 : 0;
        hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
        const signed char *mrow = s_mat + (int)qs[i] * 32;
        const int cbi = cb[i];
        int f = 0;
        uint8_t *drow = dirs + (size_t)i * cells;
        for (int u = 1; u <= uEnd; u++) {
            const int sub = (int)mrow[db[u + x - 1]] + cbi;
            const BandCell c = band_cell(i == 0, hb[u + s], eb[u + s], hc[u - 1], hb[u + s - 1], f, sub, go, ge);
            eb[u] = c.e;
            f = c.f;
            hc[u] = c.h;
            mx = max(mx, c.h);
            drow[u - 1] = (uint8_t)c.code;
        }
        for (int u = uEnd; u < cells; u++) drow[u] = 0;
        for (int u = 1; u <= uEnd; u++) hb[u] = hc[u];
    }
    return mx;
}

// Walk back from the bottom-right cell (StripedSmithWaterman.cpp:882-960).  code(i, cell, st) returns the direction code or 0.
// Ops are produced last-to-first: emit(op) with op 0 = M, 1 = I, 2 = D.  Returns false on a missing code or when the op list
// is full ("Trace back error").
template <class Code, class Emit>
__device__ __forceinline__ bool band_walk(int qLen, int dbLen, int bw, uint32_t cap, Code code, Emit emit, uint32_t &n) {
    const int cells = 2 * bw + 1;
    int i = qLen - 1, j = dbLen - 1, st = 2;
    n = 0;
    while (i > 0 || j > 0) {
        // the reference addresses its direction bytes as one flat array: a cell left / right of the row's band is the last /
        // first cell of the neighbouring row
        int r = i, cell = j - max(i - bw, 0);
        if (cell < 0) { r--; cell += cells; }
        else if (cell >= cells) { r++; cell -= cells; }
        if (r < 0 || r >= qLen || cell < 0 || cell >= cells) return false;
        int op;
        switch (code(r, cell, st)) {
            case 1: --i; --j; st = 2; op = 0; break;
            case 2: --i; st = 0; op = 1; break;
            case 3: --i; st = 2; op = 1; break;
            case 4: --j; st = 1; op = 2; break;
            case 5: --j; st = 2; op = 2; break;
            default: return false;
        }
        if (n + 1 >= cap) return false;
        emit(op);
        n++;
    }
    return true;
}

__global__ void __launch_bounds__(TB_THREADS) sw_traceback_kernel(TbArgs A) {
    __shared__ unsigned long long s_dirs[TB_ROWS * TB_THREADS];
    __shared__ uint8_t s_qs[TB_ROWS * TB_THREADS], s_ts[TB_CIG * TB_THREADS];     // the box's query rows / target columns, [i][thread]
    __shared__ signed char s_cb[TB_ROWS * TB_THREADS];
    __shared__ signed char s_mat[SPB_ALPHABET * 32];
    for (int x = threadIdx.x; x < SPB_ALPHABET * 32; x += blockDim.x)
        s_mat[x] = (x & 31) < SPB_ALPHABET ? (signed char)A.mat->align[(x >> 5) * SPB_ALPHABET + (x & 31)] : (signed char)0;
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.nAcc) return;
    const uint32_t pair = A.acc[k];
    PairOut o = A.out[pair];
    if (o.status != 1) return;
    const uint32_t qi = A.pairQ[pair], ti = A.pairT[pair];
    const uint64_t qo = A.qoff[qi];
    const uint8_t *qs = A.q + qo + o.qStart;
    const signed char *cb = A.qcb + qo + o.qStart;
    const uint8_t *db = A.tres + A.toff[ti] + o.tStart;
    const int dbLen = o.tEnd - o.tStart + 1, qLen = o.qEnd - o.qStart + 1;
    const int score = o.score, go = A.go, ge = A.ge;
    const int bw0 = abs(dbLen - qLen) + 1;
    char *cig = A.cigar + A.cigOff[k];
    const uint32_t cigCap = (uint32_t)(A.cigOff[k + 1] - A.cigOff[k]);
    uint32_t n = 0;
    int ids = 0;
    bool bad = false, done = false, idsDone = false, allM = false;
    // ops as 2-bit codes, last-to-first, in two registers
    unsigned long long opsLo = 0, opsHi = 0;
    auto emit = [&](int op) {
        if (n < 32) opsLo |= (unsigned long long)op << (2 * n);
        else opsHi |= (unsigned long long)op << (2 * (n - 32));
    };
    auto op_at = [&](uint32_t x) -> int { return (int)((x < 32 ? opsLo >> (2 * x) : opsHi >> (2 * (x - 32))) & 3ull); };
    if (bw0 <= TB_BW && qLen <= TB_ROWS && qLen + dbLen + 2 <= TB_CIG) {
        int hb[TB_W + 2], eb[TB_W + 2], hc[TB_W + 2];
#pragma unroll
        for (int u = 0; u < TB_W + 2; u++) { hb[u] = 0; eb[u] = 0; hc[u] = 0; }
        unsigned long long *dirs = s_dirs + threadIdx.x;
        // the box's sequences go to shared memory first: independent loads, and every read of the sweeps below is an LDS
        uint8_t *sq = s_qs + threadIdx.x, *st = s_ts + threadIdx.x;
        signed char *sc = s_cb + threadIdx.x;
        for (int i = 0; i < qLen; i++) { sq[i * TB_THREADS] = qs[i]; sc[i * TB_THREADS] = cb[i]; }
        for (int j = 0; j < dbLen; j++) st[j * TB_THREADS] = db[j];
        // Ungapped boxes need no DP.  If the box is square and its main diagonal sums to the forward score, the reference's DP can
        // only return all matches: every H on the diagonal is a local alignment score that the rest of the diagonal extends,
        // so none can exceed the prefix sum without beating the optimum; the diagonal then wins every tie against the gap states
        // (:872), the first band reaches the score (no doubling, :877) and the walk never leaves it.  (qLen >= 4 keeps the
        // clipped-band slot rule of :837 out of the picture.)  Nine boxes in ten of a spacer search end here.
        if (qLen == dbLen && qLen >= 4) {
            int sum = 0;
            for (int i = 0; i < qLen; i++) sum += (int)s_mat[(int)sq[i * TB_THREADS] * 32 + st[i * TB_THREADS]] + sc[i * TB_THREADS];
            if (sum == score) {
                for (int i = 0; i < qLen; i++) { cig[i] = 'M'; ids += (sq[i * TB_THREADS] == st[i * TB_THREADS]); }
                n = (uint32_t)qLen;
                done = true;
                idsDone = true;
                allM = true;
            }
        }
        int bw = bw0, mx = 0;
        bool fits = true;
        for (; !allM;) {
            if (bw == 1) mx = band_pass_reg<1>(hb, eb, hc, dirs, sq, sc, st, s_mat, qLen, dbLen, bw, go, ge);
            else if (bw == 2) mx = band_pass_reg<2>(hb, eb, hc, dirs, sq, sc, st, s_mat, qLen, dbLen, bw, go, ge);
            else mx = band_pass_reg<TB_BW>(hb, eb, hc, dirs, sq, sc, st, s_mat, qLen, dbLen, bw, go, ge);
            if (mx >= score) break;
            bw *= 2;
            if (bw > TB_BW) { fits = false; break; }     // replayed from the start on the scratch slice below
        }
        if (fits && !allM) {
            done = true;
            bad = !band_walk(qLen, dbLen, bw, min(cigCap, (uint32_t)TB_CIG),
                             [&](int r, int cell, int st) { return band_code(dirs[(size_t)r * TB_THREADS], cell, st); }, emit, n);
        }
    }
    if (!done) {
        // scratch slice: [hb | eb | hc] int32 each Wmax, then one direction byte per cell; ops go to the output slice directly
        const uint64_t bytes = A.tbOff[k + 1] - A.tbOff[k];
        unsigned char *base = A.tbScratch + A.tbOff[k];
        int bwMax = bw0;
        while (bwMax < max(qLen, dbLen)) bwMax *= 2;
        const int64_t Wmax = (int64_t)bwMax * 2 + 3 + 2;
        int *hb = reinterpret_cast<int *>(base), *eb = hb + Wmax, *hc = eb + Wmax;
        uint8_t *dirs = reinterpret_cast<uint8_t *>(hc + Wmax);
        const int64_t dirCap = (int64_t)bytes - 3 * Wmax * 4;
        for (int64_t x = 0; x < 3 * Wmax; x++) hb[x] = 0;
        int bw = bw0;
        for (;;) {
            if ((int64_t)bw * 2 + 3 + 1 > Wmax || ((int64_t)bw * 2 + 1) * qLen > dirCap) { bad = true; break; }
            if (band_pass_mem(hb, eb, hc, dirs, qs, cb, db, s_mat, qLen, dbLen, bw, go, ge) >= score) break;
            bw *= 2;
        }
        if (!bad) {
            const int cells = 2 * bw + 1;
            bad = !band_walk(qLen, dbLen, bw, cigCap,
                             [&](int r, int cell, int st) {
                                 const int bits = dirs[(size_t)r * cells + cell];
                                 return (bits & 7) == 0 ? 0 : (st == 2 ? (bits & 7) : (st == 0 ? 2 + ((bits >> 3) & 1) : 4 + ((bits >> 4) & 1)));
                             },
                             [&](int op) { cig[n] = "MID"[op]; }, n);
            if (!bad) {
                cig[n++] = 'M';
                for (uint32_t a = 0, b = n - 1; a < b; a++, b--) { const char tmp = cig[a]; cig[a] = cig[b]; cig[b] = tmp; }
            } else {
                for (uint32_t x = 0; x < cigCap; x++) cig[x] = 0;    // a failed walk leaves no partial ops behind
            }
        }
    } else if (!bad && !allM) {
        // the walk produced the path end-to-start; the start cell is the final forced 'M' of the reference.  Ops and identities
        // (Matcher.cpp:61-187) in one pass over the packed list, residues from the staged box: nothing is read back from global memory
        const uint8_t *sq = s_qs + threadIdx.x, *st = s_ts + threadIdx.x;
        cig[0] = 'M';
        ids = (sq[0] == st[0]);
        int qi2 = 1, ti2 = 1;
        for (uint32_t x = 1; x <= n; x++) {
            const int op = op_at(n - x);
            cig[x] = "MID"[op];
            if (op == 0) { ids += (sq[qi2 * TB_THREADS] == st[ti2 * TB_THREADS]); qi2++; ti2++; }
            else if (op == 1) qi2++;
            else ti2++;
        }
        n++;
        idsDone = true;
    }
    if (!bad) {
        if (!idsDone) {
            int qi2 = 0, ti2 = 0;
            for (uint32_t x = 0; x < n; x++) {
                const char op = cig[x];
                if (op == 'M') { ids += (qs[qi2] == db[ti2]); qi2++; ti2++; }
                else if (op == 'I') qi2++;
                else ti2++;
            }
        }
        o.cigLen = n;
        o.nIdent = ids;
        o.alnLen = (int)n;
    } else {
        // reference: "Trace back error" -> no CIGAR; alnLength from coordinates, no identities (Matcher.cpp:141-152)
        o.cigLen = 0;
        o.nIdent = 0;
        o.alnLen = max(abs(o.qEnd - o.qStart), abs(o.tEnd - o.tStart)) + 1;
    }
    o.cigOff = (uint32_t)A.cigOff[k];
    A.out[pair] = o;
}


// ---- work-item construction, scratch sizing and row packing (all on the device) ----
__global__ void item_count_kernel(const uint64_t *__restrict__ pairOff, const uint8_t *__restrict__ qfast, uint32_t Q,
                                  uint32_t *__restrict__ cnt) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) cnt[q] = qfast[q] ? 0u : (uint32_t)((pairOff[q + 1] - pairOff[q] + 31) >> 5);
    if (q == Q) cnt[q] = 0;
}

__global__ void item_fill_kernel(const uint64_t *__restrict__ pairOff, const uint32_t *__restrict__ itemBase, uint32_t Q,
                                 FwdItem *__restrict__ items) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const uint64_t a = pairOff[q], b = pairOff[q + 1];
    uint32_t k = itemBase[q];
    if (itemBase[q + 1] == k) return;      // query handled by the register kernel
    for (uint64_t s = a; s < b; s += 32, k++) {
        FwdItem it;
        it.query = q;
        it.pairStart = (uint32_t)s;
        it.count = (uint32_t)min((uint64_t)32, b - s);
        items[k] = it;
    }
}

__global__ void hit_ids_kernel(const spb_hit *__restrict__ hits, uint64_t n, uint32_t *__restrict__ ids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = hits[i].id;
}

__global__ void he_size_kernel(const PairOut *__restrict__ out, const uint32_t *__restrict__ acc, uint32_t nAcc, uint64_t *__restrict__ sz) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nAcc) sz[k] = (uint64_t)out[acc[k]].qEnd + 1;
    if (k == nAcc) sz[k] = 0;
}

__global__ void tb_size_kernel(const PairOut *__restrict__ out, const uint32_t *__restrict__ acc, uint32_t nAcc,
                               uint64_t *__restrict__ tbSz, uint64_t *__restrict__ cigSz, int *__restrict__ err) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == nAcc) { tbSz[k] = 0; cigSz[k] = 0; }
    if (k >= nAcc) return;
    const PairOut o = out[acc[k]];
    if (o.status != 1) { atomicMax(err, (int)k + 1); tbSz[k] = 0; cigSz[k] = 0; return; }
    const long long ql = o.qEnd - o.qStart + 1, dl = o.tEnd - o.tStart + 1;
    long long bw = llabs(dl - ql) + 1;
    while (bw < max(ql, dl)) bw *= 2;
    const long long Wmax = bw * 2 + 3 + 2;
    unsigned long long bytes = (unsigned long long)(3 * Wmax * 4 + (bw * 2 + 1) * ql * 3 + 16);
    tbSz[k] = (bytes + 15) & ~15ull;
    cigSz[k] = (uint64_t)(ql + dl + 2);
}

__global__ void count_widen_kernel(const uint32_t *__restrict__ cnt, uint32_t Q, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Q) out[i] = cnt[i];
    if (i == Q) out[i] = 0;
}

// sort key of a row in the order of Matcher::compareHits (M/alignment/Matcher.h:156-167): E-value asc (double), bit score desc,
// target length asc, target key asc
struct RowKey { double eval; int bits; uint32_t tlen, tkey, pad; };

// Util::computeSeqId (M/commons/Util.cpp:588-598): identities over the alignment length (--seq-id-mode 0), the shorter (1) or the
// longer (2) sequence
__device__ __forceinline__ float seq_id_of(int mode, int nIdent, int qLen, int tLen, int alnLen) {
    const int den = mode == 1 ? min(qLen, tLen) : mode == 2 ? max(qLen, tLen) : alnLen;
    return (float)nIdent / (float)den;
}

// Acceptance criteria beyond the E-value (all zero in the predictmatch search).  The reference skips a pair whose lengths cannot
// reach the coverage before aligning it (Util::canBeCovered, M/commons/Util.cpp:533-550; Alignment.cpp:378-381) and drops an
// alignment that fails Alignment::checkCriteria (Alignment.cpp:555-574; Util::hasCoverage Util.cpp:552-567; the coverage exit of
// ssw_align, StripedSmithWaterman.cpp:226-232, rejects the same pairs).  --max-accept / --max-rejected are unlimited on this path,
// so the outcome of a pair does not depend on its neighbours and both tests run on the finished alignment, one thread per pair.
struct Criteria { float covThr; int covMode; float seqIdThr; int alnLenThr; int seqIdMode; int swMode; };

// Alignment modes without a backtrace (Alignment::initSWMode, Alignment.cpp:174-196; Matcher::getSWResult, Matcher.cpp:140-169).
// swMode 1 = SCORE_COV: start and end positions, the alignment length is the longer span, the sequence identity is ESTIMATED from
// the raw score per column (Matcher::estimateSeqIdByScorePerCol, Matcher.cpp:205-209: float quotient, double constants).
// swMode 2 = SCORE_ONLY: forward pass only -- start positions -1, coverage 0, the estimate runs over the end positions.
__device__ __forceinline__ float est_seq_id(int score, unsigned qa, unsigned ta) {
    const float perCol = (float)(score & 0xFFFF) / (float)max(qa, ta);
    const float e = __double2float_rn(__dadd_rn(__dmul_rn((double)perCol, 0.1656), 0.1141));
    return fmaxf(0.0f, fminf(e, 1.0f));
}
struct ModeRow { int qStart, tStart, alnLen; float seqId, qCov, tCov; };
__device__ __forceinline__ ModeRow mode_row(const PairOut &o, int swMode, unsigned qLen, unsigned tLen) {
    ModeRow m;
    if (swMode == 2) {
        m.qStart = -1; m.tStart = -1;
        m.alnLen = max(abs(o.qEnd + 1), abs(o.tEnd + 1)) + 1;          // Matcher::computeAlnLength with both starts at -1
        m.seqId = est_seq_id(o.score, (unsigned)max(o.qEnd, 1), (unsigned)max(o.tEnd, 1));
        m.qCov = 0.0f; m.tCov = 0.0f;
    } else {
        m.qStart = o.qStart; m.tStart = o.tStart;
        m.alnLen = max(abs(o.qEnd - o.qStart), abs(o.tEnd - o.tStart)) + 1;
        m.seqId = est_seq_id(o.score, (unsigned)max(o.qEnd - o.qStart, 1), (unsigned)max(o.tEnd - o.tStart, 1));
        m.qCov = (float)(unsigned)(o.qEnd - o.qStart + 1) / (float)qLen;
        m.tCov = (float)(unsigned)(o.tEnd - o.tStart + 1) / (float)tLen;
    }
    return m;
}

__device__ __forceinline__ bool can_be_covered(float covThr, int covMode, float qLen, float tLen) {
    switch (covMode) {      // Parameters.h:246-251
    case 0: return (qLen / tLen >= covThr) && (tLen / qLen >= covThr);
    case 1: return (qLen / tLen) >= covThr;
    case 2: return (tLen / qLen) >= covThr;
    case 3: return ((tLen / qLen) >= covThr) && (tLen / qLen) <= 1.0f;
    case 4: return ((qLen / tLen) >= covThr) && (qLen / tLen) <= 1.0f;
    case 5: return (fminf(tLen, qLen) / fmaxf(tLen, qLen)) >= covThr;
    default: return true;
    }
}

__device__ __forceinline__ bool has_coverage(float covThr, int covMode, float qCov, float tCov) {
    switch (covMode) {
    case 0: return (qCov >= covThr) && (tCov >= covThr);
    case 1: return tCov >= covThr;
    case 2: return qCov >= covThr;
    default: return true;
    }
}

__global__ void criteria_kernel(const PairOut *__restrict__ out, const uint32_t *__restrict__ acc, uint32_t nAcc,
                                const uint32_t *__restrict__ pairQ, const uint32_t *__restrict__ pairT,
                                const uint64_t *__restrict__ qoff, const uint64_t *__restrict__ toff, Criteria cr,
                                uint8_t *__restrict__ keep) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nAcc) return;
    const uint32_t pair = acc[k];
    const PairOut o = out[pair];
    const uint32_t q = pairQ[pair], t = pairT[pair];
    const unsigned qLen = (unsigned)(qoff[q + 1] - qoff[q]), tLen = (unsigned)(toff[t + 1] - toff[t]);
    bool ok = can_be_covered(cr.covThr, cr.covMode, (float)qLen, (float)tLen);
    if (cr.swMode != 0) {
        const ModeRow m = mode_row(o, cr.swMode, qLen, tLen);
        ok = ok && has_coverage(cr.covThr, cr.covMode, m.qCov, m.tCov) && m.seqId >= cr.seqIdThr && m.alnLen >= cr.alnLenThr;
        keep[k] = ok ? 1 : 0;
        return;
    }
    // SmithWaterman::computeCov (StripedSmithWaterman.cpp:1055-1057) on the final coordinates
    const float qCov = (float)(unsigned)(o.qEnd - o.qStart + 1) / (float)qLen;
    const float tCov = (float)(unsigned)(o.tEnd - o.tStart + 1) / (float)tLen;
    ok = ok && has_coverage(cr.covThr, cr.covMode, qCov, tCov);
    ok = ok && seq_id_of(cr.seqIdMode, o.nIdent, (int)qLen, (int)tLen, o.alnLen) >= cr.seqIdThr;
    ok = ok && o.alnLen >= cr.alnLenThr;
    keep[k] = ok ? 1 : 0;
}

// --max-accept / --max-rejected (Alignment.cpp:353-404): the reference walks the prefilter list of a query in order and stops once
// max_accept pairs passed, or max_rejected pairs IN A ROW failed (lengths that cannot reach the coverage, E-value, checkCriteria).
// Pairs behind the stop are never aligned there; here every pair has its verdict already (final[pair] = 1: passes), so one thread
// per query replays the walk over the verdicts and clears the pairs behind the stop.  order[] maps list position -> pair (NULL: the
// work order is the list order).
__global__ void set_flags_kernel(const uint32_t *__restrict__ acc, uint32_t n, uint8_t *__restrict__ final_) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) final_[acc[k]] = 1;
}
__global__ void list_order_kernel(const uint64_t *__restrict__ pairOff, const uint32_t *__restrict__ pairQ, const uint32_t *__restrict__ rank,
                                  uint32_t NP, uint32_t *__restrict__ order) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < NP) order[pairOff[pairQ[k]] + rank[k]] = k;
}
__global__ void accept_limit_kernel(const uint64_t *__restrict__ pairOff, uint32_t Q, const uint32_t *__restrict__ order,
                                    uint32_t maxAccept, uint32_t maxRejected, uint8_t *__restrict__ final_) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const uint64_t a = pairOff[q], b = pairOff[q + 1];
    uint32_t passed = 0, rejected = 0;
    uint64_t r = a;
    for (; r < b && passed < maxAccept && rejected < maxRejected; r++) {
        const uint32_t pair = order ? order[r] : (uint32_t)r;
        if (final_[pair]) { passed++; rejected = 0; } else rejected++;
    }
    for (; r < b; r++) final_[order ? order[r] : (uint32_t)r] = 0;
}
__global__ void gather_flags_kernel(const uint32_t *__restrict__ acc, uint32_t n, const uint8_t *__restrict__ final_, uint8_t *__restrict__ keep) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keep[k] = final_[acc[k]];
}

__global__ void pack_rows_kernel(const PairOut *__restrict__ out, const uint32_t *__restrict__ acc, uint32_t nAcc,
                                 const uint32_t *__restrict__ pairQ, const uint32_t *__restrict__ pairT, spb_aln *__restrict__ rows,
                                 const uint64_t *__restrict__ toff, const uint32_t *__restrict__ tkeys, RowKey *__restrict__ keys,
                                 const uint64_t *__restrict__ qoff, int seqIdMode, int swMode) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nAcc) return;
    const uint32_t pair = acc[k];
    const PairOut o = out[pair];
    uint32_t tlen;
    {
        const uint32_t t = pairT[pair];
        RowKey rk;
        rk.eval = o.eval;
        rk.bits = o.bits;
        rk.tlen = tlen = (uint32_t)min((unsigned long long)(toff[t + 1] - toff[t]), 0xFFFFFFFFull);
        rk.tkey = tkeys[t];
        rk.pad = 0;
        keys[k] = rk;
    }
    spb_aln r;
    r.query = pairQ[pair];
    r.target = pairT[pair];
    r.score = o.score;
    r.bits = o.bits;
    r.eval = o.eval;
    r.q_start = o.qStart; r.q_end = o.qEnd; r.t_start = o.tStart; r.t_end = o.tEnd;
    r.n_ident = o.nIdent;
    r.aln_len = o.alnLen;
    r.seq_id = seqIdMode == 0 ? (float)o.nIdent / (float)o.alnLen
                              : seq_id_of(seqIdMode, o.nIdent, (int)(qoff[r.query + 1] - qoff[r.query]), (int)tlen, o.alnLen);
    r.cigar_off = o.cigOff;
    r.cigar_len = o.cigLen;
    r.reserved = 0;
    if (swMode != 0) {       // no backtrace was computed: estimated identity, span as length, no CIGAR (traceback fields are stale)
        const ModeRow m = mode_row(o, swMode, (unsigned)(qoff[r.query + 1] - qoff[r.query]), tlen);
        r.q_start = m.qStart; r.t_start = m.tStart;
        r.n_ident = 0;
        r.aln_len = m.alnLen;
        r.seq_id = m.seqId;
        r.cigar_off = 0;
        r.cigar_len = 0;
    }
    rows[k] = r;
}

// rows per query (accepted pairs arrive grouped by query, so row k of the packed array already lies in its query's segment)
__global__ void row_count_kernel(const uint32_t *__restrict__ acc, uint32_t nAcc, const uint32_t *__restrict__ pairQ,
                                 uint32_t *__restrict__ cnt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nAcc) atomicAdd(&cnt[pairQ[acc[k]]], 1u);
}

// Per query order of Matcher::compareHits (M/alignment/Matcher.h:156-167): E-value asc (double), score desc, target length asc,
// target key asc.  One warp per query, rank sort from the packed rows into the output array.
__device__ __forceinline__ bool row_before(double ea, int ba, uint64_t la, uint32_t ka, double eb, int bb, uint64_t lb, uint32_t kb) {
    if (ea != eb) return ea < eb;
    if (ba != bb) return ba > bb;
    if (la != lb) return la < lb;
    return ka < kb;
}

// Rank sort, one THREAD per row: the rank of a row is the number of rows of its query that sort before it (rows with equal
// keys -- the same target listed twice for a query -- keep their input order).  A query has 3.7 rows on average, so a thread
// reads a few 24-byte keys of neighbouring threads (L1) and copies its 64-byte row to its place; no per-query warp, no idle
// lanes for the queries without rows.  (The warp-per-query kernel below is kept for lists beyond RR_MAX rows.)
#define RR_MAX 1024
__global__ void __launch_bounds__(256) rank_rows_kernel(const spb_aln *__restrict__ in, const RowKey *__restrict__ keys, uint32_t nAcc,
                                                        const uint64_t *__restrict__ rowOff, spb_aln *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nAcc) return;
    const spb_aln r = in[k];
    const uint64_t o = rowOff[r.query];
    const uint32_t n = (uint32_t)(rowOff[r.query + 1] - o);
    if (n > RR_MAX) return;                 // sort_rows_kernel takes the long lists
    const RowKey me = keys[k];
    const uint32_t i = (uint32_t)(k - o);
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) {
        const RowKey x = keys[o + j];
        const bool before = row_before(x.eval, x.bits, x.tlen, x.tkey, me.eval, me.bits, me.tlen, me.tkey);
        const bool equal = !before && !row_before(me.eval, me.bits, me.tlen, me.tkey, x.eval, x.bits, x.tlen, x.tkey);
        rank += (before || (equal && j < i)) ? 1u : 0u;
    }
    out[o + rank] = r;
}

// One warp per query, rank sort from the packed rows into the output array.  The four sort fields of a query's rows are
// staged in shared memory once (up to SR_CAP rows, i.e. any list --max-seqs 300 can produce); longer lists read them
// from global memory.
#define SR_CAP 320
#define SR_WARPS 4
__global__ void __launch_bounds__(SR_WARPS * 32) sort_rows_kernel(const spb_aln *__restrict__ in, const uint64_t *__restrict__ rowOff, uint32_t Q,
                                 const uint64_t *__restrict__ toff, const uint32_t *__restrict__ tkeys, spb_aln *__restrict__ out,
                                 uint32_t minRows) {
    __shared__ double s_e[SR_WARPS][SR_CAP];
    __shared__ int s_b[SR_WARPS][SR_CAP];
    __shared__ uint32_t s_l[SR_WARPS][SR_CAP], s_k[SR_WARPS][SR_CAP];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * SR_WARPS + wid;
    if (q >= Q) return;
    const uint64_t o = rowOff[q];
    const uint32_t n = (uint32_t)(rowOff[q + 1] - o);
    if (n == 0 || n < minRows) return;
    if (n == 1) {
        if (lane == 0) out[o] = in[o];
        return;
    }
    if (n <= SR_CAP) {
        for (uint32_t i = lane; i < n; i += 32) {
            const spb_aln &x = in[o + i];
            const uint32_t t = x.target;
            s_e[wid][i] = x.eval;
            s_b[wid][i] = x.bits;
            s_l[wid][i] = (uint32_t)min((unsigned long long)(toff[t + 1] - toff[t]), 0xFFFFFFFFull);
            s_k[wid][i] = tkeys[t];
        }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const double ei = s_e[wid][i];
            const int bi = s_b[wid][i];
            const uint32_t li = s_l[wid][i], ki = s_k[wid][i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; j++) {
                const double ej = s_e[wid][j];
                const int bj = s_b[wid][j];
                const uint32_t lj = s_l[wid][j], kj = s_k[wid][j];
                // equal keys (the same target listed twice for a query) keep their input order
                const bool before = row_before(ej, bj, lj, kj, ei, bi, li, ki);
                const bool equal = !before && !row_before(ei, bi, li, ki, ej, bj, lj, kj);
                rank += (before || (equal && j < i)) ? 1u : 0u;
            }
            out[o + rank] = in[o + i];
        }
        return;
    }
    for (uint32_t i = lane; i < n; i += 32) {
        const spb_aln r = in[o + i];
        const uint64_t li = toff[r.target + 1] - toff[r.target];
        const uint32_t ki = tkeys[r.target];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; j++) {
            if (j == i) continue;
            const spb_aln &x = in[o + j];
            const uint32_t tj = x.target;
            const uint64_t lj = toff[tj + 1] - toff[tj];
            const uint32_t kj = tkeys[tj];
            if (row_before(x.eval, x.bits, lj, kj, r.eval, r.bits, li, ki) ||
                (j < i && !row_before(r.eval, r.bits, li, ki, x.eval, x.bits, lj, kj))) rank++;
        }
        out[o + rank] = r;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Exhaustive search (BASELINE config 5): every query frame against EVERY target, no k-mer stage.
// The reference equivalent is `align` fed with a prefilter DB that lists all targets for every query; its result is
// the set of pairs whose Smith-Waterman E-value passes -e.  Here a streaming kernel scores all Q x T pairs and keeps
// the (query, target) pairs whose score reaches the query's minimal accepted score; that pair list then goes through
// the ordinary alignment stage (coordinates, reverse pass, traceback, rows), so the rows are those of the list path.
//
// sw_exhaustive_kernel<ROWS>: a thread owns TWO queries (int16 halves, as sw_fast_kernel) of at most ROWS residues and
// keeps their DP columns in registers; all threads of a CTA walk the same chunk of targets, so the target residue is
// uniform over the CTA (one broadcast load) and only the profile rows come from shared memory (conflict-free
// LDS.128).  No end-coordinate bookkeeping: 7 packed instructions per row (2 cells).  Queries are grouped by length
// (tiles of 128, longest first) and each tile runs the instantiation with the fewest rows that holds it.
// ---------------------------------------------------------------------------------------------------------
#define EXH_TILE (2 * SWF_THREADS)
struct ExhArgs {
    const uint32_t *qlist;       // fast queries, longest first, padded to EXH_TILE with 0xFFFFFFFF
    const uint4 *prof16;
    const uint16_t *minScore;    // per query: smallest accepted raw score (0xFFFF: none)
    const uint8_t *tres;
    const uint64_t *toff;
    const uint32_t *chunkStart;  // [nChunks + 1] target index boundaries
    uint32_t nChunks;
    uint32_t tile0, nTiles;      // tiles of this launch
    unsigned long long *cand;    // out: query << 32 | target
    unsigned long long candCap;
    unsigned long long *nCand;
    uint32_t *qCount;            // out: candidates per query
    int go, ge;
};

template <int ROWS>
__global__ void __launch_bounds__(SWF_THREADS) sw_exhaustive_kernel(ExhArgs A) {
    __shared__ uint4 s_prof[2][SWF_PROF_ROWS][SWF_THREADS];
    const int tid = threadIdx.x;
    const uint32_t tile = A.tile0 + blockIdx.x % A.nTiles, chunk = blockIdx.x / A.nTiles;
    const uint32_t qA = A.qlist[(size_t)tile * EXH_TILE + 2 * tid], qB = A.qlist[(size_t)tile * EXH_TILE + 2 * tid + 1];
    const uint4 none = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
#pragma unroll
    for (int a = 0; a < SWF_PROF_ROWS; a++) {
        s_prof[0][a][tid] = qA != 0xFFFFFFFFu ? __ldg(A.prof16 + (size_t)qA * SWF_PROF_ROWS + a) : none;
        s_prof[1][a][tid] = qB != 0xFFFFFFFFu ? __ldg(A.prof16 + (size_t)qB * SWF_PROF_ROWS + a) : none;
    }
    const uint32_t minA = qA != 0xFFFFFFFFu ? A.minScore[qA] : 0xFFFFu, minB = qB != 0xFFFFFFFFu ? A.minScore[qB] : 0xFFFFu;
    const uint32_t MGO = ((uint32_t)(-A.go) & 0xFFFFu) * 0x10001u, MGE = ((uint32_t)(-A.ge) & 0xFFFFu) * 0x10001u;
    const uint32_t tBegin = A.chunkStart[chunk], tEnd = A.chunkStart[chunk + 1];
    uint64_t to = A.toff[tBegin];
    for (uint32_t t = tBegin; t < tEnd; t++) {
        const uint64_t tn = A.toff[t + 1];
        const int len = (int)(tn - to);
        const uint8_t *tp = A.tres + to;
        to = tn;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(tp) & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(tp) & 3) * 8u;
        uint32_t cur = __ldg(w);
        uint32_t H[ROWS], E[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) { H[i] = 0; E[i] = 0; }
        uint32_t best = 0;
        for (int j0 = 0; j0 < len; j0 += 4) {
            const uint32_t nx = __ldg(w + (j0 >> 2) + 1);
            uint32_t res = __funnelshift_r(cur, nx, sh);
            cur = nx;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                uint32_t a = res & 0xFFu;
                res >>= 8;
                if (j0 + jj >= len || a > SPB_X) a = SWF_PAD;
                const uint4 pa = s_prof[0][a][tid], pb = s_prof[1][a][tid];
                const uint32_t paw[4] = {pa.x, pa.y, pa.z, pa.w}, pbw[4] = {pb.x, pb.y, pb.z, pb.w};
                uint32_t Hd = 0, F = 0;
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    const int r = i & 3;
                    const uint32_t sel = (uint32_t)r | ((uint32_t)(r | 8) << 4) | ((uint32_t)(4 + r) << 8) | ((uint32_t)((4 + r) | 8) << 12);
                    const uint32_t sc = prmt(paw[i >> 2], pbw[i >> 2], sel);
                    const uint32_t h1 = __viaddmax_s16x2_relu(Hd, sc, E[i]);
                    const uint32_t open = __vadd2(h1, MGO);
                    E[i] = __viaddmax_s16x2_relu(E[i], MGE, open);
                    const uint32_t h = __vmaxs2(h1, F);
                    F = __viaddmax_s16x2_relu(F, MGE, open);
                    Hd = H[i];
                    H[i] = h;
                    best = __vmaxs2(best, h);
                }
            }
        }
        const bool okA = (best & 0xFFFFu) >= minA, okB = (best >> 16) >= minB;
        if (okA | okB) {
            if (okA) {
                const unsigned long long k = atomicAdd(A.nCand, 1ull);
                if (k < A.candCap) A.cand[k] = ((unsigned long long)qA << 32) | t;
                atomicAdd(A.qCount + qA, 1u);
            }
            if (okB) {
                const unsigned long long k = atomicAdd(A.nCand, 1ull);
                if (k < A.candCap) A.cand[k] = ((unsigned long long)qB << 32) | t;
                atomicAdd(A.qCount + qB, 1u);
            }
        }
    }
}

// smallest raw score whose E-value passes (the decision of sw_finish_pair: score > 0 and !(eval > thr)); every score
// below it is rejected by construction, scores above it are re-checked by the alignment stage
__global__ void exh_min_score_kernel(const uint64_t *__restrict__ qoff, uint32_t Q, spb_gumbel g, double evalThr, double dbRes,
                                     int maxScore, uint16_t *__restrict__ minScore) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const double L = (double)(qoff[q + 1] - qoff[q]);
    int s = 1;
    for (; s <= maxScore; s++)
        if (!(dev_evalue(g, (double)s, L, dbRes) > evalThr)) break;
    minScore[q] = s <= maxScore ? (uint16_t)s : (uint16_t)0xFFFFu;
}

// pairs per query: accepted candidates of a register-kernel query, all targets otherwise
__global__ void exh_pair_count_kernel(const uint32_t *__restrict__ qCount, const uint8_t *__restrict__ qfast, uint32_t Q, uint32_t T,
                                      uint64_t *__restrict__ candCnt, uint64_t *__restrict__ pairCnt) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > Q) return;
    const uint64_t c = q < Q ? qCount[q] : 0;
    candCnt[q] = c;
    pairCnt[q] = q < Q ? (qfast[q] ? c : (uint64_t)T) : 0;
}

__global__ void exh_scatter_kernel(const unsigned long long *__restrict__ cand, uint64_t n, const uint64_t *__restrict__ candOff,
                                   const uint64_t *__restrict__ pairOff, uint32_t *__restrict__ pairT) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned long long key = cand[k];
    const uint32_t q = (uint32_t)(key >> 32);
    pairT[pairOff[q] + (k - candOff[q])] = (uint32_t)key;
}

__global__ void exh_fill_all_kernel(const uint32_t *__restrict__ slow, uint32_t nSlow, uint32_t T, const uint64_t *__restrict__ pairOff,
                                    uint32_t *__restrict__ pairT) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (uint64_t)nSlow * T) return;
    const uint32_t s = (uint32_t)(k / T), t = (uint32_t)(k - (uint64_t)s * T);
    pairT[pairOff[slow[s]] + t] = t;
}

}  // namespace spb

using namespace spb;

extern "C" int spb_align_set_targets(spb_ctx *c, const uint8_t *residues, const uint64_t *offsets, uint32_t T,
                                     const uint32_t *keys) {
    SPB_CHECK(!c || !c->is_lane, "spb_align_set_targets: a lane context borrows its parent's index / targets");
    SPB_CHECK(c && residues && offsets, "spb_align_set_targets: NULL argument");
    SPB_CUDA(cudaSetDevice(c->device));
    const uint64_t N = offsets[T];
    if (c->d_ares.ensure(N + 64) || c->d_aoff.ensure(((size_t)T + 1) * 8) || c->d_akeys.ensure((size_t)T * 4 + 4)) return 1;
    SPB_CUDA(cudaMemcpyAsync(c->d_ares.p, residues, N, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemsetAsync((uint8_t *)c->d_ares.p + N, SPB_X, 64, c->stream));   // the SW kernels read aligned 4-byte windows past the end
    SPB_CUDA(cudaMemcpyAsync(c->d_aoff.p, offsets, ((size_t)T + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    c->aT = T;
    c->a_res = N;
    c->h_aoff.assign(offsets, offsets + T + 1);
    c->a_maxlen = 0;
    for (uint32_t i = 0; i < T; i++) c->a_maxlen = std::max<uint64_t>(c->a_maxlen, offsets[i + 1] - offsets[i]);
    c->h_akeys.resize(T);
    for (uint32_t i = 0; i < T; i++) c->h_akeys[i] = keys ? keys[i] : i;
    if (T) SPB_CUDA(cudaMemcpy(c->d_akeys.p, c->h_akeys.data(), (size_t)T * 4, cudaMemcpyHostToDevice));
    c->have_aln_targets = true;
    c->have_aln = false;
    return 0;
}

// results are boxed: the public struct first, then the context whose pinned staging blocks it points into
struct AlignResultBox {
    spb_align_result pub;
    spb_ctx *owner;
    cudaEvent_t done;      // spb_align_fetch_begin: recorded behind the D2H copies on the copy stream (nullptr: synchronous fetch)
};

extern "C" void spb_free_align_result(spb_align_result *r) {
    if (!r) return;
    AlignResultBox *box = reinterpret_cast<AlignResultBox *>(r);
    if (box->owner) {
        box->owner->pin_release(r->offsets);
        box->owner->pin_release(r->rows);
        box->owner->pin_release(r->cigar_ops);
    }
    if (box->done) {
        cudaEventSynchronize(box->done);     // the pinned blocks must not be handed out again under a running copy
        cudaEventDestroy(box->done);
    }
    free(box);
}

// pairs resident in c->d_pairoff (u64[Q+1]) / c->d_pairT (u32[n_pairs]); results stay in c->d_pairout / d_acc / d_cig
static int align_pipeline(spb_ctx *c, const spb_align_params *p) {
    SPB_CHECK(c->have_matrices, "spb_align: call spb_set_matrices first");
    SPB_CHECK(c->have_aln_targets, "spb_align: call spb_align_set_targets first");
    SPB_CHECK(c->have_queries && c->have_pairs, "spb_align: no resident queries / pairs");
    SPB_CHECK(p->gap_open >= 0 && p->gap_open <= 255 && p->gap_extend >= 0 && p->gap_extend <= 255, "spb_align: gap costs out of range");
    SPB_CHECK(p->byte_lanes > 0 && p->word_lanes > 0, "spb_align: lanes must be positive");
    SPB_CHECK(p->cov_mode >= 0 && p->cov_mode <= 5 && p->seq_id_mode >= 0 && p->seq_id_mode <= 2 && p->sw_mode >= 0 && p->sw_mode <= 2,
              "spb_align: cov_mode must be 0..5, seq_id_mode 0..2 and sw_mode 0..2");
    const int swMode = p->sw_mode;
    cudaStream_t st = c->stream;
    const uint32_t Q = c->Q;
    const uint64_t NP = c->n_pairs;
    c->have_aln = false;
    memset(&c->aln_stats, 0, sizeof(c->aln_stats));
    c->aln_stats.n_queries = Q;
    c->aln_stats.n_pairs = NP;
    c->n_acc = 0;
    c->n_cigar = 0;
    if (NP == 0) {
        if (c->d_rowoff.ensure(((size_t)Q + 1) * 8 + 64)) return 1;
        SPB_CUDA(cudaMemsetAsync(c->d_rowoff.p, 0, ((size_t)Q + 1) * 8, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        c->have_aln = true;
        return 0;
    }
    const uint32_t maxL = c->q_maxlen;
    const int Lpad = (int)((maxL + 3) & ~3u);
    const size_t perWarp = (size_t)SPB_ALPHABET * Lpad + (size_t)Lpad * 32 * sizeof(short2);
    SPB_CHECK(maxL > 0, "spb_align: empty query batch");
    const int warpsPerBlock = (int)std::min<size_t>(4, std::max<size_t>(1, (96 * 1024) / perWarp));
    const size_t smem = perWarp * warpsPerBlock;
    SPB_CUDA(cudaEventRecord(c->ev2, st));

    // scratch: s_a cb i8[q_res], s_b qbias int[Q], s_c itemCnt/itemBase u32[Q+1] x2, s_d items, s_e accFlag u8[NP], s_f counters,
    //          s_l prof16 [Q][22][16] i8, s_m qfast u8[Q], s_n/s_o sort keys u32[NP] x2, s_p/s_q sort values u32[NP] x2
    if (c->s_a.ensure(c->q_res + 64) || c->s_b.ensure((size_t)Q * 4 + 4) || c->s_c.ensure(((size_t)Q + 1) * 8) ||
        c->s_e.ensure(NP + 64) || c->s_f.ensure(128) || c->d_pairQ.ensure(NP * 4) || c->d_pairout.ensure(NP * sizeof(PairOut)) ||
        c->d_acc.ensure(NP * 4) || c->s_l.ensure((size_t)Q * SWF_PROF_ROWS * SWF_ROWS + 64) || c->s_m.ensure((size_t)Q + 64) ||
        c->s_n.ensure(NP * 4 + 64) || c->s_o.ensure(NP * 4 + 64) || c->s_p.ensure(NP * 4 + 64) || c->s_q.ensure(NP * 4 + 64)) return 1;
    SPB_CUDA(cudaMemsetAsync(c->s_f.p, 0, 128, st));
    SPB_CUDA(cudaMemsetAsync(c->s_e.p, 0, NP, st));
    // register kernel eligibility: segLen == 1 in both emulated SIMD kernels and at most 16 rows
    const int fastMaxL = std::min(SWF_ROWS, std::min(p->byte_lanes, p->word_lanes));
    int matMin = 0;
    for (int i = 0; i < SPB_ALPHABET * SPB_ALPHABET; i++) matMin = std::min(matMin, (int)(signed char)c->hm.align[i]);
    align_profile_kernel<<<(Q + 3) / 4, 128, 0, st>>>(c->d_q.as<uint8_t>(), c->d_qoff64.as<uint64_t>(), Q, c->d_mat.as<Matrices>(), p->comp_bias,
                                           matMin, c->s_a.as<signed char>(), c->s_b.as<int>(), c->s_l.as<signed char>(), c->s_m.as<uint8_t>(),
                                           fastMaxL);
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    pair_key_kernel<<<(Q + 3) / 4, 128, 0, st>>>(c->d_pairoff.as<uint64_t>(), c->d_pairT.as<uint32_t>(), c->d_aoff.as<uint64_t>(),
                                                 c->s_m.as<uint8_t>(), Q, c->d_pairQ.as<uint32_t>(), c->s_n.as<uint32_t>(),
                                                 c->s_p.as<uint32_t>());
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    size_t tmpBytes = 0;
    int keyBits = 1;
    while (keyBits < 32 && ((uint64_t)1 << keyBits) <= c->a_maxlen + 1) keyBits++;
    SPB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmpBytes, c->s_n.as<uint32_t>(), c->s_o.as<uint32_t>(), c->s_p.as<uint32_t>(),
                                                       c->s_q.as<uint32_t>(), (int)NP, 0, keyBits, st));
    if (c->s_cub.ensure(tmpBytes)) return 1;
    SPB_CUDA(cub::DeviceRadixSort::SortPairsDescending(c->s_cub.p, tmpBytes, c->s_n.as<uint32_t>(), c->s_o.as<uint32_t>(), c->s_p.as<uint32_t>(),
                                                       c->s_q.as<uint32_t>(), (int)NP, 0, keyBits, st));
    // every query of the batch runs the register kernel (what a spacer batch looks like): no work items of the shared-memory
    // kernel to count, no H/E slices for the reverse pass -- and no host round trip to learn that
    const bool allFast = c->q_minlen >= 1 && c->q_maxlen <= (uint32_t)fastMaxL;
    uint32_t *itemCnt = c->s_c.as<uint32_t>(), *itemBase = itemCnt + (Q + 1);
    uint32_t nItems = 0;
    if (!allFast) {
        item_count_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(c->d_pairoff.as<uint64_t>(), c->s_m.as<uint8_t>(), Q, itemCnt);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, itemCnt, itemBase, (int)(Q + 1), st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, itemCnt, itemBase, (int)(Q + 1), st));
        SPB_CUDA(cudaMemcpyAsync(&nItems, itemBase + Q, 4, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
    }

    FastArgs F;
    memset(&F, 0, sizeof(F));
    F.sortedKey = c->s_o.as<uint32_t>();
    F.sortedPair = c->s_q.as<uint32_t>();
    F.NP = (uint32_t)NP;
    F.pairQ = c->d_pairQ.as<uint32_t>();
    F.pairT = c->d_pairT.as<uint32_t>();
    F.qoff = c->d_qoff64.as<uint64_t>();
    F.qbias = c->s_b.as<int>();
    F.prof16 = c->s_l.as<uint4>();
    F.tres = c->d_ares.as<uint8_t>();
    F.toff = c->d_aoff.as<uint64_t>();
    F.out = c->d_pairout.as<PairOut>();
    F.accFlag = c->s_e.as<uint8_t>();
    F.go = p->gap_open;
    F.ge = p->gap_extend;
    F.keyMul = 16u;
    F.evalThr = p->eval_thr;
    F.dbRes = (double)p->db_res_count;
    F.g = p->gumbel;
    F.cells = c->s_f.as<unsigned long long>();

    AlnArgs A;
    memset(&A, 0, sizeof(A));
    A.q = c->d_q.as<uint8_t>();
    A.qoff = c->d_qoff64.as<uint64_t>();
    A.qcb = c->s_a.as<signed char>();
    A.qbias = c->s_b.as<int>();
    A.tres = c->d_ares.as<uint8_t>();
    A.toff = c->d_aoff.as<uint64_t>();
    A.mat = c->d_mat.as<Matrices>();
    A.items = nullptr;
    A.nItems = nItems;
    A.pairT = c->d_pairT.as<uint32_t>();
    A.pairQ = c->d_pairQ.as<uint32_t>();
    A.accFlag = c->s_e.as<uint8_t>();
    A.out = c->d_pairout.as<PairOut>();
    A.go = p->gap_open;
    A.ge = p->gap_extend;
    A.byteLanes = p->byte_lanes;
    A.wordLanes = p->word_lanes;
    A.evalThr = p->eval_thr;
    A.dbRes = (double)p->db_res_count;
    A.g = p->gumbel;
    A.Lpad = Lpad;
    A.cells = c->s_f.as<unsigned long long>();
    SPB_CUDA(cudaEventRecord(c->ev0, st));
    {
        const uint32_t nThreads = (uint32_t)((NP + 1) / 2);
        sw_fast_kernel<<<(nThreads + SWF_THREADS - 1) / SWF_THREADS, SWF_THREADS, 0, st>>>(F);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    if (nItems) {
        // queries longer than the register kernel's 16 rows: one warp per (query, <= 32 targets), H/E in shared memory
        SPB_CHECK(perWarp <= 200 * 1024, "spb_align: query length %u outside the supported range of the shared-memory kernel (1..%d)",
                  maxL, (int)(200 * 1024 / (SPB_ALPHABET + 128)));
        if (c->s_d.ensure((size_t)nItems * sizeof(FwdItem) + 16)) return 1;
        item_fill_kernel<<<(Q + 255) / 256, 256, 0, st>>>(c->d_pairoff.as<uint64_t>(), itemBase, Q, c->s_d.as<FwdItem>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        A.items = c->s_d.as<FwdItem>();
        SPB_CUDA(cudaFuncSetAttribute(sw_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sw_forward_kernel<<<(nItems + warpsPerBlock - 1) / warpsPerBlock, 32 * warpsPerBlock, smem, st>>>(A);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    SPB_CUDA(cudaEventRecord(c->ev1, st));

    // ---- accepted pairs
    unsigned long long *d_ctr = c->s_f.as<unsigned long long>();   // [0] cells, [1] nAcc (as u32), [2] error
    uint32_t *d_nAcc = reinterpret_cast<uint32_t *>(d_ctr + 1);
    int *d_err = reinterpret_cast<int *>(d_ctr + 2);
    cub::CountingInputIterator<uint32_t> cnt_it(0);
    tmpBytes = 0;
    SPB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmpBytes, cnt_it, c->s_e.as<uint8_t>(), c->d_acc.as<uint32_t>(), d_nAcc, (int)NP, st));
    if (c->s_cub.ensure(tmpBytes)) return 1;
    SPB_CUDA(cub::DeviceSelect::Flagged(c->s_cub.p, tmpBytes, cnt_it, c->s_e.as<uint8_t>(), c->d_acc.as<uint32_t>(), d_nAcc, (int)NP, st));
    unsigned long long h_ctr[3];
    SPB_CUDA(cudaMemcpyAsync(h_ctr, d_ctr, 24, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->aln_stats.kernel_ms = ms;
    c->aln_stats.n_cells = h_ctr[0];
    const uint32_t nAcc = (uint32_t)(h_ctr[1] & 0xFFFFFFFFull);
    if (nAcc && swMode != 2) {      // (score only: neither the reverse pass nor the traceback)
        // scratch: s_g he offsets u64[nAcc+1] (+ sizes), s_h he scratch, s_i tb offsets, s_j cig offsets, s_k tb scratch
        if (c->s_g.ensure(((size_t)nAcc + 1) * 16)) return 1;
        uint64_t *heSz = c->s_g.as<uint64_t>(), *heOff = heSz + (nAcc + 1);
        uint64_t heTotal = 0;
        if (!allFast) {      // H/E slices of the reverse pass: only queries outside the register kernel use them
            he_size_kernel<<<(nAcc + 1 + 255) / 256, 256, 0, st>>>(A.out, c->d_acc.as<uint32_t>(), nAcc, heSz);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, heSz, heOff, (int)(nAcc + 1), st));
            if (c->s_cub.ensure(tmpBytes)) return 1;
            SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, heSz, heOff, (int)(nAcc + 1), st));
            SPB_CUDA(cudaMemcpyAsync(&heTotal, heOff + nAcc, 8, cudaMemcpyDeviceToHost, st));
            SPB_CUDA(cudaStreamSynchronize(st));
        }
        if (c->s_h.ensure(heTotal * sizeof(short2) + 16)) return 1;
        TbArgs B;
        memset(&B, 0, sizeof(B));
        B.q = A.q; B.qoff = A.qoff; B.qcb = A.qcb; B.tres = A.tres; B.toff = A.toff; B.mat = A.mat;
        B.acc = c->d_acc.as<uint32_t>();
        B.nAcc = nAcc;
        B.pairQ = A.pairQ;
        B.pairT = A.pairT;
        B.out = A.out;
        B.go = A.go; B.ge = A.ge; B.byteLanes = A.byteLanes; B.wordLanes = A.wordLanes;
        B.heScratch = c->s_h.as<short2>();
        B.heOff = heOff;
        B.qfast = c->s_m.as<uint8_t>();
        B.prof16 = c->s_l.as<uint4>();
        sw_reverse_kernel<<<(nAcc + 127) / 128, 128, 0, st>>>(B);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        if (c->s_i.ensure(((size_t)nAcc + 1) * 16) || c->s_j.ensure(((size_t)nAcc + 1) * 16)) return 1;
        uint64_t *tbSz = c->s_i.as<uint64_t>(), *tbOff = tbSz + (nAcc + 1);
        uint64_t *cigSz = c->s_j.as<uint64_t>(), *cigOff = cigSz + (nAcc + 1);
        tb_size_kernel<<<(nAcc + 1 + 255) / 256, 256, 0, st>>>(A.out, c->d_acc.as<uint32_t>(), nAcc, tbSz, cigSz, d_err);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, tbSz, tbOff, (int)(nAcc + 1), st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, tbSz, tbOff, (int)(nAcc + 1), st));
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, cigSz, cigOff, (int)(nAcc + 1), st));
        uint64_t tbTotal = 0, cigTotal = 0;
        int h_err = 0;
        SPB_CUDA(cudaMemcpyAsync(&tbTotal, tbOff + nAcc, 8, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaMemcpyAsync(&cigTotal, cigOff + nAcc, 8, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        SPB_CHECK(h_err == 0, "spb_align: forward and reverse Smith-Waterman scores differ (accepted pair #%d)", h_err - 1);
        if (swMode == 0) {          // (score + coverage stops at the start positions)
            if (c->s_k.ensure(tbTotal + 64) || c->d_cig.ensure(cigTotal + 64)) return 1;
            // slices are sized by an upper bound of the alignment length: the slack between them is copied out with the result
            SPB_CUDA(cudaMemsetAsync(c->d_cig.p, 0, cigTotal, st));
            B.tbOff = tbOff;
            B.tbScratch = c->s_k.as<unsigned char>();
            B.cigOff = cigOff;
            B.cigar = c->d_cig.as<char>();
            sw_traceback_kernel<<<(nAcc + TB_THREADS - 1) / TB_THREADS, TB_THREADS, 0, st>>>(B);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            c->n_cigar = cigTotal;
        }
    }
    // ---- -c / --cov-mode / --min-seq-id / --min-aln-len: the accepted list shrinks to the pairs that pass (their CIGAR slices
    // stay where they are; the slices of dropped pairs become slack of the CIGAR buffer).  Not on the predictmatch path.
    uint32_t nRows = nAcc;
    const uint32_t *accRows = c->d_acc.as<uint32_t>();
    if (nAcc && (p->cov_thr != 0.0f || p->cov_mode != 0 || p->seq_id_thr != 0.0f || p->aln_len_thr != 0)) {
        Criteria cr;
        cr.covThr = p->cov_thr; cr.covMode = p->cov_mode; cr.seqIdThr = p->seq_id_thr; cr.alnLenThr = p->aln_len_thr;
        cr.seqIdMode = p->seq_id_mode;
        cr.swMode = swMode;
        uint8_t *keep = c->s_e.as<uint8_t>();            // (the E-value flags are dead: d_acc was selected from them)
        uint32_t *kept = c->s_q.as<uint32_t>();          // (the sorted pair list of the forward kernel is dead too)
        criteria_kernel<<<(nAcc + 255) / 256, 256, 0, st>>>(c->d_pairout.as<PairOut>(), accRows, nAcc, c->d_pairQ.as<uint32_t>(),
                                                            c->d_pairT.as<uint32_t>(), c->d_qoff64.as<uint64_t>(), c->d_aoff.as<uint64_t>(),
                                                            cr, keep);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmpBytes, accRows, keep, kept, d_nAcc, (int)nAcc, st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceSelect::Flagged(c->s_cub.p, tmpBytes, accRows, keep, kept, d_nAcc, (int)nAcc, st));
        SPB_CUDA(cudaMemcpyAsync(&nRows, d_nAcc, 4, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        accRows = kept;
    }
    // ---- --max-accept / --max-rejected: replay the reference's walk over each query's list on the verdicts (see accept_limit_kernel)
    const bool limAcc = p->max_accept > 0 && p->max_accept != INT32_MAX, limRej = p->max_rejected > 0 && p->max_rejected != INT32_MAX;
    if (limAcc || limRej) {
        SPB_CHECK(c->pair_rank_mode != 2, "spb_align: --max-accept / --max-rejected need a pair list with an order (not the exhaustive search)");
        uint8_t *final_ = c->s_e.as<uint8_t>();          // u8[NP] (E-value flags / criteria flags: both consumed by now)
        uint8_t *keep2 = c->s_n.as<uint8_t>();           // (sort keys of the forward kernel: dead)
        uint32_t *kept2 = c->s_o.as<uint32_t>();
        uint32_t *order = nullptr;
        SPB_CUDA(cudaMemsetAsync(final_, 0, NP, st));
        if (nRows) {
            set_flags_kernel<<<(nRows + 255) / 256, 256, 0, st>>>(accRows, nRows, final_);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
        }
        if (c->pair_rank_mode == 1) {
            order = c->s_p.as<uint32_t>();               // (unsorted sort values: dead)
            list_order_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, st>>>(c->d_pairoff.as<uint64_t>(), c->d_pairQ.as<uint32_t>(),
                                                                          c->d_pairRank.as<uint32_t>(), (uint32_t)NP, order);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
        }
        accept_limit_kernel<<<(Q + 127) / 128, 128, 0, st>>>(c->d_pairoff.as<uint64_t>(), Q, order,
                                                             limAcc ? (uint32_t)p->max_accept : 0xFFFFFFFFu,
                                                             limRej ? (uint32_t)p->max_rejected : 0xFFFFFFFFu, final_);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        if (nRows) {
            gather_flags_kernel<<<(nRows + 255) / 256, 256, 0, st>>>(accRows, nRows, final_, keep2);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            SPB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmpBytes, accRows, keep2, kept2, d_nAcc, (int)nRows, st));
            if (c->s_cub.ensure(tmpBytes)) return 1;
            SPB_CUDA(cub::DeviceSelect::Flagged(c->s_cub.p, tmpBytes, accRows, keep2, kept2, d_nAcc, (int)nRows, st));
            SPB_CUDA(cudaMemcpyAsync(&nRows, d_nAcc, 4, cudaMemcpyDeviceToHost, st));
            SPB_CUDA(cudaStreamSynchronize(st));
            accRows = kept2;
        }
    }
    // ---- result rows: per query in the order of Matcher::compareHits (part of the stage: `align` writes sorted entries)
    {
        // scratch: s_g packed rows, s_i row counts u32[Q+1] + u64[Q+1] widened
        if (c->s_g.ensure(std::max<size_t>(nRows, 1) * sizeof(spb_aln)) || c->d_rows.ensure(std::max<size_t>(nRows, 1) * sizeof(spb_aln)) ||
            c->s_i.ensure(((size_t)Q + 1) * 12 + 64) || c->d_rowoff.ensure(((size_t)Q + 1) * 8 + 64)) return 1;
        uint32_t *cnt32 = c->s_i.as<uint32_t>();
        uint64_t *cnt64 = reinterpret_cast<uint64_t *>(c->s_i.as<unsigned char>() + (((size_t)Q + 1) * 4 + 15) / 16 * 16);
        uint64_t *rowOff = c->d_rowoff.as<uint64_t>();
        SPB_CUDA(cudaMemsetAsync(cnt32, 0, ((size_t)Q + 1) * 4, st));
        if (nRows) {
            row_count_kernel<<<(nRows + 255) / 256, 256, 0, st>>>(accRows, nRows, c->d_pairQ.as<uint32_t>(), cnt32);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
        }
        count_widen_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(cnt32, Q, cnt64);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt64, rowOff, (int)(Q + 1), st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, cnt64, rowOff, (int)(Q + 1), st));
        if (nRows) {
            if (c->s_h.ensure((size_t)nRows * sizeof(RowKey) + 64)) return 1;     // (the H/E scratch of the reverse pass is dead by now)
            pack_rows_kernel<<<(nRows + 255) / 256, 256, 0, st>>>(c->d_pairout.as<PairOut>(), accRows, nRows,
                                                                 c->d_pairQ.as<uint32_t>(), c->d_pairT.as<uint32_t>(), c->s_g.as<spb_aln>(),
                                                                 c->d_aoff.as<uint64_t>(), c->d_akeys.as<uint32_t>(), c->s_h.as<RowKey>(),
                                                                 c->d_qoff64.as<uint64_t>(), p->seq_id_mode, swMode);
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            rank_rows_kernel<<<(nRows + 255) / 256, 256, 0, st>>>(c->s_g.as<spb_aln>(), c->s_h.as<RowKey>(), nRows, rowOff, c->d_rows.as<spb_aln>());
            SPB_CUDA(cudaGetLastError()); c->n_launches++;
            if (c->pairs_max_per_query > RR_MAX) {      // lists longer than the per-thread rank loop takes: one warp per query
                sort_rows_kernel<<<(Q + SR_WARPS - 1) / SR_WARPS, SR_WARPS * 32, 0, st>>>(c->s_g.as<spb_aln>(), rowOff, Q, c->d_aoff.as<uint64_t>(),
                                                                                         c->d_akeys.as<uint32_t>(), c->d_rows.as<spb_aln>(), RR_MAX + 1);
                SPB_CUDA(cudaGetLastError()); c->n_launches++;
            }
        }
    }
    SPB_CUDA(cudaEventRecord(c->ev3, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
    c->aln_stats.total_ms = ms;
    c->aln_stats.n_rows = nRows;
    c->n_acc = nRows;
    c->have_aln = true;
    return 0;
}

extern "C" int spb_pairs_set(spb_ctx *c, const uint64_t *pair_offsets, const uint32_t *pair_targets) {
    SPB_CHECK(c && pair_offsets, "spb_pairs_set: NULL argument");
    SPB_CHECK(c->have_queries, "spb_pairs_set: no resident query batch (spb_queries_set)");
    SPB_CHECK(c->have_aln_targets, "spb_pairs_set: call spb_align_set_targets first");
    SPB_CUDA(cudaSetDevice(c->device));
    const uint32_t Q = c->Q;
    const uint64_t NP = pair_offsets[Q];
    SPB_CHECK(NP < 0x7FFFFFF0ull, "spb_pairs_set: more than 2^31 pairs in one call; split the batch");
    SPB_CHECK(NP == 0 || pair_targets, "spb_pairs_set: NULL targets");
    // work order: per query, targets by length descending so that the lanes of a warp finish together
    // (the position in the caller's list travels with the pair: --max-accept / --max-rejected walk the list in its order)
    std::vector<uint32_t> pairT(NP), rank(NP);
    uint64_t maxPer = 0;
    for (uint32_t qi = 0; qi < Q; qi++) {
        const uint64_t a = pair_offsets[qi], b = pair_offsets[qi + 1];
        SPB_CHECK(b >= a && b <= NP, "spb_pairs_set: offsets not monotone at %u", qi);
        for (uint64_t k = a; k < b; k++) {
            SPB_CHECK(pair_targets[k] < c->aT, "spb_align: target id %u of query %u out of range (%u targets)", pair_targets[k], qi, c->aT);
            rank[k] = (uint32_t)(k - a);
        }
        maxPer = std::max<uint64_t>(maxPer, b - a);
        if (b - a > 1)
            std::stable_sort(rank.begin() + a, rank.begin() + b, [&](uint32_t i, uint32_t j) {
                const uint32_t x = pair_targets[a + i], y = pair_targets[a + j];
                return (c->h_aoff[x + 1] - c->h_aoff[x]) > (c->h_aoff[y + 1] - c->h_aoff[y]);
            });
        for (uint64_t k = a; k < b; k++) pairT[k] = pair_targets[a + rank[k]];
    }
    if (c->d_pairoff.ensure(((size_t)Q + 1) * 8) || c->d_pairT.ensure(NP * 4 + 16) || c->d_pairRank.ensure(NP * 4 + 16)) return 1;
    SPB_CUDA(cudaMemcpyAsync(c->d_pairoff.p, pair_offsets, ((size_t)Q + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (NP) SPB_CUDA(cudaMemcpyAsync(c->d_pairT.p, pairT.data(), NP * 4, cudaMemcpyHostToDevice, c->stream));
    if (NP) SPB_CUDA(cudaMemcpyAsync(c->d_pairRank.p, rank.data(), NP * 4, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    c->n_pairs = NP;
    c->pair_rank_mode = 1;
    c->pairs_max_per_query = maxPer;
    c->have_pairs = true;
    c->have_aln = false;
    return 0;
}

// pairs = the resident prefilter result (hit lists never leave HBM between the two stages)
extern "C" int spb_pairs_from_prefilter(spb_ctx *c) {
    SPB_CHECK(c, "spb_pairs_from_prefilter: NULL argument");
    SPB_CHECK(c->have_pref, "spb_pairs_from_prefilter: no resident prefilter result");
    SPB_CUDA(cudaSetDevice(c->device));
    const uint32_t Q = c->Q;
    const uint64_t NP = c->n_hits;
    SPB_CHECK(NP < 0x7FFFFFF0ull, "spb_pairs_from_prefilter: more than 2^31 pairs; split the batch");
    if (c->d_pairoff.ensure(((size_t)Q + 1) * 8) || c->d_pairT.ensure(NP * 4 + 16)) return 1;
    SPB_CUDA(cudaMemcpyAsync(c->d_pairoff.p, c->d_hitoff.p, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (NP) {
        hit_ids_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, c->stream>>>(c->d_hits.as<spb_hit>(), NP, c->d_pairT.as<uint32_t>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    c->n_pairs = NP;
    c->pair_rank_mode = 0;
    c->pairs_max_per_query = c->pref_max_seqs ? c->pref_max_seqs : NP;    // a hit list holds at most max-seqs targets
    c->have_pairs = true;
    c->have_aln = false;
    return 0;
}

extern "C" int spb_align_resident(spb_ctx *c, const spb_align_params *p) {
    SPB_CHECK(c && p, "spb_align_resident: NULL argument");
    SPB_CUDA(cudaSetDevice(c->device));
    return align_pipeline(c, p);
}

extern "C" int spb_align_stats(spb_ctx *c, spb_align_result *out) {
    SPB_CHECK(c && out, "spb_align_stats: NULL argument");
    SPB_CHECK(c->have_aln, "spb_align_stats: no resident alignment result");
    *out = c->aln_stats;
    out->offsets = nullptr;
    out->rows = nullptr;
    out->cigar_ops = nullptr;
    return 0;
}

extern "C" int spb_align_fetch(spb_ctx *c, spb_align_result **out) {
    SPB_CHECK(c && out, "spb_align_fetch: NULL argument");
    SPB_CHECK(c->have_aln, "spb_align_fetch: no resident alignment result");
    SPB_CUDA(cudaSetDevice(c->device));
    const auto t_begin = std::chrono::steady_clock::now();
    *out = nullptr;
    cudaStream_t st = c->stream;
    const uint32_t Q = c->Q;
    const uint32_t nAcc = (uint32_t)c->n_acc;
    AlignResultBox *box = (AlignResultBox *)calloc(1, sizeof(AlignResultBox));
    SPB_CHECK(box != nullptr, "spb_align_fetch: out of memory");
    box->owner = c;
    spb_align_result *res = &box->pub;
    *res = c->aln_stats;
    // results are staged in pinned blocks of the context (asynchronous, full-speed D2H); they stay valid until
    // spb_free_align_result -- free results before destroying the context
    res->offsets = (uint64_t *)c->pin_acquire(((size_t)Q + 1) * sizeof(uint64_t));
    res->rows = (spb_aln *)c->pin_acquire(std::max<size_t>(nAcc, 1) * sizeof(spb_aln));
    res->cigar_ops = (char *)c->pin_acquire(c->n_cigar + 1);
    if (!res->offsets || !res->rows || !res->cigar_ops) { spb_free_align_result(res); return 1; }
    res->n_cigar = c->n_cigar;
    res->n_rows = nAcc;
    const uint64_t *rowOff = c->d_rowoff.as<uint64_t>();
    cudaError_t e = cudaSuccess;
    if (nAcc) {
        e = cudaMemcpyAsync(res->rows, c->d_rows.p, (size_t)nAcc * sizeof(spb_aln), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && c->n_cigar) e = cudaMemcpyAsync(res->cigar_ops, c->d_cig.p, c->n_cigar, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(res->offsets, rowOff, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (getenv("SPB_DEBUG")) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        fprintf(stderr, "spb_align_fetch: %u rows, %llu CIGAR bytes, %.3f ms (D2H)\n", nAcc, (unsigned long long)c->n_cigar, ms);
    }
    if (e != cudaSuccess) {
        spb_free_align_result(res);
        spb::set_error("spb_align_fetch: %s", cudaGetErrorString(e));
        return 1;
    }
    *out = res;
    return 0;
}

// Pipelined fetch: the sorted rows, CIGARs and offsets of the resident result are copied device-to-device into staging
// buffers on the compute stream (microseconds) and from there to pinned host memory on a second stream, so the next batch
// (spb_queries_set .. spb_align_resident) runs while the previous result crosses PCIe.  One staging set: a second _begin
// makes the compute stream wait until the previous D2H has left the staging buffers (it has, long before the next batch
// is through its kernels).
extern "C" int spb_align_fetch_begin(spb_ctx *c, spb_align_result **out) {
    SPB_CHECK(c && out, "spb_align_fetch_begin: NULL argument");
    SPB_CHECK(c->have_aln, "spb_align_fetch_begin: no resident alignment result");
    SPB_CUDA(cudaSetDevice(c->device));
    *out = nullptr;
    if (!c->copy_stream) {
        SPB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        SPB_CUDA(cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
        SPB_CUDA(cudaEventCreateWithFlags(&c->ev_stage_free, cudaEventDisableTiming));
    }
    const uint32_t Q = c->Q;
    const size_t nAcc = (size_t)c->n_acc, nCig = (size_t)c->n_cigar;
    if (c->d_st_rows.ensure(std::max<size_t>(nAcc, 1) * sizeof(spb_aln)) || c->d_st_cig.ensure(nCig + 16) ||
        c->d_st_off.ensure(((size_t)Q + 1) * 8)) return 1;
    AlignResultBox *box = (AlignResultBox *)calloc(1, sizeof(AlignResultBox));
    SPB_CHECK(box != nullptr, "spb_align_fetch_begin: out of memory");
    box->owner = c;
    spb_align_result *res = &box->pub;
    *res = c->aln_stats;
    res->offsets = (uint64_t *)c->pin_acquire(((size_t)Q + 1) * sizeof(uint64_t));
    res->rows = (spb_aln *)c->pin_acquire(std::max<size_t>(nAcc, 1) * sizeof(spb_aln));
    res->cigar_ops = (char *)c->pin_acquire(nCig + 1);
    if (!res->offsets || !res->rows || !res->cigar_ops) { spb_free_align_result(res); return 1; }
    res->n_cigar = nCig;
    res->n_rows = nAcc;
    cudaStream_t st = c->stream, cs = c->copy_stream;
    cudaError_t e = cudaEventCreateWithFlags(&box->done, cudaEventDisableTiming);
    if (e == cudaSuccess && c->stage_busy) e = cudaStreamWaitEvent(st, c->ev_stage_free, 0);
    if (e == cudaSuccess && nAcc) e = cudaMemcpyAsync(c->d_st_rows.p, c->d_rows.p, nAcc * sizeof(spb_aln), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && nCig) e = cudaMemcpyAsync(c->d_st_cig.p, c->d_cig.p, nCig, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_st_off.p, c->d_rowoff.p, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_staged, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, c->ev_staged, 0);
    // in pieces: the control path of the next batch reads small counters back on the compute stream, and the D2H copy engine
    // arbitrates between streams per copy -- behind one 12 MB transfer each of those reads would wait for all of it
    auto d2h_chunked = [&](void *dst, const void *src, size_t bytes) {
        const size_t CH = 1u << 20;
        for (size_t o = 0; o < bytes && e == cudaSuccess; o += CH)
            e = cudaMemcpyAsync((char *)dst + o, (const char *)src + o, std::min(CH, bytes - o), cudaMemcpyDeviceToHost, cs);
    };
    if (e == cudaSuccess && nAcc) d2h_chunked(res->rows, c->d_st_rows.p, nAcc * sizeof(spb_aln));
    if (e == cudaSuccess && nCig) d2h_chunked(res->cigar_ops, c->d_st_cig.p, nCig);
    if (e == cudaSuccess) e = cudaMemcpyAsync(res->offsets, c->d_st_off.p, ((size_t)Q + 1) * 8, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_stage_free, cs);
    if (e == cudaSuccess) e = cudaEventRecord(box->done, cs);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(cs);
        spb_free_align_result(res);
        spb::set_error("spb_align_fetch_begin: %s", cudaGetErrorString(e));
        return 1;
    }
    c->stage_busy = true;
    *out = res;
    return 0;
}

extern "C" int spb_align_fetch_wait(spb_align_result *r) {
    SPB_CHECK(r != nullptr, "spb_align_fetch_wait: NULL argument");
    AlignResultBox *box = reinterpret_cast<AlignResultBox *>(r);
    if (box->done) SPB_CUDA(cudaEventSynchronize(box->done));
    return 0;
}

extern "C" int spb_align(spb_ctx *c, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                         const uint64_t *pair_offsets, const uint32_t *pair_targets, const spb_align_params *p,
                         spb_align_result **out) {
    SPB_CHECK(c && q_offsets && pair_offsets && p && out, "spb_align: NULL argument");
    *out = nullptr;
    if (spb_queries_set(c, q_residues, q_offsets, Q)) return 1;
    if (spb_pairs_set(c, pair_offsets, pair_targets)) return 1;
    if (spb_align_resident(c, p)) return 1;
    return spb_align_fetch(c, out);
}

// ---- exhaustive search: all resident queries x all alignment targets -> resident pair list ----
template <int ROWS>
static cudaError_t launch_exhaustive(const ExhArgs &A, uint32_t blocks, cudaStream_t st) {
    sw_exhaustive_kernel<ROWS><<<blocks, SWF_THREADS, 0, st>>>(A);
    return cudaGetLastError();
}

extern "C" int spb_pairs_exhaustive(spb_ctx *c, const spb_align_params *p, spb_exhaustive_stats *stats) {
    SPB_CHECK(c && p, "spb_pairs_exhaustive: NULL argument");
    SPB_CHECK(c->have_matrices, "spb_pairs_exhaustive: call spb_set_matrices first");
    SPB_CHECK(c->have_aln_targets, "spb_pairs_exhaustive: call spb_align_set_targets first");
    SPB_CHECK(c->have_queries, "spb_pairs_exhaustive: no resident query batch (spb_queries_set)");
    SPB_CHECK(p->gap_open >= 0 && p->gap_open <= 255 && p->gap_extend >= 0 && p->gap_extend <= 255, "spb_pairs_exhaustive: gap costs out of range");
    SPB_CHECK(p->byte_lanes > 0 && p->word_lanes > 0, "spb_pairs_exhaustive: lanes must be positive");
    SPB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint32_t Q = c->Q, T = c->aT;
    spb_exhaustive_stats es;
    memset(&es, 0, sizeof(es));
    es.n_queries = Q;
    es.n_targets = T;
    c->have_pairs = false;
    c->have_aln = false;
    if (c->d_pairoff.ensure(((size_t)Q + 1) * 8 + 64)) return 1;
    if (Q == 0 || T == 0) {
        SPB_CUDA(cudaMemsetAsync(c->d_pairoff.p, 0, ((size_t)Q + 1) * 8, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        if (c->d_pairT.ensure(16)) return 1;
        c->n_pairs = 0;
        c->have_pairs = true;
        if (stats) *stats = es;
        return 0;
    }
    SPB_CUDA(cudaEventRecord(c->ev2, st));
    // profiles and register-kernel eligibility exactly as the list path computes them
    const int fastMaxL = std::min(SWF_ROWS, std::min(p->byte_lanes, p->word_lanes));
    int matMin = 0, matMax = 0;
    for (int i = 0; i < SPB_ALPHABET * SPB_ALPHABET; i++) {
        matMin = std::min(matMin, (int)(signed char)c->hm.align[i]);
        matMax = std::max(matMax, (int)(signed char)c->hm.align[i]);
    }
    if (c->s_a.ensure(c->q_res + 64) || c->s_b.ensure((size_t)Q * 4 + 4) || c->s_l.ensure((size_t)Q * SWF_PROF_ROWS * SWF_ROWS + 64) ||
        c->s_m.ensure((size_t)Q + 64) || c->s_n.ensure((size_t)Q * 2 + 64) || c->s_c.ensure((size_t)Q * 4 + 64) || c->s_f.ensure(128)) return 1;
    align_profile_kernel<<<(Q + 3) / 4, 128, 0, st>>>(c->d_q.as<uint8_t>(), c->d_qoff64.as<uint64_t>(), Q, c->d_mat.as<Matrices>(), p->comp_bias,
                                           matMin, c->s_a.as<signed char>(), c->s_b.as<int>(), c->s_l.as<signed char>(), c->s_m.as<uint8_t>(),
                                           fastMaxL);
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    // a local score never exceeds rows x (largest profile entry): matrix maximum + largest composition bias (int8)
    const int maxScore = std::min(32766, SWF_ROWS * (matMax + 127));
    exh_min_score_kernel<<<(Q + 127) / 128, 128, 0, st>>>(c->d_qoff64.as<uint64_t>(), Q, p->gumbel, p->eval_thr, (double)p->db_res_count,
                                                          maxScore, c->s_n.as<uint16_t>());
    SPB_CUDA(cudaGetLastError()); c->n_launches++;

    // host plan: register-kernel queries longest first in tiles of EXH_TILE; the others take every target
    std::vector<uint32_t> qlist, slow;
    std::vector<uint8_t> mode((size_t)Q);       // 1: register kernel, 0: all targets (what exh_pair_count_kernel reads as qfast)
    uint64_t fastRes = 0;
    for (uint32_t k = 0; k < Q; k++) {
        const uint32_t q = c->h_qorder[k], L = c->h_qlen[q];
        if (L >= 1 && (int)L <= fastMaxL) { qlist.push_back(q); mode[q] = 1; fastRes += L; }
        else { slow.push_back(q); mode[q] = 0; }
    }
    const uint32_t nFast = (uint32_t)qlist.size(), nSlow = (uint32_t)slow.size();
    const uint32_t nTiles = (nFast + EXH_TILE - 1) / EXH_TILE;
    qlist.resize((size_t)nTiles * EXH_TILE, 0xFFFFFFFFu);
    // target chunks: enough (tile, chunk) units for ~16 waves of resident CTAs, at least ~2048 residues each
    std::vector<uint32_t> chunkStart;
    {
        const uint64_t wantUnits = (uint64_t)c->n_sm * 5 * 16;
        uint64_t nChunks = nTiles ? (wantUnits + nTiles - 1) / nTiles : 1;
        nChunks = std::max<uint64_t>(1, std::min<uint64_t>(nChunks, std::min<uint64_t>(T, c->a_res / 2048 + 1)));
        const uint64_t per = (c->a_res + nChunks - 1) / nChunks;
        chunkStart.push_back(0);
        uint64_t acc = 0;
        for (uint32_t t = 0; t < T; t++) {
            acc += c->h_aoff[t + 1] - c->h_aoff[t];
            if (acc >= per && t + 1 < T) { chunkStart.push_back(t + 1); acc = 0; }
        }
        chunkStart.push_back(T);
    }
    const uint32_t nChunks = (uint32_t)chunkStart.size() - 1;
    unsigned long long candCap = std::max<unsigned long long>(1ull << 22, 1024ull * Q);
    candCap = std::min<unsigned long long>(candCap, std::min<unsigned long long>((unsigned long long)Q * T, 0x7FFFFFF0ull));
    if (c->s_d.ensure(qlist.size() * 4 + 64) || c->s_e.ensure(chunkStart.size() * 4 + 64) || c->s_o.ensure(candCap * 8 + 64) ||
        c->s_p.ensure(candCap * 8 + 64) || c->s_g.ensure((size_t)nSlow * 4 + 64) || c->s_i.ensure(((size_t)Q + 1) * 32 + 64)) return 1;
    if (!qlist.empty()) SPB_CUDA(cudaMemcpyAsync(c->s_d.p, qlist.data(), qlist.size() * 4, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaMemcpyAsync(c->s_e.p, chunkStart.data(), chunkStart.size() * 4, cudaMemcpyHostToDevice, st));
    if (nSlow) SPB_CUDA(cudaMemcpyAsync(c->s_g.p, slow.data(), (size_t)nSlow * 4, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaMemcpyAsync(c->s_m.p, mode.data(), Q, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaMemsetAsync(c->s_f.p, 0, 128, st));
    SPB_CUDA(cudaMemsetAsync(c->s_c.p, 0, (size_t)Q * 4, st));

    ExhArgs A;
    memset(&A, 0, sizeof(A));
    A.qlist = c->s_d.as<uint32_t>();
    A.prof16 = c->s_l.as<uint4>();
    A.minScore = c->s_n.as<uint16_t>();
    A.tres = c->d_ares.as<uint8_t>();
    A.toff = c->d_aoff.as<uint64_t>();
    A.chunkStart = c->s_e.as<uint32_t>();
    A.nChunks = nChunks;
    A.cand = c->s_o.as<unsigned long long>();
    A.candCap = candCap;
    A.nCand = c->s_f.as<unsigned long long>();
    A.qCount = c->s_c.as<uint32_t>();
    A.go = p->gap_open;
    A.ge = p->gap_extend;
    SPB_CUDA(cudaEventRecord(c->ev0, st));
    for (uint32_t t0 = 0; t0 < nTiles;) {      // runs of tiles with the same row count
        auto rowsOf = [&](uint32_t tile) { return std::max<uint32_t>(8, c->h_qlen[qlist[(size_t)tile * EXH_TILE]]); };
        const uint32_t rows = rowsOf(t0);
        uint32_t t1 = t0 + 1;
        while (t1 < nTiles && rowsOf(t1) == rows) t1++;
        A.tile0 = t0;
        A.nTiles = t1 - t0;
        const uint64_t blocks = (uint64_t)A.nTiles * nChunks;
        SPB_CHECK(blocks < 0x7FFFFFFFull, "spb_pairs_exhaustive: too many work units; split the query batch");
        cudaError_t e = cudaSuccess;
        switch (rows) {
            case 8: e = launch_exhaustive<8>(A, (uint32_t)blocks, st); break;
            case 9: e = launch_exhaustive<9>(A, (uint32_t)blocks, st); break;
            case 10: e = launch_exhaustive<10>(A, (uint32_t)blocks, st); break;
            case 11: e = launch_exhaustive<11>(A, (uint32_t)blocks, st); break;
            case 12: e = launch_exhaustive<12>(A, (uint32_t)blocks, st); break;
            case 13: e = launch_exhaustive<13>(A, (uint32_t)blocks, st); break;
            case 14: e = launch_exhaustive<14>(A, (uint32_t)blocks, st); break;
            case 15: e = launch_exhaustive<15>(A, (uint32_t)blocks, st); break;
            default: e = launch_exhaustive<16>(A, (uint32_t)blocks, st); break;
        }
        SPB_CUDA(e); c->n_launches++;
        t0 = t1;
    }
    SPB_CUDA(cudaEventRecord(c->ev1, st));
    unsigned long long nCand = 0;
    SPB_CUDA(cudaMemcpyAsync(&nCand, c->s_f.p, 8, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    SPB_CHECK(nCand <= candCap, "spb_pairs_exhaustive: %llu pairs pass the E-value threshold, more than the %llu one batch can hold; "
              "lower -e or split the query batch", nCand, candCap);
    // candidates ordered by (query, target); CSR offsets over all queries
    size_t tmpBytes = 0;
    int endBit = 32;
    while (endBit < 64 && ((uint64_t)1 << (endBit - 32)) < (uint64_t)Q) endBit++;
    if (nCand) {
        SPB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, c->s_o.as<unsigned long long>(), c->s_p.as<unsigned long long>(),
                                                (int)nCand, 0, endBit, st));
        if (c->s_cub.ensure(tmpBytes)) return 1;
        SPB_CUDA(cub::DeviceRadixSort::SortKeys(c->s_cub.p, tmpBytes, c->s_o.as<unsigned long long>(), c->s_p.as<unsigned long long>(),
                                                (int)nCand, 0, endBit, st));
    }
    uint64_t *candCnt = c->s_i.as<uint64_t>(), *pairCnt = candCnt + (Q + 1), *candOff = pairCnt + (Q + 1);
    uint64_t *pairOff = c->d_pairoff.as<uint64_t>();
    exh_pair_count_kernel<<<(Q + 1 + 255) / 256, 256, 0, st>>>(c->s_c.as<uint32_t>(), c->s_m.as<uint8_t>(), Q, T, candCnt, pairCnt);
    SPB_CUDA(cudaGetLastError()); c->n_launches++;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, candCnt, candOff, (int)(Q + 1), st));
    if (c->s_cub.ensure(tmpBytes)) return 1;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, candCnt, candOff, (int)(Q + 1), st));
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, pairCnt, pairOff, (int)(Q + 1), st));
    uint64_t NP = 0;
    SPB_CUDA(cudaMemcpyAsync(&NP, pairOff + Q, 8, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    SPB_CHECK(NP < 0x7FFFFFF0ull, "spb_pairs_exhaustive: more than 2^31 pairs to align (%u queries longer than %d residues take every target); "
              "split the query batch", nSlow, fastMaxL);
    if (c->d_pairT.ensure(NP * 4 + 16)) return 1;
    if (nCand) {
        exh_scatter_kernel<<<(unsigned)((nCand + 255) / 256), 256, 0, st>>>(c->s_p.as<unsigned long long>(), nCand, candOff, pairOff,
                                                                            c->d_pairT.as<uint32_t>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    if (nSlow) {
        const uint64_t n = (uint64_t)nSlow * T;
        exh_fill_all_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->s_g.as<uint32_t>(), nSlow, T, pairOff, c->d_pairT.as<uint32_t>());
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    SPB_CUDA(cudaEventRecord(c->ev3, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    es.kernel_ms = ms;
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
    es.total_ms = ms;
    es.n_scored = (uint64_t)nFast * T;
    es.n_cells = fastRes * c->a_res;
    es.n_kept = nCand;
    es.n_unfiltered = (uint64_t)nSlow * T;
    es.n_pairs = NP;
    c->n_pairs = NP;
    c->pair_rank_mode = 2;
    c->pairs_max_per_query = NP;      // unknown per query: any
    c->have_pairs = true;
    if (stats) *stats = es;
    return 0;
}

extern "C" int spb_align_exhaustive(spb_ctx *c, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                                    const spb_align_params *p, spb_align_result **out, spb_exhaustive_stats *stats) {
    SPB_CHECK(c && q_offsets && p && out, "spb_align_exhaustive: NULL argument");
    *out = nullptr;
    if (spb_queries_set(c, q_residues, q_offsets, Q)) return 1;
    if (spb_pairs_exhaustive(c, p, stats)) return 1;
    if (spb_align_resident(c, p)) return 1;
    return spb_align_fetch(c, out);
}
