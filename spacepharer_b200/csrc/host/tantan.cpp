// tantan.cpp -- host pre-step of the index build: tantan repeat masking of the target residues (--mask 1).
//
// Mirrors the call in M/prefiltering/IndexBuilder.cpp:137-149 (maxCycle 50, repeatProb 0.005, repeatEndProb 0.05,
// decay 0.9, no gaps, minMaskProb 0.9, every letter masked to X) on the gapless code path of
// M/commons/tantan.cpp:221-235,318-351,365-396; likelihood ratios as ProbabilityMatrix (M/commons/BaseMatrix.h:80-109).
// All arithmetic in double except the per-letter forward value kept as float, as in the reference.  Compiled with
// -ffp-contract=off so that no a*b+c is fused (the reference build has no FMA).  Kept on the host like the
// reference does: O(50 * residues) once per target DB, OpenMP over targets; on AVX2 hosts four targets share a vector (below).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../../include/spb200.h"

namespace {

const int kMaxOff = 50, kStep = 16, kAlph = SPB_ALPHABET, kX = 20;

struct TantanTables {
    double lr[kAlph][kAlph];
    double b2f[kMaxOff];
    double b2b, f2b, f2f0;
};

int mask_one(const TantanTables &tt, uint8_t *seq, int L, std::vector<float> &prob, std::vector<double> &scale) {
    if (L <= 0) return 0;
    prob.resize((size_t)L);
    scale.assign((size_t)L / kStep + 1, 0.0);
    double fg[kMaxOff];
    double B = 1.0;
    for (int i = 0; i < kMaxOff; i++) fg[i] = 0.0;
    for (int p = 0; p < L; p++) {
        const double *row = tt.lr[seq[p] < kAlph ? seq[p] : kX];
        const int n = p < kMaxOff ? p : kMaxOff;
        const double b = B;
        double from = 0;
        for (int i = 0; i < n; i++) {
            const double f = fg[i];
            from += f;
            const uint8_t o = seq[p - i - 1];
            fg[i] = (b * tt.b2f[i] + f * tt.f2f0) * row[o < kAlph ? o : kX];
        }
        B = b * tt.b2b + from * tt.f2b;
        if (p % kStep == kStep - 1) {
            const double s = 1 / B;
            scale[p / kStep] = s;
            B *= s;
            for (int i = 0; i < kMaxOff; i++) fg[i] *= s;
        }
        prob[p] = (float)B;
    }
    double sum = 0.0;
    for (int i = 0; i < kMaxOff; i++) sum += fg[i];
    const double z = B * tt.b2b + sum * tt.f2b;
    B = tt.b2b;
    for (int i = 0; i < kMaxOff; i++) fg[i] = tt.f2b;
    for (int p = L - 1; p >= 0; p--) {
        const double nonRep = prob[p] * B / z;
        prob[p] = 1 - (float)nonRep;
        if (p % kStep == kStep - 1) {
            const double s = scale[p / kStep];
            B *= s;
            for (int i = 0; i < kMaxOff; i++) fg[i] *= s;
        }
        const double *row = tt.lr[seq[p] < kAlph ? seq[p] : kX];
        const int n = p < kMaxOff ? p : kMaxOff;
        const double toB = tt.f2b * B;
        double toF = 0;
        for (int i = 0; i < n; i++) {
            const uint8_t o = seq[p - i - 1];
            const double f = fg[i] * row[o < kAlph ? o : kX];
            toF += tt.b2f[i] * f;
            fg[i] = toB + tt.f2f0 * f;
        }
        B = tt.b2b * B + toF;
    }
    int masked = 0;
    const double minMaskProb = 0.9;
    for (int p = 0; p < L; p++)
        if (prob[p] >= minMaskProb) {
            seq[p] = kX;
            masked++;
        }
    return masked;
}

// ---------------------------------------------------------------------------------------------------------
// Four sequences at a time (AVX2).  The forward / backward recursions of one sequence are a chain of dependent double additions in a
// fixed order (`from += f`, `toF += b2f[i] * f`: 50 per residue and direction), so one sequence alone runs at the latency of the FP
// adder, and re-ordering the sums would change the rounding.  Four DIFFERENT sequences in the four lanes of a vector keep every
// lane's operation sequence exactly the scalar one (same order, no FMA: the function is compiled for avx2 without fma) and fill the
// pipeline: bit-identical masks (53 M residues of the config-2 targets and a repeat-rich adversarial set compared), 0.228 -> 0.147 s
// for 47.6 M residues on the 16 host threads of the B200 box (tools/tantan_probe.cpp; a version with vgatherdpd reached only 0.20 s:
// the gather is microcoded there).  Lanes advance in lockstep from position 0; the residues of a group are interleaved
// (`R[p][lane]`), a lane that is shorter than the group is re-initialised when the backward pass reaches its last residue, its
// partition function is taken when the forward pass leaves it.  SPB_TANTAN_SCALAR=1 selects the one-sequence path.
// ---------------------------------------------------------------------------------------------------------
#if defined(__x86_64__)
#include <immintrin.h>
#define SPB_TANTAN_AVX2 1

struct Group4 {
    std::vector<uint8_t> R;          // [Lmax][4] residues, padded with X
    std::vector<float> prob;         // [Lmax][4]
    std::vector<double> scale;       // [Lmax / 16 + 1][4]
};

__attribute__((target("avx2")))
void mask_four(const TantanTables &tt, uint8_t *const seq[4], const int len[4], Group4 &g, int masked[4]) {
    int Lmax = 0;
    for (int l = 0; l < 4; l++) Lmax = len[l] > Lmax ? len[l] : Lmax;
    for (int l = 0; l < 4; l++) masked[l] = 0;
    if (Lmax <= 0) return;
    g.R.assign((size_t)Lmax * 4, (uint8_t)kX);
    g.prob.resize((size_t)Lmax * 4);
    g.scale.assign(((size_t)Lmax / kStep + 1) * 4, 0.0);
    for (int l = 0; l < 4; l++)
        for (int p = 0; p < len[l]; p++) g.R[(size_t)p * 4 + l] = seq[l][p] < kAlph ? seq[l][p] : (uint8_t)kX;
    const double *lr = &tt.lr[0][0];
    const __m256d b2b = _mm256_set1_pd(tt.b2b), f2b = _mm256_set1_pd(tt.f2b), f2f0 = _mm256_set1_pd(tt.f2f0), one = _mm256_set1_pd(1.0);
    const __m128i len4 = _mm_loadu_si128((const __m128i *)len);
    alignas(32) double fg[kMaxOff][4];
    __m256d B = one, z = one;
    for (int i = 0; i < kMaxOff; i++) _mm256_store_pd(fg[i], _mm256_setzero_pd());
    const uint8_t *const Rp = g.R.data();
    // lr[a_l][o_l] of the four lanes without a gather instruction (microcoded and slow on current Xeons): four scalar loads from
    // the rows of the lanes' current residues
#define LR4(rows, q) _mm256_set_pd((rows)[3][Rp[(size_t)(q) * 4 + 3]], (rows)[2][Rp[(size_t)(q) * 4 + 2]], \
                                   (rows)[1][Rp[(size_t)(q) * 4 + 1]], (rows)[0][Rp[(size_t)(q) * 4 + 0]])
    for (int p = 0; p < Lmax; p++) {
        const double *const rows[4] = {lr + kAlph * Rp[(size_t)p * 4 + 0], lr + kAlph * Rp[(size_t)p * 4 + 1],
                                       lr + kAlph * Rp[(size_t)p * 4 + 2], lr + kAlph * Rp[(size_t)p * 4 + 3]};
        const int n = p < kMaxOff ? p : kMaxOff;
        const __m256d b = B;
        __m256d from = _mm256_setzero_pd();
        for (int i = 0; i < n; i++) {
            const __m256d f = _mm256_load_pd(fg[i]);
            from = _mm256_add_pd(from, f);
            const __m256d w = LR4(rows, p - i - 1);
            const __m256d t = _mm256_add_pd(_mm256_mul_pd(b, _mm256_set1_pd(tt.b2f[i])), _mm256_mul_pd(f, f2f0));
            _mm256_store_pd(fg[i], _mm256_mul_pd(t, w));
        }
        B = _mm256_add_pd(_mm256_mul_pd(b, b2b), _mm256_mul_pd(from, f2b));
        if (p % kStep == kStep - 1) {
            const __m256d sc = _mm256_div_pd(one, B);
            _mm256_storeu_pd(&g.scale[(size_t)(p / kStep) * 4], sc);
            B = _mm256_mul_pd(B, sc);
            for (int i = 0; i < kMaxOff; i++) _mm256_store_pd(fg[i], _mm256_mul_pd(_mm256_load_pd(fg[i]), sc));
        }
        _mm_storeu_ps(&g.prob[(size_t)p * 4], _mm256_cvtpd_ps(B));
        // a lane whose sequence ends here: its partition function z = B * b2b + sum(fg) * f2b
        const __m128i ends = _mm_cmpeq_epi32(len4, _mm_set1_epi32(p + 1));
        if (_mm_movemask_epi8(ends)) {
            __m256d sum = _mm256_setzero_pd();
            for (int i = 0; i < kMaxOff; i++) sum = _mm256_add_pd(sum, _mm256_load_pd(fg[i]));
            const __m256d zz = _mm256_add_pd(_mm256_mul_pd(B, b2b), _mm256_mul_pd(sum, f2b));
            z = _mm256_blendv_pd(z, zz, _mm256_castsi256_pd(_mm256_cvtepi32_epi64(ends)));
        }
    }
    B = b2b;
    for (int i = 0; i < kMaxOff; i++) _mm256_store_pd(fg[i], f2b);
    for (int p = Lmax - 1; p >= 0; p--) {
        // a lane whose last residue is p starts its backward recursion here
        const __m128i starts = _mm_cmpeq_epi32(len4, _mm_set1_epi32(p + 1));
        if (p != Lmax - 1 && _mm_movemask_epi8(starts)) {
            const __m256d m = _mm256_castsi256_pd(_mm256_cvtepi32_epi64(starts));
            B = _mm256_blendv_pd(B, b2b, m);
            for (int i = 0; i < kMaxOff; i++) _mm256_store_pd(fg[i], _mm256_blendv_pd(_mm256_load_pd(fg[i]), f2b, m));
        }
        const __m256d pr = _mm256_cvtps_pd(_mm_loadu_ps(&g.prob[(size_t)p * 4]));
        const __m256d nonRep = _mm256_div_pd(_mm256_mul_pd(pr, B), z);
        _mm_storeu_ps(&g.prob[(size_t)p * 4], _mm_sub_ps(_mm_set1_ps(1.0f), _mm256_cvtpd_ps(nonRep)));
        if (p % kStep == kStep - 1) {
            const __m256d sc = _mm256_loadu_pd(&g.scale[(size_t)(p / kStep) * 4]);
            B = _mm256_mul_pd(B, sc);
            for (int i = 0; i < kMaxOff; i++) _mm256_store_pd(fg[i], _mm256_mul_pd(_mm256_load_pd(fg[i]), sc));
        }
        const double *const rows[4] = {lr + kAlph * Rp[(size_t)p * 4 + 0], lr + kAlph * Rp[(size_t)p * 4 + 1],
                                       lr + kAlph * Rp[(size_t)p * 4 + 2], lr + kAlph * Rp[(size_t)p * 4 + 3]};
        const int n = p < kMaxOff ? p : kMaxOff;
        const __m256d toB = _mm256_mul_pd(f2b, B);
        __m256d toF = _mm256_setzero_pd();
        for (int i = 0; i < n; i++) {
            const __m256d w = LR4(rows, p - i - 1);
            const __m256d f = _mm256_mul_pd(_mm256_load_pd(fg[i]), w);
            toF = _mm256_add_pd(toF, _mm256_mul_pd(_mm256_set1_pd(tt.b2f[i]), f));
            _mm256_store_pd(fg[i], _mm256_add_pd(toB, _mm256_mul_pd(f2f0, f)));
        }
        B = _mm256_add_pd(_mm256_mul_pd(b2b, B), toF);
    }
    const double minMaskProb = 0.9;
    for (int l = 0; l < 4; l++)
        for (int p = 0; p < len[l]; p++)
            if (g.prob[(size_t)p * 4 + l] >= minMaskProb) {
                seq[l][p] = kX;
                masked[l]++;
            }
#undef LR4
}
#endif

}  // namespace

extern "C" int spb_tantan_mask(const double *prob, const double *pback, uint8_t *residues, const uint64_t *offsets, uint32_t T,
                               int threads, uint64_t *n_masked) {
    if (!prob || !pback || !residues || !offsets) return 1;
    TantanTables tt;
    for (int i = 0; i < kAlph; i++)
        for (int j = 0; j < kAlph; j++) tt.lr[i][j] = prob[i * kAlph + j] / (pback[i] * pback[j]);
    const double repeatProb = 0.005, repeatEndProb = 0.05, decay = 0.9;
    tt.b2b = 1 - repeatProb;
    tt.f2b = repeatEndProb;
    tt.f2f0 = 1 - repeatEndProb;
    double p = repeatProb * ((1 - decay) / (1 - pow(decay, kMaxOff)));
    for (int i = 0; i < kMaxOff; i++) {
        tt.b2f[i] = p;
        p *= decay;
    }
    uint64_t total = 0;
#ifdef SPB_TANTAN_AVX2
    if (T >= 8 && __builtin_cpu_supports("avx2") && !getenv("SPB_TANTAN_SCALAR")) {
        // groups of four sequences of similar length (longest first: the tail of the parallel loop is made of short groups)
        std::vector<uint32_t> order(T);
        for (uint32_t t = 0; t < T; t++) order[t] = t;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b];
        });
        bool tooLong = false;
        for (uint32_t t = 0; t < T; t++) tooLong = tooLong || (offsets[t + 1] - offsets[t] > 0x7FFFFFFFull);
        if (!tooLong) {
            const long nGroups = (long)(T / 4);
#pragma omp parallel num_threads(threads > 0 ? threads : 1) reduction(+ : total)
            {
                Group4 g;
                std::vector<float> pr;
                std::vector<double> sc;
#pragma omp for schedule(dynamic, 16)
                for (long gi = 0; gi < nGroups; gi++) {
                    uint8_t *seq[4];
                    int len[4], masked[4];
                    for (int l = 0; l < 4; l++) {
                        const uint32_t t = order[(size_t)gi * 4 + l];
                        seq[l] = residues + offsets[t];
                        len[l] = (int)(offsets[t + 1] - offsets[t]);
                    }
                    mask_four(tt, seq, len, g, masked);
                    total += (uint64_t)(masked[0] + masked[1] + masked[2] + masked[3]);
                }
#pragma omp for schedule(dynamic, 1)
                for (long r = nGroups * 4; r < (long)T; r++) {
                    const uint32_t t = order[(size_t)r];
                    total += (uint64_t)mask_one(tt, residues + offsets[t], (int)(offsets[t + 1] - offsets[t]), pr, sc);
                }
            }
            if (n_masked) *n_masked = total;
            return 0;
        }
    }
#endif
#pragma omp parallel num_threads(threads > 0 ? threads : 1) reduction(+ : total)
    {
        std::vector<float> pr;
        std::vector<double> sc;
#pragma omp for schedule(dynamic, 100)
        for (long t = 0; t < (long)T; t++)
            total += (uint64_t)mask_one(tt, residues + offsets[t], (int)(offsets[t + 1] - offsets[t]), pr, sc);
    }
    if (n_masked) *n_masked = total;
    return 0;
}
