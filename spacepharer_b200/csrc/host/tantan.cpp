// tantan.cpp -- host pre-step of the index build: tantan repeat masking of the target residues (--mask 1).
//
// Mirrors the call in M/prefiltering/IndexBuilder.cpp:137-149 (maxCycle 50, repeatProb 0.005, repeatEndProb 0.05,
// decay 0.9, no gaps, minMaskProb 0.9, every letter masked to X) on the gapless code path of
// M/commons/tantan.cpp:221-235,318-351,365-396; likelihood ratios as ProbabilityMatrix (M/commons/BaseMatrix.h:80-109).
// All arithmetic in double except the per-letter forward value kept as float, as in the reference.  Compiled with
// -ffp-contract=off so that no a*b+c is fused (the reference build has no FMA).  Kept on the host like the
// reference does: O(50 * residues) once per target DB, OpenMP over targets.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

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
