// submat.h -- substitution matrices of the host binary (embedded tables or a .out file).
//
// Mirrors SubstitutionMatrix::readProbMatrix + BaseMatrix::generateSubMatrix
// (M/commons/SubstitutionMatrix.cpp:327-407, M/commons/BaseMatrix.cpp:96-156) and the letter mapping
// (SubstitutionMatrix.cpp:257-297).  The Python side (spacepharer_b200/submat.py) computes the same doubles.
#pragma once

#include <stdint.h>

#include <string>

namespace spbhost {

struct SubMat {
    enum { N = 21 };
    int16_t sub[N * N];        // round_half_away(bitFactor * log2(P_ab / (pa pb)) + bias)
    double pback[N];           // BaseMatrix::pBack
    double prob[N * N];        // joint probabilities P_ab (tantan likelihood ratios)
    double lambda;
    uint8_t aa2num[256];
    char num2aa[N + 1];
    std::string name;          // file name without directories, e.g. "VTML40.out"

    // `spec` is what --sub-mat carries: "aa:VTML40.out", "nucl:nucleotide.out,aa:VTML40.out", a bare name or a path
    // (MultiParam<char*>, M/commons/MultiParam.cpp); the amino-acid entry is used.  Throws std::runtime_error.
    void load(const std::string &spec, double bitFactor, double scoreBias);
    void encode(const char *seq, size_t len, uint8_t *out) const {
        for (size_t i = 0; i < len; i++) out[i] = aa2num[(unsigned char)seq[i]];
    }
};

// "nucl:10,aa:16" / "16" -> amino-acid value
std::string multiparam_aa(const std::string &spec);
// ... -> nucleotide value
std::string multiparam_nucl(const std::string &spec);

}  // namespace spbhost
