// submat.cpp -- see submat.h
#include "submat.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace spbhost {

namespace {
struct Embedded {
    const char *name, *lambda, *background, *alphabet;
    const char *rows[21];
};
const Embedded kEmbedded[] = {
#include "matrix_data.inc"
};

std::vector<std::string> words(const std::string &line) {
    std::vector<std::string> w;
    std::istringstream is(line);
    std::string t;
    while (is >> t) w.push_back(t);
    return w;
}
}  // namespace

std::string multiparam_aa(const std::string &spec) {
    // "nucl:X,aa:Y" in any order; a value without prefix applies to both
    size_t start = 0;
    std::string fallback;
    while (start <= spec.size()) {
        size_t comma = spec.find(',', start);
        std::string part = spec.substr(start, comma == std::string::npos ? std::string::npos : comma - start);
        if (part.compare(0, 3, "aa:") == 0) return part.substr(3);
        if (part.compare(0, 5, "nucl:") != 0 && fallback.empty()) fallback = part;
        if (comma == std::string::npos) break;
        start = comma + 1;
    }
    return fallback;
}

std::string multiparam_nucl(const std::string &spec) {
    size_t start = 0;
    std::string fallback;
    while (start <= spec.size()) {
        size_t comma = spec.find(',', start);
        std::string part = spec.substr(start, comma == std::string::npos ? std::string::npos : comma - start);
        if (part.compare(0, 5, "nucl:") == 0) return part.substr(5);
        if (part.compare(0, 3, "aa:") != 0 && fallback.empty()) fallback = part;
        if (comma == std::string::npos) break;
        start = comma + 1;
    }
    return fallback;
}

void SubMat::load(const std::string &spec, double bitFactor, double scoreBias) {
    const std::string path = multiparam_aa(spec);
    if (path.empty()) throw std::runtime_error("no amino-acid substitution matrix in '" + spec + "'");
    std::string alph;
    std::vector<std::vector<double>> S;
    std::vector<double> bg;
    double lam = 0;
    bool haveLam = false;
    size_t slash = path.find_last_of('/');
    name = slash == std::string::npos ? path : path.substr(slash + 1);

    std::ifstream in(path.c_str());
    if (in.good()) {
        // .out layout: optional "# Background (precomputed optional):" / "# Lambda     (precomputed optional):" comment lines,
        // a header line with the alphabet, then one row per letter (SubstitutionMatrix.cpp:327-407)
        std::string line;
        std::vector<char> rowLetters;
        while (std::getline(in, line)) {
            if (line.compare(0, 36, "# Background (precomputed optional):") == 0) {
                std::vector<std::string> w = words(line);
                for (size_t i = 4; i < w.size(); i++) bg.push_back(strtod(w[i].c_str(), nullptr));
            } else if (line.compare(0, 36, "# Lambda     (precomputed optional):") == 0) {
                std::vector<std::string> w = words(line);
                if (w.size() > 4) { lam = strtod(w[4].c_str(), nullptr); haveLam = true; }
            }
            if (line.empty() || line[0] == '#') continue;
            std::vector<std::string> w = words(line);
            if (w.size() <= 1) continue;
            if (alph.empty()) {
                for (const std::string &x : w) alph.push_back((char)toupper(x[0]));
                continue;
            }
            std::vector<double> row;
            for (size_t i = 1; i < w.size() && row.size() < alph.size(); i++) row.push_back(strtod(w[i].c_str(), nullptr));
            rowLetters.push_back((char)toupper(w[0][0]));
            S.push_back(row);
        }
        // rows may come in any order: bring them into alphabet order
        std::vector<std::vector<double>> ordered(alph.size());
        for (size_t r = 0; r < rowLetters.size(); r++) {
            size_t at = alph.find(rowLetters[r]);
            if (at != std::string::npos) ordered[at] = S[r];
        }
        S.swap(ordered);
    } else {
        const Embedded *e = nullptr;
        for (const Embedded &m : kEmbedded)
            if (name == m.name) e = &m;
        if (!e) throw std::runtime_error("Could not open substitution matrix '" + path + "' (embedded: VTML40.out VTML80.out blosum62.out)");
        alph = e->alphabet;
        lam = strtod(e->lambda, nullptr);
        haveLam = true;
        for (const std::string &x : words(e->background)) bg.push_back(strtod(x.c_str(), nullptr));
        for (size_t r = 0; r < alph.size(); r++) {
            std::vector<double> row;
            for (const std::string &x : words(e->rows[r])) row.push_back(strtod(x.c_str(), nullptr));
            S.push_back(row);
        }
    }
    const size_t n = alph.size();
    if (n != (size_t)N) throw std::runtime_error("substitution matrix '" + path + "' must have the 21-letter alphabet ACDEFGHIKLMNPQRSTVWYX");
    const size_t x = alph.find('X');
    if (x == std::string::npos) throw std::runtime_error("Please add X to your substitution matrix.");
    if (!haveLam || bg.size() < n) throw std::runtime_error("substitution matrix '" + path + "' lacks the precomputed lambda/background header");
    for (size_t i = 0; i < n; i++)
        if (S[i].size() < n) throw std::runtime_error("substitution matrix '" + path + "' is incomplete");
    bool xPositive = false;
    for (size_t j = 0; j < n; j++) xPositive = xPositive || S[x][j] > 0 || S[j][x] > 0;
    double pb[N];
    for (size_t i = 0; i < n; i++) pb[i] = bg[i];
    if (!xPositive)
        for (size_t i = 0; i + 1 < n; i++) pb[i] = pb[i] * (1.0 - pb[x]);
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) prob[i * N + j] = exp(lam * S[i][j]) * pb[i] * pb[j];
    double rows[N];
    for (size_t i = 0; i < n; i++) {
        double acc = 0.0;
        for (size_t j = 0; j < n; j++) acc += prob[i * N + j];
        rows[i] = acc;
    }
    rows[n - 1] = 1e-5;   // ANY_BACK
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            const double v = bitFactor * log2(prob[i * N + j] / (rows[i] * rows[j])) + scoreBias;
            sub[i * N + j] = (int16_t)(v < 0.0 ? (int)(v - 0.5) : (int)(v + 0.5));
        }
    for (size_t i = 0; i < n; i++) pback[i] = pb[i];
    lambda = lam;
    // letter mapping
    for (int c = 0; c < 256; c++) aa2num[c] = (uint8_t)x;
    for (int c = 0; c < 256; c++) {
        const char up = (char)toupper(c);
        size_t at = std::string::npos;
        if (strchr("ATGCDEFHIKLMNPQRSVWYX", up) && up != '\0') at = alph.find(up);
        else if (up == 'J') at = alph.find('L');
        else if (up == 'Z') at = alph.find('E');
        else if (up == 'B') at = alph.find('D');
        if (at != std::string::npos) aa2num[c] = (uint8_t)at;
    }
    for (size_t i = 0; i < n; i++) num2aa[i] = alph[i];
    num2aa[n] = '\0';
}

}  // namespace spbhost
