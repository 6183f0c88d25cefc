// setscore.cpp -- see setscore.h.  Compiled with -ffp-contract=off: the reference build (-mavx2, no FMA) evaluates every
// expression below as separate IEEE double operations, and the printed "%.3E" values must agree to the last digit.
#include "setscore.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <stdexcept>

#include "db_io.h"

namespace spbhost {

// ---------------------------------------------------------------------------------------------------------
// in-memory DBs
// ---------------------------------------------------------------------------------------------------------
size_t TextDB::find(uint32_t key) const {
    auto it = std::lower_bound(keys.begin(), keys.end(), key);
    if (it == keys.end() || *it != key) return SIZE_MAX;
    return (size_t)(it - keys.begin());
}

void TextDB::load(const std::string &name) {
    DBReader r;
    r.open(name);
    dbtype = r.dbtype();
    keys.resize(r.size());
    data.resize(r.size());
    for (size_t i = 0; i < r.size(); i++) {
        const DBEntry &e = r.byKey(i);
        keys[i] = e.key;
        const char *d = r.data(e);
        // an entry ends at its NUL (readers walk `while (*data != '\0')`), never beyond its indexed length
        data[i].assign(d, strnlen(d, e.length ? (size_t)e.length - 1 : 0));
    }
}

void TextDB::save(const std::string &name, int threads) const {
    DBWriter w;
    const int nt = threads > 0 ? threads : 1;
    w.open(name, dbtype, nt);
    const size_t n = keys.size();
    for (size_t i = 0; i < n; i++) w.write((int)(i * (size_t)nt / (n ? n : 1)), keys[i], data[i].data(), data[i].size());
    w.close();
}

// Util::fast_atoi<unsigned int> (M/commons/Util.h:100-114)
static uint32_t fast_atou(const char *s) {
    uint32_t v = 0;
    while (*s >= '0' && *s <= '9') v = v * 10u + (uint32_t)(*s++ - '0');
    return v;
}

void KeyMap::load(const std::string &name) {
    TextDB t;
    t.load(name);
    keys = t.keys;
    vals.resize(keys.size());
    for (size_t i = 0; i < keys.size(); i++) vals[i] = fast_atou(t.data[i].c_str());
}

bool KeyMap::get(uint32_t key, uint32_t *val) const {
    auto it = std::lower_bound(keys.begin(), keys.end(), key);
    if (it == keys.end() || *it != key) return false;
    *val = vals[(size_t)(it - keys.begin())];
    return true;
}

// SSTR(double) (M/commons/Util.cpp:713-718)
static void append_e3(std::string &out, double v) {
    char b[40];
    out.append(b, (size_t)snprintf(b, sizeof(b), "%.3E", v));
}

// ---------------------------------------------------------------------------------------------------------
// combineprotnuclaln (src/util/CombineProtNuclAln.cpp:55-95)
// ---------------------------------------------------------------------------------------------------------
// Util::getWordsOfLine (M/commons/Util.h:167-190): words separated by blanks / tabs up to the end of the line
static size_t words_of_line(const char *d, const char *end, const char **w, size_t maxw) {
    size_t n = 0;
    while (d < end && *d != '\n' && n < maxw) {
        while (d < end && (*d == ' ' || *d == '\t')) d++;
        if (d >= end || *d == '\n') break;
        w[n++] = d;
        while (d < end && *d != ' ' && *d != '\t' && *d != '\n') d++;
    }
    return n;
}

void combine_prot_nucl_entry(const std::string &protEntry, const std::string &nuclEntry, std::string &out) {
    out.clear();
    const char *pd = protEntry.data(), *pend = pd + protEntry.size();
    const char *nbeg = nuclEntry.data(), *nend = nbeg + nuclEntry.size();
    const char *pw[16], *nw[16];
    while (pd < pend) {
        const char *nl = (const char *)memchr(pd, '\n', (size_t)(pend - pd));
        const char *le = nl ? nl : pend;
        if (words_of_line(pd, le, pw, 16) < 10) throw std::runtime_error("Invalid alignment result record");
        const uint32_t target = fast_atou(pw[0]);
        const double protEval = strtod(pw[3], nullptr);
        double updatedEval = FLT_MAX, nuclSeqId = 0;
        for (const char *nd = nbeg; nd < nend;) {
            const char *nnl = (const char *)memchr(nd, '\n', (size_t)(nend - nd));
            const char *nle = nnl ? nnl : nend;
            if (fast_atou(nd) == target) {
                if (words_of_line(nd, nle, nw, 16) < 10) throw std::runtime_error("Invalid alignment result record");
                nuclSeqId = strtod(nw[2], nullptr);
                const double nuclEval = strtod(nw[3], nullptr);
                updatedEval = ((log(protEval) * 0.5 + log(nuclEval) * 0.5) < log(nuclEval)) ? exp(log(protEval) * 0.5 + log(nuclEval) * 0.5)
                                                                                             : nuclEval;
            }
            nd = nnl ? nnl + 1 : nend;
        }
        out.append(pw[0], (size_t)(pw[2] - pw[0]));
        append_e3(out, nuclSeqId);
        out.push_back('\t');
        append_e3(out, updatedEval);
        out.push_back('\t');
        out.append(pw[4], (size_t)(le - pw[4]));
        out.push_back('\n');
        pd = nl ? nl + 1 : pend;
    }
}

void combine_prot_nucl_aln(const TextDB &prot, const TextDB &nucl, int threads, TextDB &out) {
    out.dbtype = prot.dbtype;
    std::vector<std::string> tmp(prot.keys.size());
    std::vector<char> have(prot.keys.size(), 0);
    std::string err;
#pragma omp parallel for schedule(dynamic, 10) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)prot.keys.size(); i++) {
        const size_t j = nucl.find(prot.keys[(size_t)i]);
        if (j == SIZE_MAX) continue;
        try {
            combine_prot_nucl_entry(prot.data[(size_t)i], nucl.data[j], tmp[(size_t)i]);
        } catch (const std::exception &e) {
#pragma omp critical
            err = e.what();
        }
        have[(size_t)i] = 1;
    }
    if (!err.empty()) throw std::runtime_error(err);
    out.keys.clear();
    out.data.clear();
    for (size_t i = 0; i < prot.keys.size(); i++) {
        if (!have[i]) {
            fprintf(stderr, "Missing nucleotide alignment result for key %u \n", prot.keys[i]);
            continue;
        }
        out.keys.push_back(prot.keys[i]);
        out.data.push_back(std::move(tmp[i]));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Aggregation (M/multihit/Aggregation.cpp:24-90): the lines of one query entry grouped by the SET of their target,
// sets visited in ascending key order (std::map), lines of a set in entry order; columns split at tabs (Util::split =
// strtok: empty columns vanish).
// ---------------------------------------------------------------------------------------------------------
typedef std::vector<std::string> Columns;
typedef std::map<uint32_t, std::vector<Columns> > BySet;

static void split_tabs(const char *b, const char *e, Columns &out) {
    out.clear();
    while (b < e) {
        while (b < e && *b == '\t') b++;
        const char *s = b;
        while (b < e && *b != '\t') b++;
        if (b > s) out.emplace_back(s, (size_t)(b - s));
    }
}

static void build_map(const std::string &entry, const KeyMap &memberToSet, BySet &out) {
    out.clear();
    const char *d = entry.data(), *end = d + entry.size();
    Columns cols;
    while (d < end) {
        const char *nl = (const char *)memchr(d, '\n', (size_t)(end - d));
        const char *le = nl ? nl : end;
        if (le > d) {
            split_tabs(d, le, cols);
            if (!cols.empty()) {
                uint32_t setKey = 0;
                if (!memberToSet.get(fast_atou(cols[0].c_str()), &setKey))
                    throw std::runtime_error("Invalid target database key " + cols[0] + ".");
                out[setKey].push_back(cols);
            }
        }
        d = nl ? nl + 1 : end;
    }
}

template <class F>
static void aggregate(const TextDB &in, const KeyMap &memberToSet, int threads, TextDB &out, F entryOf) {
    out.keys = in.keys;
    out.data.assign(in.keys.size(), std::string());
    out.dbtype = DBTYPE_ALIGNMENT_RES;
    std::string err;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        BySet sets;
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)in.keys.size(); i++) {
            try {
                build_map(in.data[(size_t)i], memberToSet, sets);
                std::string &buf = out.data[(size_t)i];
                for (BySet::const_iterator it = sets.begin(); it != sets.end(); ++it) {
                    entryOf(it->second, in.keys[(size_t)i], it->first, buf);
                    buf.push_back('\n');
                }
            } catch (const std::exception &e) {
#pragma omp critical
                err = e.what();
            }
        }
    }
    if (!err.empty()) throw std::runtime_error(err);
}

// ---------------------------------------------------------------------------------------------------------
// besthitperset (M/multihit/besthitperset.cpp:29-115): per (query, target set) the best hit; its score column is replaced by
// the log of the best-hit P-value.  --simple-best-hit 1 (predictmatch, PredictMatch.cpp:26): the first line with the
// strictly smallest E-value; log P = log(E) below 1e-3, log(1 - exp(-E)) above, log(DBL_MIN) for E == 0.
// ---------------------------------------------------------------------------------------------------------
void best_hit_per_set(const TextDB &result, const KeyMap &targetMemberToSet, const KeyMap &targetSetSize, bool simpleBestHit,
                      int threads, TextDB &out) {
    aggregate(result, targetMemberToSet, threads, out,
              [&](const std::vector<Columns> &rows, uint32_t, uint32_t targetSetKey, std::string &buf) {
                  uint32_t nbrGenes = 0;
                  if (!targetSetSize.get(targetSetKey, &nbrGenes))
                      throw std::runtime_error("Invalid target size database key " + std::to_string(targetSetKey) + ".");
                  double bestScore = -DBL_MAX, secondBestScore = -DBL_MAX, bestEval = DBL_MAX;
                  const double logBestHitCalibration = log(1.0);
                  const bool simple = simpleBestHit || rows.size() < 2;
                  const Columns *best = nullptr;
                  for (size_t i = 0; i < rows.size(); i++) {
                      if (rows[i].size() < 4) throw std::runtime_error("Invalid result record.");
                      const double eval = strtod(rows[i][3].c_str(), nullptr);
                      double pval = eval / nbrGenes;
                      if (pval == 0) pval = DBL_MIN;
                      const double score = -log(pval);
                      if (simple) {
                          if (bestEval > eval) { bestEval = eval; best = &rows[i]; }
                      } else if (score >= bestScore) {
                          secondBestScore = bestScore;
                          bestScore = score;
                          best = &rows[i];
                      } else if (score > secondBestScore) {
                          secondBestScore = score;
                      }
                  }
                  double logCorrectedPval;
                  if (simple) {
                      if (bestEval == 0) logCorrectedPval = log(DBL_MIN) - logBestHitCalibration;
                      else if (bestEval > 0 && bestEval < 10e-4) logCorrectedPval = log(bestEval) - logBestHitCalibration;
                      else logCorrectedPval = log(1 - exp(-bestEval)) - logBestHitCalibration;
                  } else {
                      logCorrectedPval = secondBestScore - bestScore;
                  }
                  if (!best) return;
                  for (size_t c = 0; c < best->size(); c++) {
                      if (c == 1) append_e3(buf, logCorrectedPval); else buf += (*best)[c];
                      if (c + 1 != best->size()) buf.push_back('\t');
                  }
              });
}

// ---------------------------------------------------------------------------------------------------------
// mergeresultsbyset (M/util/mergeresultsbyset.cpp:10-55): per query set, the entries of its members back to back in the
// order of the <query>_set_to_member entry
// ---------------------------------------------------------------------------------------------------------
void merge_results_by_set(const TextDB &setToMember, const TextDB &result, int threads, TextDB &out) {
    out.keys = setToMember.keys;
    out.data.assign(out.keys.size(), std::string());
    out.dbtype = result.dbtype;
    std::string err;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (long i = 0; i < (long)setToMember.keys.size(); i++) {
        const std::string &members = setToMember.data[(size_t)i];
        const char *d = members.data(), *end = d + members.size();
        while (d < end) {
            const uint32_t key = fast_atou(d);
            const size_t id = result.find(key);
            if (id == SIZE_MAX) {
#pragma omp critical
                err = "Invalid key " + std::to_string(key) + " in entry " + std::to_string(i) + ".";
                break;
            }
            out.data[(size_t)i] += result.data[id];
            const char *nl = (const char *)memchr(d, '\n', (size_t)(end - d));
            d = nl ? nl + 1 : end;
        }
    }
    if (!err.empty()) throw std::runtime_error(err);
}

// ---------------------------------------------------------------------------------------------------------
// combinescore (src/util/CombineScore.cpp:41-98): per (query set, target set) the combined score of the best hits of the
// set's members (modified truncated-product method).  With t = log(1 / (#query ORFs + 1)):
//   hits with log P < t contribute t - log P; k of them: score = sum - t, seq. id. = their mean;
//   none: score = -log P and seq. id. of the LAST line whose log P is below the running value (the reference's update
//   rule, kept literally: the running value turns positive after the first line, so every later negative log P replaces it).
// ---------------------------------------------------------------------------------------------------------
void combine_score(const TextDB &merged, const KeyMap &targetMemberToSet, const KeyMap &querySetSize, int threads, TextDB &out) {
    aggregate(merged, targetMemberToSet, threads, out,
              [&](const std::vector<Columns> &rows, uint32_t querySetKey, uint32_t targetSetKey, std::string &buf) {
                  char kb[16];
                  buf.append(kb, (size_t)snprintf(kb, sizeof(kb), "%u", targetSetKey));
                  buf.push_back('\t');
                  uint32_t orfCount = 0;
                  if (!querySetSize.get(querySetKey, &orfCount))
                      throw std::runtime_error("Invalid query size database key " + std::to_string(querySetKey) + ".");
                  const double logPvalThreshold = log(1.0 / (orfCount + 1u));
                  double minLogPval = 0, sumLogPval = 0, minSeqId = 0, sumSeqId = 0;
                  size_t k = 0;
                  for (size_t i = 0; i < rows.size(); i++) {
                      if (rows[i].size() < 3) throw std::runtime_error("Invalid result record.");
                      const double logPvalue = strtod(rows[i][1].c_str(), nullptr);
                      const double seqId = strtod(rows[i][2].c_str(), nullptr);
                      if (logPvalue < minLogPval) {
                          minLogPval = (logPvalue == 0) ? logPvalue : -logPvalue;
                          minSeqId = seqId;
                      }
                      if (logPvalue < logPvalThreshold) {
                          sumLogPval -= logPvalue - logPvalThreshold;
                          sumSeqId += seqId;
                          k++;
                      }
                  }
                  if (k == 0) {
                      append_e3(buf, minLogPval);
                      buf.push_back('\t');
                      append_e3(buf, minSeqId);
                  } else {
                      const double avgSeqId = sumSeqId / k;
                      append_e3(buf, sumLogPval - logPvalThreshold);
                      buf.push_back('\t');
                      append_e3(buf, avgSeqId);
                  }
              });
}

// ---------------------------------------------------------------------------------------------------------
// filtermatchbyfdr (src/util/FilterMatchbyFdr.cpp:13-263): combined scores of the target run against those of the
// reversed-ORF control run -> score threshold at the requested false discovery rate; lines at or above it are kept.
// ---------------------------------------------------------------------------------------------------------
static bool is_ws(char c) { return c == ' ' || c == '\t'; }

// second column of every line (Util::getWordsOfLine + strtod)
static void collect_scores(const TextDB &db, std::vector<double> &out) {
    for (const std::string &entry : db.data) {
        const char *d = entry.c_str();
        while (*d != '\0') {
            const char *p = d;
            int words = 0;
            const char *w1 = nullptr;
            while (*p != '\n' && *p != '\0') {
                while (is_ws(*p)) p++;
                if (words == 1) w1 = p;
                words++;
                while (*p != '\n' && *p != '\0' && !is_ws(*p)) p++;
            }
            if (words < 2) throw std::runtime_error("Invalid result record.");
            out.push_back(strtod(w1, nullptr));
            while (*p != '\n' && *p != '\0') p++;
            d = (*p == '\n') ? p + 1 : p;
        }
    }
}

void filter_match_by_fdr(const TextDB &pos, const TextDB &neg, float fdrCutoffF, int reportFdr, int threads, TextDB &out, FdrReport *rep) {
    const double fdrCutoff = (double)fdrCutoffF;   // par.fdrCutoff is a float compared against doubles
    std::vector<double> P, N;
    collect_scores(pos, P);
    std::sort(P.begin(), P.end(), [](double a, double b) { return a > b; });
    if (P.empty()) P.push_back(0);
    double threshold = 0;
    FdrReport R;
    std::vector<double> fdrList, uniq;
    if (!neg.keys.empty()) {
        collect_scores(neg, N);
        std::sort(N.begin(), N.end(), [](double a, double b) { return a > b; });
        // one point per distinct score s of the target run: x = (#control scores > s + 0.5) / (#control + 1),
        // y = (#target scores > s) / #target; closed by (1, 1)
        std::vector<double> x, y;
        {
            double current = DBL_MAX;
            size_t negCounter = 0;
            for (size_t i = 0; i < P.size(); i++) {
                if (P[i] < current) {
                    current = P[i];
                    while (negCounter < N.size() && current < N[negCounter]) negCounter++;
                    uniq.push_back(current);
                    y.push_back(1.0 * i / P.size());
                    x.push_back(1.0 * (negCounter + 0.5) / (N.size() + 1));
                }
            }
            x.push_back(1.0);
            y.push_back(1.0);
        }
        // upper hull walked greedily: from point i the farthest point of maximal slope (ties go to the later point)
        std::vector<double> slopeList;
        std::vector<size_t> idxList;
        for (size_t i = 0; i < x.size() - 1;) {
            size_t jMax = i + 1;
            double slopeMax = 0;
            for (size_t j = i + 1; j < x.size(); j++) {
                const double slope = 1.0 * (y[j] - y[i]) / (x[j] - x[i]);
                if (slope >= slopeMax) { jMax = j; slopeMax = slope; }
            }
            i = jMax;
            slopeList.push_back(slopeMax);
            idxList.push_back(jMax);
        }
        // pi0 = mean of the last two slopes (with a single hull segment the reference reads before the array: undefined;
        // the one slope is used for both here)
        const double pi0 = ((slopeList.size() >= 2 ? slopeList[slopeList.size() - 2] : slopeList.back()) + slopeList.back()) / 2;
        size_t i = 0;
        double currentFDR = 0;
        if (x[idxList.back()] * pi0 / y[idxList.back()] >= fdrCutoff) {
            while (currentFDR <= fdrCutoff && i < idxList.size()) {
                currentFDR = x[idxList[i]] * pi0 / y[idxList[i]];
                i++;
            }
        }
        if (i < 2) {
            threshold = P.back();
            R.mode = 1;
        } else {
            size_t j = idxList[i - 2];
            double TPFP = y[j];
            double FP = x[j] * pi0;
            currentFDR = 0;
            while (currentFDR <= fdrCutoff && j + 1 < x.size()) {
                j++;
                const double deltaX = x[j] - x[j - 1];
                TPFP += deltaX * slopeList[i - 1];
                FP += deltaX * pi0;
                currentFDR = FP / TPFP;
            }
            if (j >= uniq.size()) j = uniq.size() - 1;   // the closing point has no score (the reference would read past the list)
            threshold = uniq[j];
            R.passed = y[j] * P.size();
        }
        if (reportFdr) {
            for (size_t j = 0; j < idxList[0]; j++) fdrList.push_back(0.0);
            for (size_t s = 0; s + 1 < idxList.size(); s++) {
                double TPFP = y[idxList[s]];
                double FP = x[idxList[s]] * pi0;
                for (size_t j = idxList[s]; j < idxList[s + 1]; j++) {
                    if (std::isinf(slopeList[s])) {
                        fdrList.push_back(x[idxList[s]] * pi0 / y[idxList[s]]);
                    } else {
                        TPFP += (x[j] - x[j - 1]) * slopeList[s];
                        FP += (x[j] - x[j - 1]) * pi0;
                        fdrList.push_back(FP / TPFP);
                    }
                }
            }
        }
    } else {
        R.mode = 2;
        if (reportFdr) {
            for (size_t i = 0; i < P.size(); i++)
                if (DBL_MIN < P[i]) { uniq.push_back(P[i]); fdrList.push_back(0.0); }
        }
        threshold = P.back();
    }
    R.threshold = threshold;
    if (rep) *rep = R;

    out.keys = pos.keys;
    out.data.assign(pos.keys.size(), std::string());
    out.dbtype = DBTYPE_GENERIC_DB;
    std::string err;
#pragma omp parallel for schedule(dynamic, 10) num_threads(threads > 0 ? threads : 1)
    for (long e = 0; e < (long)pos.keys.size(); e++) {
        const char *d = pos.data[(size_t)e].c_str();
        std::string &buf = out.data[(size_t)e];
        while (*d != '\0') {
            // words[0], words[1], words[3] of Util::getWordsOfLine: words[3] is the 4th column or, on a 3-column line, its end
            const char *p = d, *w0 = nullptr, *w1 = nullptr, *w3 = nullptr;
            int words = 0;
            while (*p != '\n' && *p != '\0') {
                while (is_ws(*p)) p++;
                if (words == 0) w0 = p;
                if (words == 1) w1 = p;
                if (words == 3) w3 = p;
                words++;
                while (*p != '\n' && *p != '\0' && !is_ws(*p)) p++;
            }
            if (words <= 3) w3 = p;
            if (words < 2) {
#pragma omp critical
                err = "Invalid result record.";
                break;
            }
            const double score = strtod(w1, nullptr);
            if (score >= threshold) {
                buf.append(w0, (size_t)(w3 - w0));
                if (reportFdr) {
                    buf.push_back('\t');
                    const size_t idx = (size_t)(std::find(uniq.begin(), uniq.end(), score) - uniq.begin());
                    append_e3(buf, idx < fdrList.size() ? fdrList[idx] : 0.0);
                }
                buf.push_back('\n');
            }
            while (*p != '\n' && *p != '\0') p++;
            d = (*p == '\n') ? p + 1 : p;
        }
    }
    if (!err.empty()) throw std::runtime_error(err);
}

}  // namespace spbhost
