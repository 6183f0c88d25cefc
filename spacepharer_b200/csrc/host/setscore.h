// setscore.h -- the set-level stages that consume the search result DBs in predictmatch:
//   besthitperset     lib/mmseqs/src/multihit/besthitperset.cpp:29-115 (+ Aggregation.cpp:24-90)
//   mergeresultsbyset lib/mmseqs/src/util/mergeresultsbyset.cpp:10-55
//   combinescore      src/util/CombineScore.cpp:41-98
//   filtermatchbyfdr  src/util/FilterMatchbyFdr.cpp:13-263
//   combineprotnuclaln src/util/CombineProtNuclAln.cpp:55-95 (text half of the nucleotide re-alignment branch, 8f row 3)
// Host C++ like the reference (text DBs in, text DBs out; cheap arithmetic, double precision, printed with "%.3E").
// Every stage works on in-memory text DBs, so the fused `spb200 predictmatch` runs search -> ... -> match without
// re-reading its intermediates while the module commands stay byte-compatible with the reference's DBs.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

namespace spbhost {

enum { DBTYPE_GENERIC_DB = 12 };

// a result DB held in memory: entries sorted by key, payload without the terminating NUL
struct TextDB {
    std::vector<uint32_t> keys;
    std::vector<std::string> data;
    int dbtype = 5;
    size_t find(uint32_t key) const;            // position or SIZE_MAX
    void load(const std::string &name);         // MMseqs2 DB on disk -> memory
    void save(const std::string &name, int threads) const;
};

// key -> unsigned value DBs (<db>_member_to_set, <db>_set_size): first number of every entry (Util::fast_atoi)
struct KeyMap {
    std::vector<uint32_t> keys, vals;
    void load(const std::string &name);
    bool get(uint32_t key, uint32_t *val) const;
};

struct FdrReport {
    double threshold = 0;
    double passed = 0;          // y[j] * #scores, as the reference prints it
    int mode = 0;               // 0 threshold from the FDR curve, 1 list too short / --fdr 1, 2 empty control set
};

void best_hit_per_set(const TextDB &result, const KeyMap &targetMemberToSet, const KeyMap &targetSetSize, bool simpleBestHit,
                      int threads, TextDB &out);
void merge_results_by_set(const TextDB &setToMember, const TextDB &result, int threads, TextDB &out);
void combine_score(const TextDB &merged, const KeyMap &targetMemberToSet, const KeyMap &querySetSize, int threads, TextDB &out);
// combineprotnuclaln (src/util/CombineProtNuclAln.cpp:55-95), nucleotide re-alignment branch of predictmatch: per protein row
// the LAST row of the nucleotide entry with the same target key replaces the sequence identity (printed "%.3E") and gives
// the E-value min(sqrt(protE * nuclE), nuclE) computed from the PRINTED values of both DBs; no such row: 0 and FLT_MAX.
void combine_prot_nucl_entry(const std::string &protEntry, const std::string &nuclEntry, std::string &out);
// entries of `prot` without an entry of the same key in `nucl` are dropped with the reference's warning on stderr
void combine_prot_nucl_aln(const TextDB &prot, const TextDB &nucl, int threads, TextDB &out);
void filter_match_by_fdr(const TextDB &pos, const TextDB &neg, float fdrCutoff, int reportFdr, int threads, TextDB &out, FdrReport *rep);

}  // namespace spbhost
