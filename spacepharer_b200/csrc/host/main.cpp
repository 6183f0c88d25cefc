// main.cpp -- spb200: drop-in `prefilter` and `align` modules over libspb200 (the C ABI in include/spb200.h).
//
// Accepts the argv the reference's workflow generates for these two steps (lib/mmseqs/data/workflow/blastp.sh:33,41;
// flag sets M/commons/Parameters.cpp:337-398; the effective SpacePHARER command lines are in SURVEY.md 3.2), reads and
// writes the MMseqs2 DB format byte-compatibly, and prints the rows exactly like QueryMatcher::prefilterHitToBuffer
// (M/prefiltering/QueryMatcher.h:114-126) and Matcher::resultToBuffer (M/alignment/Matcher.cpp:323-370).
//   spb200 prefilter <queryDB> <targetDB> <prefilterDB> [flags]     Prefiltering (M/prefiltering/Prefiltering.cpp)
//   spb200 align     <queryDB> <targetDB> <prefilterDB> <alnDB> [flags]   Alignment (M/alignment/Alignment.cpp)
//   spb200 search    <queryDB> <targetDB> <alnDB> [flags]           both steps with the hit lists kept in HBM
//   spb200 extractframes <nuclSeqDB> <outDB> [--min-length 9]        extractorfs --orf-start-mode 1 + translatenucs fused on the device:
//                                                                   <outDB>_nucl_orf(+_h) and the amino-acid query DB <outDB>(+_h)
//   spb200 createindex <targetDB> <tmpDir> [flags]                  writes <targetDB>.spbidx (tantan-masked residues; the k-mer
//                                                                   table itself is rebuilt on the device), picked up by the above
// and the set-level stages that consume the alignment DB in predictmatch (data/predictmatch.sh:57-116; setscore.h):
//   spb200 besthitperset <querySetDB> <targetSetDB> <alnDB> <outDB> [--simple-best-hit 1]
//   spb200 mergeresultsbyset <querySetDB>_set_to_member <resultDB> <outDB>
//   spb200 combinescore <querySetDB> <targetSetDB> <mergedDB> <outDB> <tmpDir>
//   spb200 filtermatchbyfdr <cScoreDB> <controlcScoreDB> <matchDB> [--fdr 0.05 --report-fdr 0]
//   spb200 predictmatch <querySetDB> <targetSetDB> <controlTargetSetDB> <matchDB> <tmpDir> [search flags, --fdr ...]
//       search against the target and the reversed-ORF control setDB on the GPU, then the four stages in memory;
//       writes the `match` DB of predictmatch.sh; --perform-nucl-aln 1 (the reference's default) re-scores every alignment on
//       the <db>_nucl_orf entries on the device and combines the E-values before the set stages.
// and the two modules of that nucleotide branch (data/predictmatch.sh:35-45):
//   spb200 proteinaln2nucl <queryDB>_nucl_orf <targetDB>_nucl_orf <queryDB> <targetDB> <alnDB> <outDB>
//   spb200 combineprotnuclaln <protAlnDB> <nuclAlnDB> <outDB>
// Host-only helpers used by the CPU tests: dumpmatrix, dbcopy, showparams.
// Errors follow the reference: message on stderr, exit code 1 (Debug(Debug::ERROR) ... EXIT(EXIT_FAILURE)).
#include <ctype.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <atomic>
#include <chrono>
#include <stdexcept>
#include <exception>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../../include/spb200.h"
#include "db_io.h"
#include "fastfmt.h"
#include "setscore.h"
#include "submat.h"

using namespace spbhost;

namespace {

struct Fail : std::runtime_error {
    explicit Fail(const std::string &m) : std::runtime_error(m) {}
};

// ---------------------------------------------------------------------------------------------------------
// parameters (names, defaults and meaning of M/commons/Parameters.cpp; SpacePHARER's predictmatch defaults are passed
// explicitly by the workflow, the values below are MMseqs2's own defaults)
// ---------------------------------------------------------------------------------------------------------
struct Params {
    std::vector<std::string> db;
    std::string subMat = "nucl:nucleotide.out,aa:blosum62.out", seedSubMat = "nucl:nucleotide.out,aa:VTML80.out";
    float sensitivity = 5.7f;
    int kmerSize = 0, kmerScore = INT_MAX, maxSeqs = 300, minUngapped = 15, mask = 1, compBias = 1, diagScore = 1;
    int split = 0, splitMode = 2, exactKmer = 0, spacedMode = 1, covMode = 0, addSelf = 0, maskLower = 0;
    std::string spacedPattern;
    double evalue = 1e-3, seqIdThr = 0, covThr = 0;
    int addBacktrace = 0, alignmentMode = 0, altAli = 0, realign = 0, seqIdMode = 0, alnLenThr = 0;
    int maxRejected = INT_MAX, maxAccept = INT_MAX;
    std::string gapOpen = "nucl:5,aa:11", gapExtend = "nucl:2,aa:1";
    double scoreBias = 0.0, realignScoreBias = -0.2;
    int realignMaxSeqs = INT_MAX;
    int threads = 1, verbosity = 3, compressed = 0;
    // spb200-specific
    // flags whose non-default values would change the reference's result but are not implemented: parsed so that they can be
    // refused (check_unimplemented_flags) instead of being silently ignored
    std::string alphSize = "nucl:5,aa:21", taxonList;
    int maxSeqLen = 65535;
    int device = 0, binSize = 0, byteLanes = 32, wordLanes = 16;
    bool realignOk = false;     // set by the align module: --realign 1 is served there (two passes over the same spb_align call)
    bool altAliOk = false;      // set by the align module: --alt-ali N is served there (host rounds over masked target copies)
    bool targetSplitOk = false; // set by the prefilter module: --split N --split-mode 0 (the reference's own target split) is served
    int gpus = 1;              // search / predictmatch: devices device .. device + gpus - 1, k-mer table sharded over them (index split)
    int lanes = 2;             // search / predictmatch: contexts per device that share its index and overlap their batches (one host thread each)
    int oversubscribe = 0;     // 1: with fewer devices than --gpus, several shard contexts share a device (same code path; tests)
    std::string gumbel, gumbelNucl;
    long batch = 250000;       // frames per GPU pass: formatting of batch k overlaps the kernels of batch k+1 (1 M spacers: 4.2 -> 3.7 s)
    // set-level stages (M/commons/Parameters.cpp:217,2277; src/commons/LocalParameters.h:75-77,171-176)
    int simpleBestHit = 1, reportFdr = 0, performNuclAln = 1, keepIntermediate = 0;
    // 1: no k-mer stage, Smith-Waterman of every query against every target (BASELINE config 5; the result of the reference's
    // `align` over an all-pairs prefilter DB -- the reference's own flag of this name is the sliced search)
    int exhaustive = 0;
    float fdr = 0.05f;
};

struct FlagDef { const char *name; int kind; void *dst; };   // kind: 0 int, 1 float, 2 double, 3 string, 4 ignored

std::vector<FlagDef> flag_table(Params &p) {
    static int ignoredInt;
    (void)ignoredInt;
    return {
        {"--sub-mat", 3, &p.subMat}, {"--seed-sub-mat", 3, &p.seedSubMat}, {"-s", 1, &p.sensitivity}, {"-k", 0, &p.kmerSize},
        {"--k-score", 0, &p.kmerScore}, {"--max-seqs", 0, &p.maxSeqs}, {"--min-ungapped-score", 0, &p.minUngapped},
        {"--mask", 0, &p.mask}, {"--mask-lower-case", 0, &p.maskLower}, {"--comp-bias-corr", 0, &p.compBias},
        {"--diag-score", 0, &p.diagScore}, {"--split", 0, &p.split}, {"--split-mode", 0, &p.splitMode},
        {"--exact-kmer-matching", 0, &p.exactKmer}, {"--spaced-kmer-mode", 0, &p.spacedMode},
        {"--spaced-kmer-pattern", 3, &p.spacedPattern}, {"--cov-mode", 0, &p.covMode}, {"-c", 2, &p.covThr},
        {"--add-self-matches", 0, &p.addSelf}, {"-e", 2, &p.evalue}, {"--min-seq-id", 2, &p.seqIdThr},
        {"-a", 0, &p.addBacktrace}, {"--alignment-mode", 0, &p.alignmentMode}, {"--alt-ali", 0, &p.altAli},
        {"--realign", 0, &p.realign}, {"--seq-id-mode", 0, &p.seqIdMode}, {"--min-aln-len", 0, &p.alnLenThr},
        {"--max-rejected", 0, &p.maxRejected}, {"--max-accept", 0, &p.maxAccept}, {"--gap-open", 3, &p.gapOpen},
        {"--gap-extend", 3, &p.gapExtend}, {"--score-bias", 2, &p.scoreBias}, {"--threads", 0, &p.threads},
        {"-v", 0, &p.verbosity}, {"--compressed", 0, &p.compressed},
        // accepted and without influence on this path (profile / nucleotide / memory-management flags of the same modules)
        {"--alph-size", 3, &p.alphSize}, {"--max-seq-len", 0, &p.maxSeqLen}, {"--split-memory-limit", 4, nullptr},
        {"--db-load-mode", 4, nullptr}, {"--pca", 4, nullptr}, {"--pcb", 4, nullptr}, {"--alignment-output-mode", 4, nullptr},
        {"--wrapped-scoring", 4, nullptr}, {"--realign-score-bias", 2, &p.realignScoreBias}, {"--realign-max-seqs", 0, &p.realignMaxSeqs},
        {"--zdrop", 4, nullptr}, {"--local-tmp", 4, nullptr}, {"--taxon-list", 3, &p.taxonList}, {"--exhaustive-search-filter", 4, nullptr},
        {"--exhaustive-search", 0, &p.exhaustive},
        {"--simple-best-hit", 0, &p.simpleBestHit}, {"--fdr", 1, &p.fdr}, {"--report-fdr", 0, &p.reportFdr},
        {"--perform-nucl-aln", 0, &p.performNuclAln}, {"--tax-fdr", 4, nullptr}, {"--remove-tmp-files", 4, nullptr},
        // spb200-specific
        {"--keep-intermediate", 0, &p.keepIntermediate},
        {"--device", 0, &p.device}, {"--gpus", 0, &p.gpus}, {"--lanes", 0, &p.lanes}, {"--oversubscribe", 0, &p.oversubscribe}, {"--bin-size", 0, &p.binSize}, {"--byte-lanes", 0, &p.byteLanes}, {"--word-lanes", 0, &p.wordLanes},
        {"--gumbel", 3, &p.gumbel}, {"--gumbel-nucl", 3, &p.gumbelNucl}, {"--batch", 0, nullptr},
    };
}

void parse_args(int argc, const char **argv, int first, size_t nDb, Params &p) {
    std::vector<FlagDef> table = flag_table(p);
    for (int i = first; i < argc; i++) {
        const std::string a = argv[i];
        if (a.size() > 1 && a[0] == '-' && !(a.size() > 1 && (isdigit((unsigned char)a[1]) || a[1] == '.'))) {
            const FlagDef *def = nullptr;
            for (const FlagDef &f : table)
                if (a == f.name) def = &f;
            if (!def) throw Fail("Unrecognized parameter \"" + a + "\"");
            if (i + 1 >= argc) throw Fail("Missing argument " + a);
            const char *v = argv[++i];
            char *end = nullptr;
            if (a == "--batch") { p.batch = strtol(v, &end, 10); if (*end || p.batch < 1) throw Fail("Error in argument " + a); continue; }
            switch (def->kind) {
                case 0: {
                    long x = strtol(v, &end, 10);
                    if (end == v || *end) throw Fail("Error in argument " + a);
                    *(int *)def->dst = (int)std::max<long>(INT_MIN, std::min<long>(INT_MAX, x));
                    break;
                }
                case 1: *(float *)def->dst = strtof(v, &end); if (end == v || *end) throw Fail("Error in argument " + a); break;
                case 2: *(double *)def->dst = strtod(v, &end); if (end == v || *end) throw Fail("Error in argument " + a); break;
                case 3: *(std::string *)def->dst = v; break;
                default: break;
            }
        } else {
            p.db.push_back(a);
        }
    }
    if (p.db.size() != nDb) throw Fail("expected " + std::to_string(nDb) + " database arguments, got " + std::to_string(p.db.size()));
    if (p.compressed != 0 && p.compressed != 1) throw Fail("Error in argument --compressed");
    DBWriter::setCompressedDefault(p.compressed == 1);   // every DB this module writes (DBWriter.cpp:32-41: WRITER_COMPRESSED_MODE)
    if (p.threads < 1) p.threads = 1;
}

// Prefiltering::getKmerThreshold (M/prefiltering/Prefiltering.cpp:987-1021), sequence (non-profile) search
int kmer_threshold(float sensitivity, int kmerScore, int kmerSize) {
    double best = kmerScore;
    if (kmerScore == INT_MAX) {
        if (kmerSize == 6) {
            float base = 163.2f;
            best = base - (sensitivity * 8.917);
        } else {
            throw Fail("The k-mer size " + std::to_string(kmerSize) + " is not valid.");
        }
    }
    return (int)best;
}

// QueryMatcher::initDiagonalMatcher (M/prefiltering/QueryMatcher.cpp:383-411) with Util::getL2CacheSize
int bin_size_for(size_t dbsize) {
    uint64_t l2 = 262144;
#if defined(_SC_LEVEL2_CACHE_SIZE)
    long c = sysconf(_SC_LEVEL2_CACHE_SIZE);
    if (c > 0) l2 = (uint64_t)c;
#endif
    for (int b = 2; b <= 1024; b *= 2)
        if (dbsize / (size_t)b < l2) return b;
    return 2048;
}

// ALP parameter sets (M/alignment/EvalueComputation.h:56-75 table; VTML40 16/2 as estimated by the reference with seed 42,
// SURVEY.md 8a-5).  Other combinations must come from the reference's estimator via --gumbel.
bool gumbel_for(const std::string &mat, int go, int ge, const std::string &override_, spb_gumbel *g) {
    double v[12];
    if (!override_.empty()) {
        const char *p = override_.c_str();
        for (int i = 0; i < 12; i++) {
            char *e = nullptr;
            v[i] = strtod(p, &e);
            if (e == p) return false;
            p = (*e == ',') ? e + 1 : e;
        }
    } else if (mat == "VTML40.out" && go == 16 && ge == 2) {
        const double t[12] = {0.34420393530174209, 0.27487374979561635, 0.22018895496520485, 0.026237512825550646,
                              0.18481499377626534, 1.7126891360391483, 0.22018895496520485, 0.026237512825550646,
                              0.18481499377626534, 1.7126891360391483, 0.18283871043913499, 1.7838353361758408};
        memcpy(v, t, sizeof(v));
    } else if (mat == "nucleotide.out" && go == 10 && ge == 2) {
        // estimated by the reference at start-up like the VTML40 set (EvalueComputation.h:48-154, seed 42); probed from the
        // reference build for proteinaln2nucl's NucleotideMatrix(nucleotide.out, 1.0, 0.0) with --gap-open nucl:10 --gap-extend 2
        const double t[12] = {0.63361984397325744, 0.42150719938226583, 0.68666204026925159, 0.18923575026627315,
                              0.71225621255786731, 2.9058945364387716, 0.68666204026925159, 0.18923575026627315,
                              0.71225621255786731, 2.9058945364387716, 0.7125673897037631, 2.8984262849372726};
        memcpy(v, t, sizeof(v));
    } else if (mat == "nucleotide.out" && go == 7 && ge == 1) {   // EvalueComputation.h:57-62
        const double t[12] = {1.0960171987681839, 0.33538787507026158, 2.0290734315292083, -0.46514786408422282,
                              5.0543294182155085, 15.130999712620039, 2.0290734315292083, -0.46514786408422282,
                              5.0543294182155085, 15.130999712620039, 5.0543962679167036, 15.129930117400917};
        memcpy(v, t, sizeof(v));
    } else if (mat == "blosum62.out" && go == 11 && ge == 1) {
        const double t[12] = {0.27359865037097330642, 0.044620920658722244834, 1.5938724404943873658, -19.959867650284412122,
                              30.455610143099914211, -622.28684628915891608, 1.5938724404943873658, -19.959867650284412122,
                              30.455610143099914211, -622.28684628915891608, 29.602444874818868215, -601.81087985041381216};
        memcpy(v, t, sizeof(v));
    } else {
        return false;
    }
    g->lambda = v[0]; g->K = v[1]; g->a_I = v[2]; g->b_I = v[3]; g->alpha_I = v[4]; g->beta_I = v[5];
    g->a_J = v[6]; g->b_J = v[7]; g->alpha_J = v[8]; g->beta_J = v[9]; g->sigma = v[10]; g->tau = v[11];
    return true;
}

void check(int rc, const char *what) {
    if (rc != 0) throw Fail(std::string(what) + ": " + spb_last_error());
}

// ---------------------------------------------------------------------------------------------------------
// sequence DBs -> numeric residues
// ---------------------------------------------------------------------------------------------------------
struct SeqSet {
    std::vector<uint8_t> res;
    std::vector<uint64_t> off;
    std::vector<uint32_t> keys;   // DB key of every sequence (position = id)
    std::vector<uint8_t> lower;   // 1 where the DB holds a lower-case letter (only filled for --mask-lower-case 1)
};

// ids in LINEAR_ACCESS order (what Prefiltering / IndexBuilder use for both DBs, Prefiltering.cpp:163-164,187-188)
void load_sequences(const DBReader &db, const SubMat &m, SeqSet &s, bool wantLower = false) {
    const size_t n = db.size();
    s.off.assign(n + 1, 0);
    s.keys.resize(n);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        const DBEntry &e = db.linear(i);
        const uint64_t len = e.length >= 2 ? e.length - 2 : 0;   // "residues\n\0"
        s.off[i] = total;
        s.keys[i] = e.key;
        total += len;
    }
    s.off[n] = total;
    s.res.resize(total + 1);
    if (wantLower) s.lower.assign(total + 1, 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        const DBEntry &e = db.linear((size_t)i);
        const size_t len = (size_t)(s.off[i + 1] - s.off[i]);
        m.encode(db.data(e), len, s.res.data() + s.off[i]);
        if (wantLower) {
            const char *c = db.data(e);
            for (size_t k = 0; k < len; k++) s.lower[s.off[i] + k] = islower((unsigned char)c[k]) ? 1 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// row formatting
// ---------------------------------------------------------------------------------------------------------
// Util::fastSeqIdToBuffer (M/commons/Util.cpp:278-307): the caller overwrites the last byte with '\t', hence "1.00"
size_t seqid_to_buffer(float seqId, char *b) {
    char *s = b;
    if (seqId == 1.0f) {
        memcpy(b, "1.00", 4);
        return 4;
    }
    *b++ = '0';
    *b++ = '.';
    if (seqId < 0.10) *b++ = '0';
    if (seqId < 0.01) *b++ = '0';
    b = put_int(b, (int)(seqId * 1000));
    return (size_t)(b - s);
}

// Matcher::compressAlignment (M/alignment/Matcher.cpp:211-230): run lengths, counts always printed
void compress_cigar(const char *ops, uint32_t n, std::string &out) {
    char state = 'M';
    unsigned count = 0;
    char tmp[16];
    auto run = [&]() {
        char *w = put_uint(tmp, count);
        *w++ = state;
        out.append(tmp, (size_t)(w - tmp));
    };
    for (uint32_t i = 0; i < n; i++) {
        if (ops[i] != state) {
            run();
            state = ops[i];
            count = 1;
        } else {
            count++;
        }
    }
    run();
}

struct Engine {
    spb_ctx *ctx = nullptr;
    SubMat seed, ungapped, align;
    ~Engine() { if (ctx) spb_ctx_destroy(ctx); }
    // host part (matrices) and device part (context) separately: a search reads and masks its DBs while the CUDA context
    // of every device comes up in the background
    void load_matrices(const Params &p) {
        seed.load(p.seedSubMat, 8.0, -0.2);        // Prefiltering.cpp:64
        ungapped.load(p.subMat, 2.0, -0.2);        // Prefiltering.cpp:65
        align.load(p.subMat, 2.0, p.scoreBias);    // Alignment.cpp:159
    }
    void create_context(const Params &p) {
        check(spb_ctx_create(p.device, &ctx), "spb_ctx_create");
        check(spb_set_matrices(ctx, seed.sub, seed.pback, ungapped.sub, align.sub, align.pback), "spb_set_matrices");
    }
    void open(const Params &p) {
        load_matrices(p);
        create_context(p);
    }
};

void log_line(const Params &p, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void log_line(const Params &p, const char *fmt, ...) {
    if (p.verbosity < 3) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
    fputc('\n', stdout);
}

// ---------------------------------------------------------------------------------------------------------
// prefilter
// ---------------------------------------------------------------------------------------------------------
// ---- persistent index (the reference: createindex -> <targetDB>.idx, PrefilteringIndexReader.cpp) ----
// The reference serialises the whole k-mer table because building it costs CPU-seconds.  Here the table is rebuilt on the device
// in ~20 ms (index.cu) -- faster than reading 1.7 GB back -- so the cache holds what is expensive on the HOST side: the
// tantan-masked residues (SEQINDEXDATA of the reference's index) plus the parameters they depend on.  Layout of
// <targetDB>.spbidx: IdxHeader, keys u32[T], offsets u64[T+1], masked residues u8[N].
// phase wall clock of the host binary on stderr (SPB_TIMING=1; development aid)
struct PhaseClock;
extern PhaseClock g_clock;
struct PhaseClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool on = getenv("SPB_TIMING") != nullptr;
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[spb200 timing] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};

PhaseClock g_clock;

struct IdxHeader {
    char magic[8];            // "SPBIDX1"
    uint32_t version, mask, maskLower, pad;
    uint64_t nTargets, nResidues, srcDataSize, checksum;   // checksum: FNV-1a over keys, offsets and the UNMASKED residues
    char seedMatrix[64];
};

uint64_t fnv1a(const void *d, size_t n, uint64_t h) {
    const unsigned char *p = (const unsigned char *)d;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

uint64_t seqset_checksum(const SeqSet &T) {
    uint64_t h = 1469598103934665603ULL;
    h = fnv1a(T.keys.data(), T.keys.size() * 4, h);
    h = fnv1a(T.off.data(), T.off.size() * 8, h);
    h = fnv1a(T.res.data(), (size_t)T.off.back(), h);
    if (!T.lower.empty()) h = fnv1a(T.lower.data(), (size_t)T.off.back(), h);   // --mask-lower-case 1: the case is part of the input
    return h;
}

void fill_idx_header(const Params &p, const Engine &E, const SeqSet &T, uint64_t srcDataSize, IdxHeader *h) {
    memset(h, 0, sizeof(*h));
    memcpy(h->magic, "SPBIDX1", 8);
    h->version = 1;
    h->mask = (uint32_t)p.mask;
    h->maskLower = (uint32_t)p.maskLower;
    h->nTargets = T.keys.size();
    h->nResidues = T.off.back();
    h->srcDataSize = srcDataSize;
    h->checksum = seqset_checksum(T);
    snprintf(h->seedMatrix, sizeof(h->seedMatrix), "%s", E.seed.name.c_str());
}

void mask_targets(Engine &E, const Params &p, const SeqSet &T, std::vector<uint8_t> &masked, uint64_t *nMasked) {
    masked.assign(T.res.begin(), T.res.end());
    *nMasked = 0;
    if (p.mask)   // IndexBuilder.cpp:137-149
        check(spb_tantan_mask(E.seed.prob, E.seed.pback, masked.data(), T.off.data(), (uint32_t)T.keys.size(), p.threads, nMasked), "spb_tantan_mask");
    if (p.maskLower) {   // IndexBuilder.cpp:151-160: after tantan, every lower-case letter of the DB becomes X in the index
        if (T.lower.size() != T.res.size()) throw Fail("--mask-lower-case 1: the target sequences were read without their case");
        const uint8_t x = E.seed.aa2num[(unsigned char)'X'];
        const size_t n = (size_t)T.off.back();
        uint64_t cnt = 0;
#pragma omp parallel for reduction(+ : cnt) schedule(static)
        for (long i = 0; i < (long)n; i++)
            if (T.lower[(size_t)i]) { masked[(size_t)i] = x; cnt++; }     // (counted like the reference: also where tantan masked already)
        *nMasked += cnt;
    }
}

// true if <name>.spbidx exists and belongs to exactly this target set and these masking parameters
bool load_index_cache(const std::string &name, const Params &p, const Engine &E, const SeqSet &T, uint64_t srcDataSize,
                      std::vector<uint8_t> &masked) {
    FILE *f = fopen((name + ".spbidx").c_str(), "rb");
    if (!f) return false;
    IdxHeader want, got;
    fill_idx_header(p, E, T, srcDataSize, &want);
    bool ok = fread(&got, sizeof(got), 1, f) == 1 && memcmp(&got, &want, sizeof(got)) == 0;
    if (ok) {
        ok = fseek(f, (long)(sizeof(got) + T.keys.size() * 4 + T.off.size() * 8), SEEK_SET) == 0;
        masked.resize(T.res.size());
        ok = ok && fread(masked.data(), 1, (size_t)want.nResidues, f) == (size_t)want.nResidues;
    }
    fclose(f);
    if (!ok) log_line(p, "Index cache %s.spbidx does not match this database / these parameters: ignored", name.c_str());
    return ok;
}

// masked target residues: from the <targetDB>.spbidx cache when it validates, else tantan on the host (once per run, whatever the
// number of GPUs)
bool masked_targets(Engine &E, const Params &p, const SeqSet &T, const std::string &tName, uint64_t srcDataSize,
                    std::vector<uint8_t> &masked, uint64_t *nMasked) {
    const bool cached = !tName.empty() && load_index_cache(tName, p, E, T, srcDataSize, masked);
    if (!cached) mask_targets(E, p, T, masked, nMasked);
    return cached;
}

void build_index_from_masked(Engine &E, const Params &p, const SeqSet &T, const std::vector<uint8_t> &masked, int kmerThr, bool cached,
                             uint64_t nMasked, const std::string &tName) {
    const std::string pattern = p.spacedPattern.empty() ? (p.spacedMode ? "11011101" : "111111") : p.spacedPattern;
    // index-side threshold (Prefiltering.cpp:510-513): only for exact matching or when the k-mer threshold is small enough
    // (an exact-matching search indexes every k-mer: localKmerThr = 0)
    check(spb_index_build(E.ctx, masked.data(), T.off.data(), (uint32_t)T.keys.size(), pattern.c_str(), p.exactKmer ? 0 : kmerThr), "spb_index_build");
    spb_index_stats st;
    check(spb_index_stats_get(E.ctx, &st), "spb_index_stats_get");
    if (cached)
        log_line(p, "Index table: %llu targets, %llu residues (masked residues from %s.spbidx), %llu postings, %.1f ms on the device",
                 (unsigned long long)st.n_targets, (unsigned long long)st.n_residues, tName.c_str(), (unsigned long long)st.n_postings, st.build_ms);
    else
        log_line(p, "Index table: %llu targets, %llu residues (%llu masked), %llu postings, %.1f ms on the device",
                 (unsigned long long)st.n_targets, (unsigned long long)st.n_residues, (unsigned long long)nMasked,
                 (unsigned long long)st.n_postings, st.build_ms);
}

void build_target_index(Engine &E, const Params &p, const SeqSet &T, int kmerThr, const std::string &tName = std::string(),
                        uint64_t srcDataSize = 0) {
    std::vector<uint8_t> masked;
    uint64_t nMasked = 0;
    const bool cached = masked_targets(E, p, T, tName, srcDataSize, masked, &nMasked);
    build_index_from_masked(E, p, T, masked, kmerThr, cached, nMasked, tName);
}

// Flags of the prefilter / align modules that change the reference's output and that this path does not implement keep the
// reference's default here; any other value is refused (a drop-in must not exit 0 with a different DB).  Flags that stay in
// the "accepted, no influence" class are inert under the settings this path already requires: --cov-mode (needs -c > 0),
// --split-mode (needs --split > 1), --realign-score-bias / --realign-max-seqs (need --realign 1), --zdrop (nucleotide
// alignment only), --alignment-output-mode 0, --pca / --pcb (profiles), --db-load-mode, --split-memory-limit (memory management).
void check_unimplemented_flags(const Params &p) {
    // a DB searched against itself: the reference lists the identity hit first and scores it with scoreIdentical
    // (Prefiltering.cpp:539-552,808-820; Alignment.cpp:62,384-394).  SpacePHARER never does that (spacers vs phages); refused.
    if (p.db.size() >= 2 && p.db[0] == p.db[1])
        throw Fail("query and target database are the same: the self-match handling of the reference is not supported on this path");
    if (p.maskLower != 0 && p.maskLower != 1) throw Fail("Error in argument --mask-lower-case");
    {
        // "--alph-size nucl:5,aa:21" or a single number for both
        std::string aa = p.alphSize;
        const size_t at = aa.find("aa:");
        if (at != std::string::npos) aa = aa.substr(at + 3);
        const size_t comma = aa.find(',');
        if (comma != std::string::npos) aa = aa.substr(0, comma);
        if (aa.find("nucl:") == 0) aa = "21";
        if (atoi(aa.c_str()) != 21) throw Fail("--alph-size " + p.alphSize + ": only the full amino-acid alphabet (aa:21) is supported");
    }
    if (!p.taxonList.empty()) throw Fail("--taxon-list is not supported");
    if (p.maxSeqLen < 1) throw Fail("Error in argument --max-seq-len");
}

// --max-seq-len: the reference splits / truncates longer sequences (Sequence.cpp); here longer entries are refused
void check_seq_len(const Params &p, const std::vector<uint64_t> &off, const char *what) {
    for (size_t i = 0; i + 1 < off.size(); i++)
        if (off[i + 1] - off[i] > (uint64_t)p.maxSeqLen)
            throw Fail(std::string(what) + " sequence " + std::to_string(i) + " is longer than --max-seq-len " + std::to_string(p.maxSeqLen) +
                       "; splitting long sequences is not supported on this path");
}

void check_search_params(const Params &p) {
    check_unimplemented_flags(p);
    if (p.kmerSize != 0 && p.kmerSize != 6) throw Fail("The k-mer size " + std::to_string(p.kmerSize) + " is not valid. (spb200 implements k = 6)");
    if (p.diagScore != 1) throw Fail("--diag-score 0 is not supported");
    if (p.exactKmer != 0 && p.exactKmer != 1) throw Fail("Error in argument --exact-kmer-matching");
    // --split N with --split-mode 1 (query split) or 2 (auto: query split whenever the memory suffices, Prefiltering.cpp:258-290) gives
    // the unsplit result and needs nothing here; --split-mode 0 is the reference's lossy target split (prefilter module only)
    if (p.split < 0 || p.splitMode < 0 || p.splitMode > 2) throw Fail("Error in argument --split / --split-mode");
    if (p.split > 1 && p.splitMode == 0 && !p.targetSplitOk)
        throw Fail("--split > 1 with --split-mode 0 (the reference's lossy target split) is only served by the prefilter module; spread one "
                   "DB over several GPUs with --gpus N (k-mer table sharded over the devices, unsplit result)");
    if (p.gpus < 1 || p.gpus > 8) throw Fail("--gpus must be 1..8");
    if (p.lanes < 1 || p.lanes > 8) throw Fail("--lanes must be 1..8");
    if (p.addSelf != 0) throw Fail("--add-self-matches is not supported on this path");
    if (p.covThr < 0 || p.covThr > 1) throw Fail("Error in argument -c");
    if (p.covMode < 0 || p.covMode > 5) throw Fail("Error in argument --cov-mode");
}

// Util::canBeCovered (M/commons/Util.cpp:533-550), float arithmetic as there; modes by value (Parameters.h:246-251)
bool can_be_covered(float covThr, int covMode, float qLen, float tLen) {
    switch (covMode) {
    case 0: return (qLen / tLen >= covThr) && (tLen / qLen >= covThr);
    case 1: return (qLen / tLen) >= covThr;
    case 2: return (tLen / qLen) >= covThr;
    case 3: return ((tLen / qLen) >= covThr) && (tLen / qLen) <= 1.0;
    case 4: return ((qLen / tLen) >= covThr) && (qLen / tLen) <= 1.0;
    case 5: return (std::min(tLen, qLen) / std::max(tLen, qLen)) >= covThr;
    default: return true;
    }
}

// DBReader::decomposeDomainByAminoAcid (M/commons/DBReader.cpp:1171-1213) for every rank: entries in id order, entry length =
// residues + 2 as in the index; a split is closed once it holds ceil(dataSize / world) bytes
std::vector<std::pair<size_t, size_t>> decompose_by_amino_acid(const SeqSet &T, size_t world) {
    const size_t n = T.keys.size();
    std::vector<std::pair<size_t, size_t>> out(world, std::make_pair((size_t)0, (size_t)0));
    if (world == 1) { out[0] = std::make_pair((size_t)0, n); return out; }
    if (n <= world) {
        for (size_t r = 0; r < world; r++) out[r] = std::make_pair(r < n ? r : (size_t)0, r < n ? (size_t)1 : (size_t)0);
        return out;
    }
    const uint64_t dataSize = T.off.back() + 2 * (uint64_t)n;
    const uint64_t chunk = (uint64_t)ceil((double)dataSize / (double)world);
    size_t rank = 0;
    uint64_t acc = 0;
    for (size_t i = 0; i < n; i++) {
        if (acc >= chunk) { acc = 0; rank++; }
        acc += T.off[i + 1] - T.off[i] + 2;
        out[rank].second++;
    }
    size_t start = 0;
    for (size_t r = 0; r < world; r++) { out[r].first = start; start += out[r].second; }
    return out;
}

struct PrefHit { uint32_t key; int32_t score, diag; };

// `prefilter --split N --split-mode 0`: MMseqs2's own target split (Prefiltering.cpp:258-362 set-up, :712-733 one index per split,
// :364-481 mergeTargetSplits).  It is LOSSY by design: every split keeps max-seqs / N + 4 sqrt(max-seqs / N) hits per query and the
// merge only concatenates and sorts, so the DB differs from the unsplit one -- a drop-in called with these flags has to write
// what the reference writes.  The splits run one after the other on the context (index rebuilt per split, ~ms on the device).
int run_prefilter_target_split(Engine &E, const Params &p, const SeqSet &Q, const SeqSet &T, int kmerThr, const std::string &tName,
                               uint64_t srcDataSize) {
    std::vector<uint8_t> masked;
    uint64_t nMasked = 0;
    masked_targets(E, p, T, tName, srcDataSize, masked, &nMasked);
    const size_t nq = Q.keys.size(), split = (size_t)p.split;
    size_t maxRes = std::min<size_t>(T.keys.size(), (size_t)p.maxSeqs);
    maxRes = std::max<size_t>(1, maxRes / split + (size_t)(4 * sqrt((double)maxRes / (double)split)));      // Prefiltering.cpp:343-346
    log_line(p, "Target split mode. Searching through %zu splits, %zu hits per query and split", split, maxRes);
    const std::string pattern = p.spacedPattern.empty() ? (p.spacedMode ? "11011101" : "111111") : p.spacedPattern;
    const bool covFilter = p.covThr > 0.0 && (p.covMode == 0 || p.covMode == 2 || p.covMode == 5);
    std::vector<std::vector<PrefHit>> rows(nq);
    uint64_t nHits = 0;
    double kernelMs = 0;
    const auto ranges = decompose_by_amino_acid(T, split);
    for (size_t sIdx = 0; sIdx < split; sIdx++) {
        const size_t from = ranges[sIdx].first, cnt = ranges[sIdx].second;
        if (cnt == 0) continue;
        std::vector<uint64_t> toff(cnt + 1);
        for (size_t i = 0; i <= cnt; i++) toff[i] = T.off[from + i] - T.off[from];
        check(spb_index_build(E.ctx, masked.data() + T.off[from], toff.data(), (uint32_t)cnt, pattern.c_str(), p.exactKmer ? 0 : kmerThr),
              "spb_index_build");
        spb_prefilter_params pp;
        pp.kmer_thr = kmerThr;
        pp.max_seqs = (int32_t)maxRes;
        pp.min_diag_score = p.minUngapped;
        pp.bin_size = p.binSize ? p.binSize : bin_size_for(cnt);      // QueryMatcher sees the split's dbSize
        pp.comp_bias = p.compBias;
        pp.id_base = 0;                                               // ids are local to the split (Prefiltering.cpp:827)
        pp.emit_trunc_keys = 0;
        pp.exact_kmer = p.exactKmer;
        for (size_t b0 = 0, b1 = 0; b0 < nq; b0 = b1) {
            b1 = std::min(nq, b0 + (size_t)p.batch);
            std::vector<uint64_t> off(b1 - b0 + 1);
            for (size_t i = b0; i <= b1; i++) off[i - b0] = Q.off[i] - Q.off[b0];
            spb_prefilter_result *r = nullptr;
            check(spb_prefilter(E.ctx, Q.res.data() + Q.off[b0], off.data(), (uint32_t)(b1 - b0), &pp, &r), "spb_prefilter");
            kernelMs += r->kernel_ms;
#pragma omp parallel for num_threads(p.threads) schedule(static)
            for (long i = 0; i < (long)(b1 - b0); i++) {
                const float qLen = (float)(Q.off[b0 + (size_t)i + 1] - Q.off[b0 + (size_t)i]);
                std::vector<PrefHit> &dst = rows[b0 + (size_t)i];
                for (uint64_t h = r->offsets[i]; h < r->offsets[i + 1]; h++) {
                    const spb_hit &x = r->hits[h];
                    const size_t tid = from + x.id;
                    if (covFilter && !can_be_covered((float)p.covThr, p.covMode, qLen, (float)(T.off[tid + 1] - T.off[tid]))) continue;
                    dst.push_back(PrefHit{T.keys[tid], x.score, x.diag});
                }
            }
            spb_free_prefilter_result(r);
        }
        log_line(p, "Process prefiltering step %zu of %zu: targets %zu .. %zu", sIdx + 1, split, from, from + cnt);
    }
    // mergeTargetSplits: the lists of a query concatenated, sorted by hit_t::compareHitsByScoreAndId; never compressed
    DBWriter::setCompressedDefault(false);
    DBWriter out;
    out.open(p.db[2], DBTYPE_PREFILTER_RES, p.threads);
#pragma omp parallel num_threads(p.threads)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        std::string buf;
        char tmp[64];
#pragma omp for schedule(static) reduction(+ : nHits)
        for (long i = 0; i < (long)nq; i++) {
            std::vector<PrefHit> &v = rows[(size_t)i];
            std::sort(v.begin(), v.end(), [](const PrefHit &a, const PrefHit &b) {
                if (abs(a.score) != abs(b.score)) return abs(a.score) > abs(b.score);
                return a.key < b.key;
            });
            buf.clear();
            for (const PrefHit &x : v) {
                char *w = put_uint(tmp, x.key); *w++ = '\t';
                w = put_int(w, x.score); *w++ = '\t';
                w = put_int(w, x.diag); *w++ = '\n';
                buf.append(tmp, (size_t)(w - tmp));
            }
            nHits += v.size();
            out.write(th, Q.keys[(size_t)i], buf.data(), buf.size());
        }
    }
    out.close();
    log_line(p, "%llu prefilter hits for %zu queries over %zu target splits; match kernels %.3f ms", (unsigned long long)nHits, nq, split, kernelMs);
    return 0;
}

int run_prefilter(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 3, p);
    p.targetSplitOk = true;
    check_search_params(p);
    Engine E;
    E.open(p);
    DBReader qdb, tdb;
    qdb.open(p.db[0]);
    tdb.open(p.db[1]);
    log_line(p, "Query database size: %zu type: Aminoacid", qdb.size());
    log_line(p, "Target database size: %zu type: Aminoacid", tdb.size());
    SeqSet Q, T;
    load_sequences(qdb, E.seed, Q);
    load_sequences(tdb, E.seed, T, p.maskLower == 1);
    check_seq_len(p, Q.off, "query");
    check_seq_len(p, T.off, "target");
    const int kmerThr = kmer_threshold(p.sensitivity, p.kmerScore, 6);
    log_line(p, "k-mer similarity threshold: %d", kmerThr);
    if (p.split > 1 && p.splitMode == 0) return run_prefilter_target_split(E, p, Q, T, kmerThr, p.db[1], tdb.dataSize());
    build_target_index(E, p, T, kmerThr, p.db[1], tdb.dataSize());

    spb_prefilter_params pp;
    pp.kmer_thr = kmerThr;
    pp.max_seqs = p.maxSeqs;
    pp.min_diag_score = p.minUngapped;
    pp.bin_size = p.binSize ? p.binSize : bin_size_for(T.keys.size());
    pp.comp_bias = p.compBias;
    pp.id_base = 0;
    pp.emit_trunc_keys = 0;
    pp.exact_kmer = p.exactKmer;

    DBWriter out;
    out.open(p.db[2], DBTYPE_PREFILTER_RES, p.threads);
    const size_t nq = Q.keys.size();
    double kernelMs = 0;
    uint64_t nHits = 0;
    const bool covFilter = p.covThr > 0.0 && (p.covMode == 0 || p.covMode == 2 || p.covMode == 5);
    for (size_t b0 = 0, b1 = 0; b0 < nq; b0 = b1) {
        b1 = std::min(nq, b0 + (size_t)p.batch);
        std::vector<uint64_t> off(b1 - b0 + 1);
        for (size_t i = b0; i <= b1; i++) off[i - b0] = Q.off[i] - Q.off[b0];
        spb_prefilter_result *r = nullptr;
        check(spb_prefilter(E.ctx, Q.res.data() + Q.off[b0], off.data(), (uint32_t)(b1 - b0), &pp, &r), "spb_prefilter");
        kernelMs += r->kernel_ms;
        nHits += r->n_hits;
#pragma omp parallel num_threads(p.threads)
        {
            int th = 0;
#ifdef _OPENMP
            th = omp_get_thread_num();
#endif
            std::string buf;
            char tmp[64];
            // contiguous blocks per thread keep every thread's part of the data file in query order
#pragma omp for schedule(static)
            for (long i = 0; i < (long)(b1 - b0); i++) {
                buf.clear();
                const float qLen = (float)(Q.off[b0 + (size_t)i + 1] - Q.off[b0 + (size_t)i]);
                for (uint64_t h = r->offsets[i]; h < r->offsets[i + 1]; h++) {
                    const spb_hit &x = r->hits[h];   // id -> DB key (Prefiltering.cpp:827-829)
                    // -c with --cov-mode 0 / 2 / 5: hits whose lengths cannot reach the coverage are not written (after the
                    // max-seqs cut, Prefiltering.cpp:834-842)
                    if (covFilter && !can_be_covered((float)p.covThr, p.covMode, qLen, (float)(T.off[x.id + 1] - T.off[x.id]))) continue;
                    char *w = put_uint(tmp, T.keys[x.id]); *w++ = '\t';
                    w = put_int(w, x.score); *w++ = '\t';
                    w = put_int(w, x.diag); *w++ = '\n';
                    buf.append(tmp, (size_t)(w - tmp));
                }
                out.write(th, Q.keys[b0 + (size_t)i], buf.data(), buf.size());
            }
        }
        spb_free_prefilter_result(r);
    }
    out.close();
    log_line(p, "%llu prefilter hits for %zu queries; match kernels %.3f ms", (unsigned long long)nHits, nq, kernelMs);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// align
// ---------------------------------------------------------------------------------------------------------
int sw_mode_of(const Params &p) {
    if (p.addBacktrace || p.alignmentMode == 3) return 0;
    if (p.alignmentMode == 2) return 1;
    if (p.alignmentMode == 0 && (float)p.covThr > 0.0f) return ((float)p.seqIdThr == 0.0f) ? 1 : 0;
    return 2;
}

void fill_align_params(const Params &p, const Engine &E, uint64_t dbRes, spb_align_params *ap) {
    const int go = atoi(multiparam_aa(p.gapOpen).c_str()), ge = atoi(multiparam_aa(p.gapExtend).c_str());
    memset(ap, 0, sizeof(*ap));
    ap->gap_open = go;
    ap->gap_extend = ge;
    ap->eval_thr = p.evalue;
    ap->db_res_count = dbRes;
    ap->comp_bias = p.compBias;
    ap->byte_lanes = p.byteLanes;
    ap->word_lanes = p.wordLanes;
    ap->cov_thr = (float)p.covThr;          // Parameters holds -c and --min-seq-id as float
    ap->cov_mode = p.covMode;
    ap->seq_id_thr = (float)p.seqIdThr;
    ap->aln_len_thr = p.alnLenThr;
    ap->seq_id_mode = p.seqIdMode;
    // Alignment::Alignment + initSWMode (Alignment.cpp:29-37,174-196): -a forces score + coverage + identity with backtrace;
    // --alignment-mode 0 picks by the thresholds, 1 = score only, 2 = score + coverage, 3 = the full mode without the backtrace column
    ap->sw_mode = sw_mode_of(p);
    ap->max_accept = p.maxAccept;        // INT_MAX = no limit
    ap->max_rejected = p.maxRejected;
    if (!gumbel_for(E.align.name, go, ge, p.gumbel, &ap->gumbel))
        throw Fail("no Gumbel parameters for " + E.align.name + " gap open " + std::to_string(go) + " extend " + std::to_string(ge) +
                   ": pass the twelve ALP values of the reference run with --gumbel lambda,K,a_I,b_I,alpha_I,beta_I,a_J,b_J,alpha_J,beta_J,sigma,tau");
}

void check_align_params(const Params &p) {
    check_unimplemented_flags(p);
    if (p.realign != 0 && p.realign != 1) throw Fail("Error in argument --realign");
    if (p.realign && !p.realignOk) throw Fail("--realign is only served by the align module (not supported in the fused search)");
    if (p.realign && (p.altAli > 0 || p.maxRejected != INT_MAX || p.maxAccept != INT_MAX))
        throw Fail("--realign together with --alt-ali / --max-accept / --max-rejected is not supported");
    if (p.realign && p.realignMaxSeqs < 1) throw Fail("Error in argument --realign-max-seqs");
    if (p.altAli < 0) throw Fail("Error in argument --alt-ali");
    if (p.altAli > 0 && !p.altAliOk) throw Fail("--alt-ali is only served by the align module (not supported in the fused search)");
    if (p.altAli > 0 && sw_mode_of(p) == 2)
        throw Fail("--alt-ali needs start positions: not supported with the score-only alignment mode");
    if (p.seqIdThr < 0 || p.seqIdThr > 1) throw Fail("Error in argument --min-seq-id");
    if (p.covThr < 0 || p.covThr > 1) throw Fail("Error in argument -c");
    if (p.covMode < 0 || p.covMode > 5) throw Fail("Error in argument --cov-mode");
    if (p.seqIdMode < 0 || p.seqIdMode > 2) throw Fail("Error in argument --seq-id-mode");
    if (p.alnLenThr < 0) throw Fail("Error in argument --min-aln-len");
    if (p.maxRejected < 1 || p.maxAccept < 1) throw Fail("Error in argument --max-rejected / --max-accept (must be >= 1)");
    if (p.exhaustive && (p.maxRejected != INT_MAX || p.maxAccept != INT_MAX))
        throw Fail("--max-rejected / --max-accept walk a prefilter list in its order: not supported with --exhaustive-search 1");
    // -a forces SCORE_COV_SEQID (Alignment.cpp:35-37); without it the rows have ten columns (Matcher::resultToBuffer,
    // Matcher.cpp:355-367) and --alignment-mode selects how much is computed (fill_align_params)
    if (p.addBacktrace != 0 && p.addBacktrace != 1) throw Fail("Error in argument -a");
    if (p.alignmentMode < 0 || p.alignmentMode > 3)
        throw Fail("--alignment-mode " + std::to_string(p.alignmentMode) + " (ungapped) is not supported on this path");
}

// rows of one batch as text, one entry per query (Matcher::resultToBuffer, M/alignment/Matcher.cpp:323-370); sink(thread,
// batch-local query index, payload) is called once per query from an OpenMP team
template <class Sink>
void format_alignment_rows(const Params &p, const spb_align_result *r, const SeqSet &Q, size_t b0, const SeqSet &T, Sink sink) {
#pragma omp parallel num_threads(p.threads)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        std::string buf, cig;
        char tmp[256];
#pragma omp for schedule(static)
        for (long i = 0; i < (long)r->n_queries; i++) {
            buf.clear();
            const int qLen = (int)(Q.off[b0 + i + 1] - Q.off[b0 + i]);
            for (uint64_t k = r->offsets[i]; k < r->offsets[i + 1]; k++) {
                const spb_aln &a = r->rows[k];
                char *w = tmp;
                w = put_uint(w, T.keys[a.target]); *w++ = '\t';
                w = put_int(w, a.bits); *w++ = '\t';
                w += seqid_to_buffer(a.seq_id, w);
                *w++ = '\t';
                w = put_e3(w, a.eval); *w++ = '\t';
                w = put_int(w, a.q_start); *w++ = '\t';
                w = put_int(w, a.q_end); *w++ = '\t';
                w = put_int(w, qLen); *w++ = '\t';
                w = put_int(w, a.t_start); *w++ = '\t';
                w = put_int(w, a.t_end); *w++ = '\t';
                w = put_int(w, (int)(T.off[a.target + 1] - T.off[a.target]));
                if (!p.addBacktrace) {      // ten columns (Matcher.cpp:355-367)
                    *w++ = '\n';
                    buf.append(tmp, (size_t)(w - tmp));
                    continue;
                }
                *w++ = '\t';
                buf.append(tmp, (size_t)(w - tmp));
                cig.clear();
                if (a.cigar_len) compress_cigar(r->cigar_ops + a.cigar_off, a.cigar_len, cig);
                buf += cig;
                buf.push_back('\n');
            }
            sink(th, (size_t)i, buf);
        }
    }
}

void write_alignment_rows(const Params &p, const spb_align_result *r, const SeqSet &Q, size_t b0, const SeqSet &T, DBWriter &out) {
    format_alignment_rows(p, r, Q, b0, T, [&](int th, size_t i, const std::string &buf) { out.write(th, Q.keys[b0 + i], buf.data(), buf.size()); });
}

// one result row as text (Matcher::resultToBuffer) -- the row formatter of format_alignment_rows for a single row
void format_one_row(const Params &p, const spb_aln &a, uint32_t tKey, int qLen, int tLen, const char *cigarOps, std::string &out) {
    char tmp[256];
    char *w = tmp;
    w = put_uint(w, tKey); *w++ = '\t';
    w = put_int(w, a.bits); *w++ = '\t';
    w += seqid_to_buffer(a.seq_id, w);
    *w++ = '\t';
    w = put_e3(w, a.eval); *w++ = '\t';
    w = put_int(w, a.q_start); *w++ = '\t';
    w = put_int(w, a.q_end); *w++ = '\t';
    w = put_int(w, qLen); *w++ = '\t';
    w = put_int(w, a.t_start); *w++ = '\t';
    w = put_int(w, a.t_end); *w++ = '\t';
    w = put_int(w, tLen);
    if (!p.addBacktrace) {
        *w++ = '\n';
        out.append(tmp, (size_t)(w - tmp));
        return;
    }
    *w++ = '\t';
    out.append(tmp, (size_t)(w - tmp));
    std::string cig;
    if (a.cigar_len) compress_cigar(cigarOps + a.cigar_off, a.cigar_len, cig);
    out += cig;
    out.push_back('\n');
}

// `align --alt-ali N` (Alignment::computeAlternativeAlignment, Alignment.cpp:576-611): for every result of the first pass the aligned
// target stretch [dbStart, dbEnd) is masked with X and the pair is aligned again, up to N times while the criteria hold; every
// further alignment masks its own stretch too.  Here the masked targets of one round form a temporary target set (one copy per
// live result) and go through the same spb_align call, N rounds per batch; nothing changes on the device.  The rows of a query
// are then ordered like the reference orders them: first-pass results in the order of the prefilter list, behind them the
// alternative alignments of result 0, 1, ... in the order they were found, and std::sort with Matcher::compareHits over that
// sequence (rows that tie in E-value, score, length and key keep the order the library's introsort gives them -- the reference
// binary is built with the same libstdc++).
struct AltRow { double eval; int bits; uint32_t tlen, tkey; std::string text; };

void align_batch_with_alternatives(Engine &E, const Params &p, const spb_align_params &ap, const SeqSet &PQ, size_t b0, size_t b1,
                                   const std::vector<uint64_t> &pairOff, const std::vector<uint32_t> &pairT, const SeqSet &T,
                                   const spb_align_result *first, DBWriter &out, uint64_t *nAltRows) {
    const size_t nq = b1 - b0;
    // live results: one masked copy of the target each
    struct Live { uint32_t q; uint32_t t; std::vector<uint8_t> res; size_t slot; };   // slot: index into rowsOf[q]
    std::vector<std::vector<AltRow>> firstRows(nq);            // first pass, list order
    std::vector<std::vector<std::vector<AltRow>>> altRows(nq); // per first-pass result: its alternatives in order
    std::vector<Live> live;
    for (size_t i = 0; i < nq; i++) {
        const uint64_t a = pairOff[b0 + i] - pairOff[b0], b = pairOff[b0 + i + 1] - pairOff[b0];
        const int qLen = (int)(PQ.off[b0 + i + 1] - PQ.off[b0 + i]);
        // the device returns the rows sorted; the reference starts from the list order: position of every row's target in the list
        std::vector<char> used((size_t)(b - a), 0);
        std::vector<std::pair<uint64_t, uint64_t>> byPos;     // (list position, row index)
        for (uint64_t k = first->offsets[i]; k < first->offsets[i + 1]; k++) {
            uint64_t pos = b - a;
            for (uint64_t j = 0; j < b - a; j++)
                if (!used[(size_t)j] && pairT[pairOff[b0] + a + j] == first->rows[k].target) { pos = j; break; }
            if (pos == b - a) throw Fail("align --alt-ali: a result row names a target that is not in the prefilter list");
            used[(size_t)pos] = 1;
            byPos.push_back(std::make_pair(pos, k));
        }
        std::sort(byPos.begin(), byPos.end());
        altRows[i].resize(byPos.size());
        for (size_t s0 = 0; s0 < byPos.size(); s0++) {
            const spb_aln &r = first->rows[byPos[s0].second];
            const uint32_t t = r.target;
            const int tLen = (int)(T.off[t + 1] - T.off[t]);
            AltRow row{r.eval, r.bits, (uint32_t)tLen, T.keys[t], std::string()};
            format_one_row(p, r, T.keys[t], qLen, tLen, first->cigar_ops, row.text);
            firstRows[i].push_back(std::move(row));
            Live l;
            l.q = (uint32_t)i; l.t = t; l.slot = s0;
            l.res.assign(T.res.begin() + (long)T.off[t], T.res.begin() + (long)T.off[t + 1]);
            for (int x = std::max(r.t_start, 0); x < r.t_end && x < tLen; x++) l.res[(size_t)x] = 20;      // [dbStart, dbEnd): X
            live.push_back(std::move(l));
        }
    }
    spb_align_params ap2 = ap;
    ap2.max_accept = 0;          // the limits of the list walk do not apply to the alternatives
    ap2.max_rejected = 0;
    for (int round = 0; round < p.altAli && !live.empty(); round++) {
        // temporary target set = the masked copies, in the order of `live` (grouped by query, ascending)
        std::vector<uint8_t> tres;
        std::vector<uint64_t> toff(1, 0);
        std::vector<uint32_t> tkeys;
        std::vector<uint64_t> po(nq + 1, 0);
        std::vector<uint32_t> pt;
        for (size_t k = 0; k < live.size(); k++) {
            tres.insert(tres.end(), live[k].res.begin(), live[k].res.end());
            toff.push_back(tres.size());
            tkeys.push_back(T.keys[live[k].t]);
            po[live[k].q + 1]++;
            pt.push_back((uint32_t)k);
        }
        tres.push_back(0);
        for (size_t i = 0; i < nq; i++) po[i + 1] += po[i];
        std::vector<uint64_t> qo(nq + 1);
        for (size_t i = 0; i <= nq; i++) qo[i] = PQ.off[b0 + i] - PQ.off[b0];
        check(spb_align_set_targets(E.ctx, tres.data(), toff.data(), (uint32_t)live.size(), tkeys.data()), "spb_align_set_targets");
        spb_align_result *r = nullptr;
        check(spb_align(E.ctx, PQ.res.data() + PQ.off[b0], qo.data(), (uint32_t)nq, po.data(), pt.data(), &ap2, &r), "spb_align");
        std::vector<char> passed(live.size(), 0);
        for (size_t i = 0; i < nq; i++) {
            const int qLen = (int)(PQ.off[b0 + i + 1] - PQ.off[b0 + i]);
            for (uint64_t k = r->offsets[i]; k < r->offsets[i + 1]; k++) {
                const spb_aln &a = r->rows[k];
                Live &l = live[a.target];
                const int tLen = (int)l.res.size();
                AltRow row{a.eval, a.bits, (uint32_t)tLen, T.keys[l.t], std::string()};
                format_one_row(p, a, T.keys[l.t], qLen, tLen, r->cigar_ops, row.text);
                altRows[i][l.slot].push_back(std::move(row));
                for (int x = std::max(a.t_start, 0); x < a.t_end && x < tLen; x++) l.res[(size_t)x] = 20;
                passed[a.target] = 1;
                (*nAltRows)++;
            }
        }
        spb_free_align_result(r);
        std::vector<Live> next;
        for (size_t k = 0; k < live.size(); k++)
            if (passed[k]) next.push_back(std::move(live[k]));
        live.swap(next);
    }
    // the original targets back for the next batch
    check(spb_align_set_targets(E.ctx, T.res.data(), T.off.data(), (uint32_t)T.keys.size(), T.keys.data()), "spb_align_set_targets");
#pragma omp parallel num_threads(p.threads)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
#pragma omp for schedule(static)
        for (long i = 0; i < (long)nq; i++) {
            std::vector<AltRow> all;
            for (AltRow &r : firstRows[(size_t)i]) all.push_back(std::move(r));
            for (std::vector<AltRow> &v : altRows[(size_t)i])
                for (AltRow &r : v) all.push_back(std::move(r));
            if (all.size() > 1)
                std::sort(all.begin(), all.end(), [](const AltRow &x, const AltRow &y) {      // Matcher::compareHits
                    if (x.eval != y.eval) return x.eval < y.eval;
                    if (x.bits != y.bits) return x.bits > y.bits;
                    if (x.tlen != y.tlen) return x.tlen < y.tlen;
                    return x.tkey < y.tkey;
                });
            std::string buf;
            for (const AltRow &r : all) buf += r.text;
            out.write(th, PQ.keys[b0 + (size_t)i], buf.data(), buf.size());
        }
    }
}

// `align --realign 1` (Alignment.cpp:47-56,415-447): the first pass only scores (SCORE_ONLY, coverage threshold off, the length
// pre-check of canBeCovered on) and is sorted like Matcher::compareHits; the second pass aligns the survivors in that order with the
// REALIGN matrix (score bias + --realign-score-bias; its own composition-bias profile), without E-value limit, keeps those whose
// coverage reaches -c (at most --realign-max-seqs of them) and reports their boundaries, identity and backtrace with the score and
// E-value of the first pass.  Both passes are the same spb_align call; the matrices are swapped in between.
void align_batch_realign(Engine &E, const Params &p, const spb_align_params &ap, const SubMat &realignMat, const SeqSet &PQ, size_t b0,
                         size_t b1, const std::vector<uint64_t> &pairOff, const std::vector<uint32_t> &pairT, const SeqSet &T,
                         DBWriter &out, uint64_t *nRows, uint64_t *nCells, double *kernelMs) {
    const size_t nq = b1 - b0;
    std::vector<uint64_t> qo(nq + 1), po(nq + 1, 0);
    std::vector<uint32_t> pt;
    for (size_t i = 0; i <= nq; i++) qo[i] = PQ.off[b0 + i] - PQ.off[b0];
    for (size_t i = 0; i < nq; i++) {      // Util::canBeCovered with the real -c before the first pass (Alignment.cpp:378)
        const float qLen = (float)(qo[i + 1] - qo[i]);
        for (uint64_t k = pairOff[b0 + i]; k < pairOff[b0 + i + 1]; k++) {
            const uint32_t t = pairT[k];
            if (can_be_covered((float)p.covThr, p.covMode, qLen, (float)(T.off[t + 1] - T.off[t]))) pt.push_back(t);
        }
        po[i + 1] = pt.size();
    }
    spb_align_params ap1 = ap;
    ap1.sw_mode = 2;
    ap1.cov_thr = 0.0f;
    ap1.cov_mode = 0;
    spb_align_result *r1 = nullptr;
    check(spb_align(E.ctx, PQ.res.data() + PQ.off[b0], qo.data(), (uint32_t)nq, po.data(), pt.data(), &ap1, &r1), "spb_align");
    *nCells += r1->n_cells;
    *kernelMs += r1->kernel_ms;
    // second pass: the rows of the first, in their order
    std::vector<uint64_t> po2(nq + 1, 0);
    std::vector<uint32_t> pt2((size_t)r1->n_rows);
    for (size_t i = 0; i < nq; i++) {
        for (uint64_t k = r1->offsets[i]; k < r1->offsets[i + 1]; k++) pt2[(size_t)k] = r1->rows[k].target;
        po2[i + 1] = r1->offsets[i + 1];
    }
    spb_align_params ap2 = ap;
    ap2.sw_mode = 0;
    ap2.eval_thr = (double)FLT_MAX;
    ap2.seq_id_thr = 0.0f;
    ap2.aln_len_thr = 0;
    ap2.max_accept = p.realignMaxSeqs;      // (INT_MAX: no limit) at most that many results pass, walked in the first pass's order
    ap2.max_rejected = 0;
    const bool swap = p.realignScoreBias != 0.0;
    if (swap) check(spb_set_matrices(E.ctx, E.seed.sub, E.seed.pback, E.ungapped.sub, realignMat.sub, realignMat.pback), "spb_set_matrices");
    spb_align_result *r2 = nullptr;
    check(spb_align(E.ctx, PQ.res.data() + PQ.off[b0], qo.data(), (uint32_t)nq, po2.data(), pt2.data(), &ap2, &r2), "spb_align");
    if (swap) check(spb_set_matrices(E.ctx, E.seed.sub, E.seed.pback, E.ungapped.sub, E.align.sub, E.align.pback), "spb_set_matrices");
    *nCells += r2->n_cells;
    *kernelMs += r2->kernel_ms;
    uint64_t rows = 0;
#pragma omp parallel num_threads(p.threads) reduction(+ : rows)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        std::string buf;
#pragma omp for schedule(static)
        for (long i = 0; i < (long)nq; i++) {
            buf.clear();
            const int qLen = (int)(qo[(size_t)i + 1] - qo[(size_t)i]);
            std::vector<char> used((size_t)(r2->offsets[i + 1] - r2->offsets[i]), 0);
            for (uint64_t k = r1->offsets[i]; k < r1->offsets[i + 1]; k++) {      // first-pass order = the order of the result
                const spb_aln &f = r1->rows[k];
                for (uint64_t j = r2->offsets[i]; j < r2->offsets[i + 1]; j++) {
                    if (used[(size_t)(j - r2->offsets[i])] || r2->rows[j].target != f.target) continue;
                    used[(size_t)(j - r2->offsets[i])] = 1;
                    spb_aln a = r2->rows[j];
                    a.bits = f.bits;             // res.score / res.eval of the first pass (Alignment.cpp:433-434)
                    a.eval = f.eval;
                    format_one_row(p, a, T.keys[a.target], qLen, (int)(T.off[a.target + 1] - T.off[a.target]), r2->cigar_ops, buf);
                    rows++;
                    break;
                }
            }
            out.write(th, PQ.keys[b0 + (size_t)i], buf.data(), buf.size());
        }
    }
    *nRows += rows;
    spb_free_align_result(r1);
    spb_free_align_result(r2);
}

int run_align(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 4, p);
    p.altAliOk = true;
    p.realignOk = true;
    check_align_params(p);
    Engine E;
    E.open(p);
    DBReader qdb, tdb, pdb;
    qdb.open(p.db[0]);
    tdb.open(p.db[1]);
    pdb.open(p.db[2]);
    SeqSet Q, T;
    // align reads the raw (unmasked) sequences; ids = key order here, rows carry DB keys anyway
    load_sequences(qdb, E.align, Q);
    load_sequences(tdb, E.align, T);
    check_seq_len(p, Q.off, "query");
    check_seq_len(p, T.off, "target");
    spb_align_params ap;
    fill_align_params(p, E, tdb.aminoAcidDBSize(), &ap);   // Alignment.cpp:269
    check(spb_align_set_targets(E.ctx, T.res.data(), T.off.data(), (uint32_t)T.keys.size(), T.keys.data()), "spb_align_set_targets");

    // work list: one entry of the prefilter DB per query key, first column = target key (Alignment.cpp:348-392)
    const size_t np = pdb.size();
    SeqSet PQ;   // queries in prefilter-DB order
    std::vector<uint64_t> pairOff(1, 0);
    std::vector<uint32_t> pairT;
    PQ.off.assign(1, 0);
    for (size_t i = 0; i < np; i++) {
        const DBEntry &e = pdb.byKey(i);
        const size_t qid = qdb.linearIdOfKey(e.key);
        if (qid == SIZE_MAX) throw Fail("Invalid query sequence " + std::to_string(e.key));
        PQ.keys.push_back(e.key);
        PQ.res.insert(PQ.res.end(), Q.res.begin() + (long)Q.off[qid], Q.res.begin() + (long)Q.off[qid + 1]);
        PQ.off.push_back(PQ.res.size());
        const char *d = pdb.data(e), *end = d + (e.length ? e.length - 1 : 0);
        while (d < end) {
            char *next = nullptr;
            const unsigned long key = strtoul(d, &next, 10);
            if (next == d) break;
            const size_t tid = tdb.linearIdOfKey((uint32_t)key);
            if (tid == SIZE_MAX) throw Fail("Invalid target sequence " + std::to_string(key) + " in the prefilter list of query " + std::to_string(e.key));
            pairT.push_back((uint32_t)tid);
            const char *nl = (const char *)memchr(next, '\n', (size_t)(end - next));
            d = nl ? nl + 1 : end;
        }
        pairOff.push_back(pairT.size());
    }
    PQ.res.push_back(0);
    DBWriter out;
    out.open(p.db[3], DBTYPE_ALIGNMENT_RES, p.threads);
    uint64_t nRows = 0, nCells = 0, nAlt = 0;
    double kernelMs = 0;
    SubMat realignMat;
    if (p.realign) {
        // -a is switched on for a realignment (Alignment.cpp:52-55); the matrix of the second pass (Alignment.cpp:164-170)
        p.addBacktrace = 1;
        realignMat.load(p.subMat, 2.0, p.scoreBias + p.realignScoreBias);
    }
    for (size_t b0 = 0, b1 = 0; b0 < np; b0 = b1) {
        b1 = std::min(np, b0 + (size_t)p.batch);
        while (b1 > b0 + 1 && pairOff[b1] - pairOff[b0] > 0x7FFFFFF0ull) b1 = b0 + (b1 - b0) / 2;
        std::vector<uint64_t> qo(b1 - b0 + 1), po(b1 - b0 + 1);
        for (size_t i = b0; i <= b1; i++) { qo[i - b0] = PQ.off[i] - PQ.off[b0]; po[i - b0] = pairOff[i] - pairOff[b0]; }
        if (p.realign) {
            align_batch_realign(E, p, ap, realignMat, PQ, b0, b1, pairOff, pairT, T, out, &nRows, &nCells, &kernelMs);
            continue;
        }
        spb_align_result *r = nullptr;
        check(spb_align(E.ctx, PQ.res.data() + PQ.off[b0], qo.data(), (uint32_t)(b1 - b0), po.data(), pairT.data() + pairOff[b0], &ap, &r),
              "spb_align");
        nRows += r->n_rows;
        nCells += r->n_cells;
        kernelMs += r->kernel_ms;
        if (p.altAli > 0) align_batch_with_alternatives(E, p, ap, PQ, b0, b1, pairOff, pairT, T, r, out, &nAlt);
        else write_alignment_rows(p, r, PQ, b0, T, out);
        spb_free_align_result(r);
    }
    out.close();
    log_line(p, "%llu alignments accepted of %zu pairs; %.2f GCUPS in the forward kernels", (unsigned long long)nRows, pairT.size(),
             kernelMs > 0 ? (double)nCells / kernelMs / 1e6 : 0.0);
    if (p.altAli > 0) log_line(p, "%llu alternative alignments (--alt-ali %d)", (unsigned long long)nAlt, p.altAli);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// nucleotide re-alignment branch (proteinaln2nucl on the device, combineprotnuclaln in setscore.cpp)
// ---------------------------------------------------------------------------------------------------------
void set_predictmatch_defaults(struct Params &p);

struct NuclSet {
    std::vector<uint8_t> nt;          // raw ASCII nucleotides, entries back to back
    std::vector<uint64_t> off;
    std::vector<uint8_t> startCodon;  // the amino-acid entry begins with '*' (proteinaln2nucl.cpp:92-96,113-118)
    uint64_t dbSize = 0;              // getAminoAcidDBSize of the nucleotide DB
};

// entry i = the <db>_nucl_orf entry with key keys[i] (the reference looks both DBs up by key, proteinaln2nucl.cpp:88-112)
void load_nucl(const std::string &nuclName, const DBReader &aaDb, const std::vector<uint32_t> &keys, NuclSet &N) {
    DBReader ndb;
    ndb.open(nuclName);
    if (ndb.dbtype() != 1 || aaDb.dbtype() != DBTYPE_AMINO_ACIDS) throw Fail("Wrong query and target database input");
    N.dbSize = ndb.aminoAcidDBSize();
    const size_t n = keys.size();
    N.off.assign(n + 1, 0);
    N.startCodon.assign(n + 1, 0);
    std::vector<size_t> pos(n);
    for (size_t i = 0; i < n; i++) {
        pos[i] = ndb.findKey(keys[i]);
        if (pos[i] == SIZE_MAX) throw Fail("Invalid database key " + std::to_string(keys[i]) + " in " + nuclName);
        const DBEntry &e = ndb.byKey(pos[i]);
        N.off[i + 1] = N.off[i] + (e.length >= 2 ? e.length - 2 : 0);     // DBReader::getSeqLen
        const size_t a = aaDb.findKey(keys[i]);
        if (a != SIZE_MAX && aaDb.byKey(a).length > 0 && aaDb.data(aaDb.byKey(a))[0] == '*') N.startCodon[i] = 1;
    }
    N.nt.resize(N.off[n] + 1);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) memcpy(N.nt.data() + N.off[i], ndb.data(ndb.byKey(pos[(size_t)i])), (size_t)(N.off[i + 1] - N.off[i]));
}

void fill_nucl_params(const Params &p, uint64_t dbSize, spb_nucl_params *np) {
    memset(np, 0, sizeof(*np));
    std::string mat = multiparam_nucl(p.subMat);
    const size_t slash = mat.find_last_of('/');
    if (slash != std::string::npos) mat = mat.substr(slash + 1);
    if (mat != "nucleotide.out") throw Fail("nucleotide re-alignment: only nucleotide.out is supported as nucleotide matrix, got '" + mat + "'");
    np->gap_open = atoi(multiparam_nucl(p.gapOpen).c_str());
    np->gap_extend = atoi(multiparam_nucl(p.gapExtend).c_str());
    np->db_res_count = dbSize;
    if (!gumbel_for(mat, np->gap_open, np->gap_extend, p.gumbelNucl, &np->gumbel))
        throw Fail("no Gumbel parameters for " + mat + " gap open " + std::to_string(np->gap_open) + " extend " + std::to_string(np->gap_extend) +
                   ": pass the twelve ALP values of the reference run with --gumbel-nucl");
    // NucleotideMatrix::setupLetterMapping (M/commons/NucleotideMatrix.cpp:18-63), alphabet order of nucleotide.out: A C T G X
    for (int c = 0; c < 256; c++) {
        int n = 4;
        switch (c < 128 ? toupper(c) : 0) {
            case 'A': n = 0; break;
            case 'C': case 'M': case 'Y': case 'H': n = 1; break;
            case 'T': case 'U': case 'W': n = 2; break;
            case 'G': case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': n = 3; break;
            default: break;
        }
        np->letter[c] = (uint8_t)n;
    }
    // nucleotide.out at bit factor 1, bias 0 (lib/mmseqs/data/nucleotide.out): match 2, everything else -3
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) np->sub[i * 5 + j] = (int16_t)((i == j && i < 4) ? 2 : -3);
}

void nucl_row_error(const spb_nucl_aln &n) {
    if (n.err == 1) throw Fail("Alignment contains unalignable character");
    if (n.err) throw Fail("nucleotide re-alignment: an alignment row does not fit its nucleotide entry (code " + std::to_string(n.err) + ")");
}

// the row proteinaln2nucl writes (Matcher::resultToBuffer, proteinaln2nucl.cpp:160-172): nucleotide coordinates / lengths and
// the protein backtrace with every run length tripled
void append_nucl_row(std::string &buf, uint32_t tKey, const spb_nucl_aln &n, const char *ops, uint32_t nOps) {
    char tmp[256];
    char *w = tmp;
    w += sprintf(w, "%u\t%d\t", tKey, n.bits);
    w += seqid_to_buffer(n.seq_id, w);
    w += sprintf(w, "\t%.3E\t%d\t%d\t%u\t%d\t%d\t%u\t", n.eval, n.q_start, n.q_end, n.q_len, n.t_start, n.t_end, n.t_len);
    buf.append(tmp, (size_t)(w - tmp));
    for (uint32_t i = 0; i < nOps;) {
        uint32_t j = i;
        while (j < nOps && ops[j] == ops[i]) j++;
        buf.append(tmp, (size_t)sprintf(tmp, "%u%c", (j - i) * 3, ops[i]));
        i = j;
    }
    buf.push_back('\n');
}

// <queryDB>_nucl_orf <targetDB>_nucl_orf <queryDB> <targetDB> <alnDB> <outDB>   (proteinaln2nucl.cpp:13-72)
int run_proteinaln2nucl(int argc, const char **argv) {
    Params p;
    set_predictmatch_defaults(p);
    parse_args(argc, argv, 2, 6, p);
    DBReader qaa, taa, aln;
    qaa.open(p.db[2]);
    taa.open(p.db[3]);
    aln.open(p.db[4]);
    std::vector<uint32_t> qkeys(aln.size()), tkeys(taa.size());
    for (size_t i = 0; i < aln.size(); i++) qkeys[i] = aln.byKey(i).key;
    for (size_t i = 0; i < taa.size(); i++) tkeys[i] = taa.byKey(i).key;
    NuclSet QN, TN;
    load_nucl(p.db[0], qaa, qkeys, QN);
    load_nucl(p.db[1], taa, tkeys, TN);
    spb_nucl_params np;
    fill_nucl_params(p, TN.dbSize, &np);
    // Matcher::parseAlignmentRecord (M/alignment/Matcher.cpp:232-321): key score seqId eval qStart qEnd qLen dbStart dbEnd dbLen backtrace
    std::vector<spb_aln> rows;
    std::vector<uint64_t> rowOff(aln.size() + 1, 0);
    std::string ops;
    for (size_t i = 0; i < aln.size(); i++) {
        const DBEntry &e = aln.byKey(i);
        const char *d = aln.data(e), *end = d + strnlen(d, e.length ? (size_t)e.length - 1 : 0);
        while (d < end) {
            const char *nl = (const char *)memchr(d, '\n', (size_t)(end - d));
            const char *le = nl ? nl : end;
            std::vector<std::string> f;
            for (const char *c = d; c < le;) {
                const char *t = (const char *)memchr(c, '\t', (size_t)(le - c));
                f.emplace_back(c, (size_t)((t ? t : le) - c));
                c = t ? t + 1 : le;
            }
            if (f.size() < 11 || f[10].empty()) throw Fail("This module only supports database input with backtrace string");
            spb_aln a;
            memset(&a, 0, sizeof(a));
            a.query = (uint32_t)i;
            const uint32_t tKey = (uint32_t)strtoul(f[0].c_str(), nullptr, 10);
            const size_t tid = taa.findKey(tKey);
            if (tid == SIZE_MAX) throw Fail("Invalid target database key " + f[0]);
            a.target = (uint32_t)tid;
            a.q_start = atoi(f[4].c_str()); a.q_end = atoi(f[5].c_str());
            a.t_start = atoi(f[7].c_str()); a.t_end = atoi(f[8].c_str());
            a.cigar_off = (uint32_t)ops.size();
            for (const char *c = f[10].c_str(); *c;) {
                unsigned cnt = 0;
                while (isdigit((unsigned char)*c)) cnt = cnt * 10 + (unsigned)(*c++ - '0');
                if (!*c) break;
                if (*c == 'M' || *c == 'D' || *c == 'I') ops.append(cnt, *c);
                c++;
            }
            a.cigar_len = (uint32_t)ops.size() - a.cigar_off;
            rows.push_back(a);
            d = nl ? nl + 1 : end;
        }
        rowOff[i + 1] = rows.size();
    }
    Engine E;
    check(spb_ctx_create(p.device, &E.ctx), "spb_ctx_create");
    check(spb_nucl_set_targets(E.ctx, TN.nt.data(), TN.off.data(), (uint32_t)tkeys.size(), TN.startCodon.data()), "spb_nucl_set_targets");
    check(spb_nucl_set_queries(E.ctx, QN.nt.data(), QN.off.data(), (uint32_t)qkeys.size(), QN.startCodon.data()), "spb_nucl_set_queries");
    spb_nucl_aln *nr = nullptr;
    check(spb_nucl_rescore(E.ctx, &np, rows.data(), rows.size(), ops.data(), ops.size(), &nr), "spb_nucl_rescore");
    DBWriter out;
    out.open(p.db[5], DBTYPE_ALIGNMENT_RES, p.threads);
    std::string buf;
    for (size_t i = 0; i < aln.size(); i++) {
        buf.clear();
        for (uint64_t k = rowOff[i]; k < rowOff[i + 1]; k++) {
            nucl_row_error(nr[k]);
            append_nucl_row(buf, tkeys[rows[k].target], nr[k], ops.data() + rows[k].cigar_off, rows[k].cigar_len);
        }
        out.write((int)(i * (size_t)p.threads / (aln.size() ? aln.size() : 1)), qkeys[i], buf.data(), buf.size());
    }
    out.close();
    spb_free_nucl_result(E.ctx, nr);
    return 0;
}

// <protAlnDB> <nuclAlnDB> <outDB>   (CombineProtNuclAln.cpp:13-25)
int run_combineprotnuclaln(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 3, p);
    TextDB prot, nucl, out;
    prot.load(p.db[0]);
    nucl.load(p.db[1]);
    combine_prot_nucl_aln(prot, nucl, p.threads, out);
    out.save(p.db[2], p.threads);
    return 0;
}

// both steps, hit lists stay in HBM (what blastp.sh does with a pref DB in between); sink(thread, query key, rows text)
// nucl: the nucleotide re-alignment branch (predictmatch.sh:27-45) fused in -- after the alignment stage the resident rows are
// re-scored on the resident <db>_nucl_orf entries (spb_nucl_rescore_resident) and combined like combineprotnuclaln; the sink
// then receives the combined rows, and extra(thread, key, protein rows, nucleotide rows), if given, the two intermediates
typedef std::function<void(int, uint32_t, const std::string &, const std::string &)> NuclExtraSink;
// batchDone(first sink thread, threads): called by a GPU's formatter thread after the rows of a batch went through the sink
typedef std::function<void(int, int)> BatchDone;
template <class Sink>
void search_dbs(const Params &p, const std::string &qName, const std::string &tName, Sink sink, bool nucl = false,
                const NuclExtraSink *extra = nullptr, const BatchDone *batchDone = nullptr) {
    // --gpus G: one context + one host thread per device.  Every device gets the whole target set for alignment (53 MB at
    // GenBank-phage size) but only its k-mer range of the index table (spb_index_shard); the ranges of the other devices are
    // read through NVLink peer memory by the lookup kernel.  Query batches are handed out to the devices one by one; rows
    // are the unsplit rows by construction (Prefiltering.cpp:258-362,559-643 / Alignment.cpp:226-249 are the reference's split).
    const int G = std::max(1, p.gpus);
    std::vector<std::unique_ptr<Engine>> eng((size_t)G);
    for (int g = 0; g < G; g++) {
        eng[(size_t)g].reset(new Engine());
        eng[(size_t)g]->load_matrices(p);
    }
    // The CUDA driver and the context of every device come up in the background (0.6 - 1 s) while this thread reads, encodes and
    // masks the DBs -- none of that needs the device.
    std::exception_ptr ctxErr;
    std::thread ctxThread([&]() {
        try {
            const int nDev = spb_device_count();
            if (G > 1 && p.device + G > nDev && !p.oversubscribe)
                throw Fail("--gpus " + std::to_string(G) + " from device " + std::to_string(p.device) + ": only " + std::to_string(nDev) +
                           " CUDA device(s) visible (--oversubscribe 1 places several shard contexts on one device)");
            std::vector<std::thread> th;
            std::vector<std::exception_ptr> err((size_t)G);
            for (int g = 0; g < G; g++) {
                Params pg = p;
                pg.device = (p.device + g < nDev || nDev == 0) ? p.device + g : p.device + g % std::max(1, nDev - p.device);
                th.emplace_back([&, g, pg]() { try { eng[(size_t)g]->create_context(pg); } catch (...) { err[(size_t)g] = std::current_exception(); } });
            }
            for (std::thread &x : th) x.join();
            for (std::exception_ptr &e : err) if (e) std::rethrow_exception(e);
        } catch (...) {
            ctxErr = std::current_exception();
        }
    });
    struct JoinGuardCtx {       // an exception while the DBs are read must not leave the context thread running
        std::thread &t;
        ~JoinGuardCtx() { if (t.joinable()) t.join(); }
    } joinCtx{ctxThread};
    Engine &E0 = *eng[0];
    PhaseClock &pc = g_clock;
    pc.mark("start-up");
    DBReader qdb, tdb;
    qdb.open(qName);
    tdb.open(tName);
    pc.mark("open DBs");
    SeqSet Q, T;
    load_sequences(qdb, E0.seed, Q);
    load_sequences(tdb, E0.seed, T, p.maskLower == 1);
    pc.mark("encode sequences");
    check_seq_len(p, Q.off, "query");
    check_seq_len(p, T.off, "target");
    NuclSet QN, TN;
    spb_nucl_params np;
    if (nucl) {
        load_nucl(qName + "_nucl_orf", qdb, Q.keys, QN);
        load_nucl(tName + "_nucl_orf", tdb, T.keys, TN);
        fill_nucl_params(p, TN.dbSize, &np);
    }
    const int kmerThr = kmer_threshold(p.sensitivity, p.kmerScore, 6);
    std::vector<uint8_t> masked;
    uint64_t nMasked = 0;
    bool cached = false;
    if (!p.exhaustive) cached = masked_targets(E0, p, T, tName, tdb.dataSize(), masked, &nMasked);
    pc.mark("tantan");
    ctxThread.join();
    if (ctxErr) std::rethrow_exception(ctxErr);
    pc.mark("(contexts ready)");
    std::vector<std::exception_ptr> errs((size_t)G);
    auto run_on_all = [&](const std::function<void(int)> &f) {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g]() { try { f(g); } catch (...) { errs[(size_t)g] = std::current_exception(); } });
        for (std::thread &t : th) t.join();
        for (std::exception_ptr &e : errs)
            if (e) { std::exception_ptr x = e; e = nullptr; std::rethrow_exception(x); }
    };
    run_on_all([&](int g) {
        Engine &E = *eng[(size_t)g];
        if (nucl) check(spb_nucl_set_targets(E.ctx, TN.nt.data(), TN.off.data(), (uint32_t)T.keys.size(), TN.startCodon.data()), "spb_nucl_set_targets");
        if (!p.exhaustive) {
            Params pq = p;
            if (g) pq.verbosity = 0;          // one "Index table" line
            build_index_from_masked(E, pq, T, masked, kmerThr, cached, nMasked, tName);
            if (G > 1) check(spb_index_shard(E.ctx, g, G), "spb_index_shard");
        }
        check(spb_align_set_targets(E.ctx, T.res.data(), T.off.data(), (uint32_t)T.keys.size(), T.keys.data()), "spb_align_set_targets");
    });
    masked.clear();
    masked.shrink_to_fit();
    pc.mark("index + targets");
    if (G > 1 && !p.exhaustive) {
        for (int g = 0; g < G; g++)
            for (int h = 0; h < G; h++)
                if (h != g) check(spb_index_shard_attach_ctx(eng[(size_t)g]->ctx, h, eng[(size_t)h]->ctx), "spb_index_shard_attach_ctx");
        log_line(p, "Index table: k-mer ranges sharded over %d GPUs (devices %d..%d), peer loads over NVLink", G, p.device, p.device + G - 1);
    }
    // Lanes: further contexts on every device that borrow its index / targets (spb_ctx_create_lane) and are driven by their own
    // host thread, so that upload, host-side sizing, kernels and read-back of different batches overlap on the device.
    const int L = std::max(1, p.lanes), W = G * L;
    struct Lanes {
        std::vector<spb_ctx *> ctx;
        ~Lanes() { for (spb_ctx *c : ctx) if (c) spb_ctx_destroy(c); }
    } lanes;
    std::vector<spb_ctx *> worker((size_t)W, nullptr);      // worker w = g * L + lane
    for (int g = 0; g < G; g++) {
        worker[(size_t)g * L] = eng[(size_t)g]->ctx;
        for (int l = 1; l < L; l++) {
            spb_ctx *c = nullptr;
            check(spb_ctx_create_lane(eng[(size_t)g]->ctx, &c), "spb_ctx_create_lane");
            lanes.ctx.push_back(c);
            worker[(size_t)g * L + l] = c;
        }
    }
    spb_prefilter_params pp;
    pp.kmer_thr = kmerThr;
    pp.max_seqs = p.maxSeqs;
    pp.min_diag_score = p.minUngapped;
    pp.bin_size = p.binSize ? p.binSize : bin_size_for(T.keys.size());
    pp.comp_bias = p.compBias;
    pp.id_base = 0;
    pp.emit_trunc_keys = 0;
    pp.exact_kmer = p.exactKmer;
    spb_align_params ap;
    fill_align_params(p, E0, tdb.aminoAcidDBSize(), &ap);
    const size_t nq = Q.keys.size();
    // frames per GPU pass; with several workers every one of them should see a few batches
    const size_t batch = W > 1 ? std::min<size_t>((size_t)p.batch, std::max<size_t>(4096, (nq + 2 * (size_t)W - 1) / (2 * (size_t)W))) : (size_t)p.batch;
    const size_t nBatches = (nq + batch - 1) / batch;
    std::atomic<size_t> nextBatch(0);
    Params pf = p;                         // formatter teams share the host threads
    pf.threads = std::max(1, p.threads / W);
    std::vector<std::exception_ptr> werrs((size_t)W);
    auto run_on_workers = [&](const std::function<void(int)> &f) {
        std::vector<std::thread> th;
        for (int w = 0; w < W; w++)
            th.emplace_back([&, w]() { try { f(w); } catch (...) { werrs[(size_t)w] = std::current_exception(); } });
        for (std::thread &t : th) t.join();
        for (std::exception_ptr &e : werrs)
            if (e) { std::exception_ptr x = e; e = nullptr; std::rethrow_exception(x); }
    };
    run_on_workers([&](int w) {
        struct { spb_ctx *ctx; } E{worker[(size_t)w]};
        const int thBase = w * pf.threads;
        std::thread formatter;
        std::exception_ptr fmtError;
        spb_align_result *prevR = nullptr;
        spb_nucl_aln *prevNr = nullptr;
        auto finish_previous = [&]() {
            if (formatter.joinable()) formatter.join();
            if (prevNr) spb_free_nucl_result(E.ctx, prevNr);
            if (prevR) spb_free_align_result(prevR);
            prevNr = nullptr;
            prevR = nullptr;
            if (fmtError) {
                std::exception_ptr e = fmtError;
                fmtError = nullptr;
                std::rethrow_exception(e);
            }
        };
        struct JoinGuard {       // an exception on the GPU path must not leave the formatter running on freed state
            std::thread &t;
            ~JoinGuard() { if (t.joinable()) t.join(); }
        } guard{formatter};
        for (;;) {
            const size_t bi = nextBatch.fetch_add(1);
            if (bi >= nBatches) break;
            const size_t b0 = bi * batch, b1 = std::min(nq, b0 + batch);
            std::vector<uint64_t> off(b1 - b0 + 1);
            for (size_t i = b0; i <= b1; i++) off[i - b0] = Q.off[i] - Q.off[b0];
            check(spb_queries_set(E.ctx, Q.res.data() + Q.off[b0], off.data(), (uint32_t)(b1 - b0)), "spb_queries_set");
            if (p.exhaustive) {
                spb_exhaustive_stats es;
                check(spb_pairs_exhaustive(E.ctx, &ap, &es), "spb_pairs_exhaustive");
                log_line(p, "Exhaustive search: %llu pairs scored, %llu kept, %.2f GCUPS in the streaming kernel", (unsigned long long)es.n_scored,
                         (unsigned long long)es.n_pairs, es.kernel_ms > 0 ? (double)es.n_cells / es.kernel_ms / 1e6 : 0.0);
            } else {
                check(spb_prefilter_resident(E.ctx, &pp), "spb_prefilter_resident");
                check(spb_pairs_from_prefilter(E.ctx), "spb_pairs_from_prefilter");
            }
            check(spb_align_resident(E.ctx, &ap), "spb_align_resident");
            spb_nucl_aln *nr = nullptr;
            if (nucl) {
                std::vector<uint64_t> noff(b1 - b0 + 1);
                for (size_t i = b0; i <= b1; i++) noff[i - b0] = QN.off[i] - QN.off[b0];
                check(spb_nucl_set_queries(E.ctx, QN.nt.data() + QN.off[b0], noff.data(), (uint32_t)(b1 - b0), QN.startCodon.data() + b0),
                      "spb_nucl_set_queries");
                uint64_t nRows = 0;
                double ms = 0;
                check(spb_nucl_rescore_resident(E.ctx, &np, &nr, &nRows, &ms), "spb_nucl_rescore_resident");
                log_line(p, "Nucleotide re-alignment: %llu rows re-scored on the device in %.3f ms", (unsigned long long)nRows, ms);
            }
            // the rows leave the device on a second stream (spb_align_fetch_begin) and are turned into text by a formatter thread
            // while this thread already drives the next batch through the GPU; results are freed here, by the thread that owns
            // the context, once the formatter has finished with them
            spb_align_result *r = nullptr;
            check(spb_align_fetch_begin(E.ctx, &r), "spb_align_fetch_begin");
            finish_previous();
            prevR = r;
            prevNr = nr;
            formatter = std::thread([&, r, nr, b0]() {
                try {
                    check(spb_align_fetch_wait(r), "spb_align_fetch_wait");
                    if (!nucl) {
                        format_alignment_rows(pf, r, Q, b0, T, [&](int th, size_t i, const std::string &buf) { sink(thBase + th, Q.keys[b0 + i], buf); });
                        if (batchDone) (*batchDone)(thBase, pf.threads);
                    } else {
                        std::string err;
                        format_alignment_rows(pf, r, Q, b0, T, [&](int th, size_t i, const std::string &prot) {
                            std::string nuc, comb;
                            try {
                                for (uint64_t k = r->offsets[i]; k < r->offsets[i + 1]; k++) {
                                    const spb_aln &a = r->rows[k];
                                    nucl_row_error(nr[k]);
                                    append_nucl_row(nuc, T.keys[a.target], nr[k], r->cigar_ops + a.cigar_off, a.cigar_len);
                                }
                                combine_prot_nucl_entry(prot, nuc, comb);
                            } catch (const std::exception &e) {
#pragma omp critical
                                err = e.what();
                            }
                            if (extra) (*extra)(thBase + th, Q.keys[b0 + i], prot, nuc);
                            sink(thBase + th, Q.keys[b0 + i], comb);
                        });
                        if (!err.empty()) throw Fail(err);
                        if (batchDone) (*batchDone)(thBase, pf.threads);
                    }
                } catch (...) {
                    fmtError = std::current_exception();
                }
            });
        }
        finish_previous();
    });
    pc.mark("batches (GPU + formatting)");
}

// spb200 extractframes <nuclSeqDB> <outDB> [--min-length 9 --threads N]
// The two query-preparation steps createsetdb runs back to back for spacers (data/createsetdb.sh:88-113):
//   extractorfs <db>_nucl <db>_nucl_orf --min-length 9 --orf-start-mode 1   (M/util/extractorfs.cpp:19-160)
//   translatenucs <db>_nucl_orf <db>                                        (M/util/translatenucs.cpp:13-115)
// fused on the device (spb_spacer_frames): writes <outDB>_nucl_orf (+ _h: the Orf::writeOrfHeader lines, Orf.cpp:440-457) and
// the amino-acid DB <outDB> (+ _h) with the fragment keys 0..n-1 in the reference's order.
int run_extractframes(int argc, const char **argv) {
    Params p;
    int minLength = 9;
    std::vector<const char *> rest;
    for (int i = 0; i < argc; i++) {
        if (i + 1 < argc && std::string(argv[i]) == "--min-length") { minLength = atoi(argv[++i]); continue; }
        if (i + 1 < argc && (std::string(argv[i]) == "--orf-start-mode" || std::string(argv[i]) == "--create-lookup")) { ++i; continue; }
        rest.push_back(argv[i]);
    }
    parse_args((int)rest.size(), rest.data(), 2, 2, p);
    DBReader in;
    in.open(p.db[0]);
    if (in.dbtype() != 1) throw Fail("extractframes: " + p.db[0] + " is not a nucleotide database");
    const size_t n = in.size();
    // entries in key order (extractorfs walks the DB by id); numeric nucleotides 0..3 = A C G T, everything else ambiguous
    std::vector<uint64_t> off(n + 1, 0);
    for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + (in.byKey(i).length >= 2 ? in.byKey(i).length - 2 : 0);
    std::vector<uint8_t> nt(off[n] + 1);
    uint8_t code[256];
    memset(code, 4, sizeof(code));
    code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = 3; code['U'] = code['u'] = 3;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        const char *d = in.data(in.byKey((size_t)i));
        for (uint64_t k = off[i]; k < off[i + 1]; k++) nt[k] = code[(unsigned char)d[k - off[i]]];
    }
    Engine E;
    check(spb_ctx_create(p.device, &E.ctx), "spb_ctx_create");
    spb_frames_result *fr = nullptr;
    check(spb_spacer_frames(E.ctx, nt.data(), off.data(), (uint32_t)n, minLength, 32734, 0, &fr), "spb_spacer_frames");
    log_line(p, "Fragments: %u from %zu sequences, %.3f ms on the device", fr->n_frames, n, fr->kernel_ms);
    // Orf::iupacReverseComplementTable (M/commons/Orf.cpp:48-52), letters only
    static const char *comp = "TVGH..CD..M.KN...YSAABW.R.";   // complements of 'A'..'Z'
    auto complement = [&](char c) -> char {
        if (c >= 'A' && c <= 'Z') return comp[c - 'A'];
        if (c >= 'a' && c <= 'z') return (char)tolower(comp[c - 'a']);
        return '.';
    };
    DBWriter wNuc, wNucH, wAA, wAAH;
    wNuc.open(p.db[1] + "_nucl_orf", 1, p.threads);
    wNucH.open(p.db[1] + "_nucl_orf_h", DBTYPE_GENERIC_DB, p.threads);
    wAA.open(p.db[1], DBTYPE_AMINO_ACIDS, p.threads);
    wAAH.open(p.db[1] + "_h", DBTYPE_GENERIC_DB, p.threads);
    static const char AAS[] = "ACDEFGHIKLMNPQRSTVWYX";
    const uint32_t F = fr->n_frames;
#pragma omp parallel num_threads(p.threads)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        std::string buf;
        char hdr[64];
#pragma omp for schedule(static)
        for (long f = 0; f < (long)F; f++) {
            const spb_orf &o = fr->orfs[f];
            const DBEntry &e = in.byKey(o.contig);
            const char *d = in.data(e);
            const long L = (long)(off[o.contig + 1] - off[o.contig]);
            buf.clear();
            if (o.strand == 0) {
                buf.assign(d + o.from, (size_t)(o.to - o.from + 1));
            } else {
                for (uint32_t x = o.from; x <= o.to; x++) buf.push_back(complement(d[L - 1 - (long)x]));
            }
            buf.push_back('\n');
            wNuc.write(th, (uint32_t)f, buf.data(), buf.size());
            // positions on the forward strand of the contig (extractorfs.cpp:93-99), Orf::writeOrfHeader
            long from = o.from, to = o.to;
            if (o.strand == 1) { from = (L - 1) - from; to = (L - 1) - to; }
            const int flags = (int)o.incomplete_start | ((int)o.incomplete_end << 1);
            int hl = sprintf(hdr, "%u\t%ld%c%ld", e.key, from, from < to ? '+' : '-', labs(from - to));
            if (flags) hl += sprintf(hdr + hl, "\t%d", flags);
            hdr[hl++] = '\n';
            wNucH.write(th, (uint32_t)f, hdr, (size_t)hl);
            wAAH.write(th, (uint32_t)f, hdr, (size_t)hl);
            buf.clear();
            for (uint64_t k = fr->offsets[f]; k < fr->offsets[f + 1]; k++) buf.push_back(AAS[fr->residues[k] <= 20 ? fr->residues[k] : 20]);
            buf.push_back('\n');
            wAA.write(th, (uint32_t)f, buf.data(), buf.size());
        }
    }
    wNuc.close(); wNucH.close(); wAA.close(); wAAH.close();
    spb_free_frames_result(fr);
    return 0;
}

// spb200 createindex <targetDB> <tmpDir> [prefilter flags]   (M/workflow/CreateIndex.cpp; host-only: no GPU needed)
int run_createindex(int argc, const char **argv) {
    Params p;
    set_predictmatch_defaults(p);
    parse_args(argc, argv, 2, 2, p);
    Engine E;
    E.seed.load(p.seedSubMat, 8.0, -0.2);
    DBReader tdb;
    tdb.open(p.db[0]);
    SeqSet T;
    load_sequences(tdb, E.seed, T, p.maskLower == 1);
    std::vector<uint8_t> masked;
    uint64_t nMasked = 0;
    mask_targets(E, p, T, masked, &nMasked);
    IdxHeader h;
    fill_idx_header(p, E, T, tdb.dataSize(), &h);
    const std::string tmp = p.db[0] + ".spbidx.tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) throw Fail("Could not open " + tmp + " for writing");
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(T.keys.data(), 4, T.keys.size(), f) == T.keys.size() &&
              fwrite(T.off.data(), 8, T.off.size(), f) == T.off.size() && fwrite(masked.data(), 1, (size_t)h.nResidues, f) == (size_t)h.nResidues;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), (p.db[0] + ".spbidx").c_str()) != 0) throw Fail("Could not write " + p.db[0] + ".spbidx");
    log_line(p, "Index cache %s.spbidx: %llu targets, %llu residues, %llu masked", p.db[0].c_str(), (unsigned long long)h.nTargets,
             (unsigned long long)h.nResidues, (unsigned long long)nMasked);
    return 0;
}

int run_search(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 3, p);
    check_search_params(p);
    check_align_params(p);
    DBWriter out;
    out.open(p.db[2], DBTYPE_ALIGNMENT_RES, std::max(p.threads, p.gpus * p.lanes));    // sink thread ids: worker * (threads / workers) + team thread
    // the rows of a batch go to disk as soon as they are formatted, while the next batch is on the device
    const BatchDone flush = [&](int first, int count) { out.flush(first, count); };
    search_dbs(p, p.db[0], p.db[1], [&](int th, uint32_t key, const std::string &buf) { out.write(th, key, buf.data(), buf.size()); }, false,
               nullptr, &flush);
    g_clock.mark("(join)");
    out.close();
    g_clock.mark("write DB");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// set-level stages (setscore.h): module commands with the reference's argv, and the fused predictmatch
// ---------------------------------------------------------------------------------------------------------
int run_besthitperset(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 4, p);     // <queryDB> <targetDB> <resultDB> <outDB>  (besthitperset.cpp:123-128: db2, db3, db4)
    KeyMap m2s, ssz;
    m2s.load(p.db[1] + "_member_to_set");
    ssz.load(p.db[1] + "_set_size");
    TextDB in, out;
    in.load(p.db[2]);
    best_hit_per_set(in, m2s, ssz, p.simpleBestHit != 0, p.threads, out);
    out.save(p.db[3], p.threads);
    return 0;
}

int run_mergeresultsbyset(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 3, p);     // <setDB> <resultDB> <outDB>
    TextDB sets, in, out;
    sets.load(p.db[0]);
    in.load(p.db[1]);
    merge_results_by_set(sets, in, p.threads, out);
    out.save(p.db[2], p.threads);
    return 0;
}

int run_combinescore(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 5, p);     // <queryDB> <targetDB> <resultDB> <outDB> <tmpDir>
    KeyMap m2s, qsz;
    m2s.load(p.db[1] + "_member_to_set");
    qsz.load(p.db[0] + "_set_size");
    TextDB in, out;
    in.load(p.db[2]);
    combine_score(in, m2s, qsz, p.threads, out);
    out.save(p.db[3], p.threads);
    return 0;
}

void log_fdr(const Params &p, const FdrReport &rep) {
    if (rep.mode == 0) {
        log_line(p, "Combined score threshold is %f with FDR of %f.", rep.threshold, (double)p.fdr);
        log_line(p, "%g matches passed combined score threshold.", rep.passed);
    } else if (rep.mode == 1) {
        if (p.fdr < 1) log_line(p, "Combined score list too short. Using threshold %f", 0.0);
        else log_line(p, "FDR cutoff is set to 1. Printing all matches. ");
    } else {
        log_line(p, "Combined score list of control set is empty. Printing all matches");
    }
}

int run_filtermatchbyfdr(int argc, const char **argv) {
    Params p;
    parse_args(argc, argv, 2, 3, p);     // <cScoreDB> <controlcScoreDB> <outDB>
    TextDB pos, neg, out;
    pos.load(p.db[0]);
    neg.load(p.db[1]);
    FdrReport rep;
    filter_match_by_fdr(pos, neg, p.fdr, p.reportFdr, p.threads, out, &rep);
    log_fdr(p, rep);
    out.save(p.db[2], p.threads);
    return 0;
}

// setpredictmatchDefaults (src/workflow/PredictMatch.cpp:9-31) + the search defaults the workflow passes down
// (SURVEY.md 3.2): what `predictmatch` runs when no flag is given
void set_predictmatch_defaults(Params &p) {
    p.sensitivity = 5.7f;
    p.evalue = 200;
    p.addBacktrace = 1;
    p.alignmentMode = 2;
    p.kmerSize = 6;
    p.spacedMode = 1;
    p.spacedPattern = "11011101";
    p.subMat = "nucl:nucleotide.out,aa:VTML40.out";
    p.gapOpen = "nucl:10,aa:16";
    p.gapExtend = "nucl:2,aa:2";
    p.simpleBestHit = 1;
}

// data/predictmatch.sh:27-116: search (+ the nucleotide re-alignment branch unless --perform-nucl-aln 0: proteinaln2nucl on the
// device, combineprotnuclaln) + besthitperset + mergeresultsbyset + combinescore against the target and the control setDB,
// then filtermatchbyfdr; intermediates stay in memory
int run_predictmatch(int argc, const char **argv) {
    Params p;
    set_predictmatch_defaults(p);
    parse_args(argc, argv, 2, 5, p);     // <querySetDB> <targetSetDB> <controlTargetSetDB> <matchDB> <tmpDir>
    check_search_params(p);
    check_align_params(p);
    const bool nucl = p.performNuclAln != 0;   // the reference's default (LocalParameters.h:176)
    KeyMap qsz;
    qsz.load(p.db[0] + "_set_size");
    TextDB qSets;
    qSets.load(p.db[0] + "_set_to_member");
    TextDB cScore[2];
    for (int side = 0; side < 2; side++) {
        const std::string &tName = p.db[1 + side];
        const char *suffix = side ? "_rev" : "";
        // the search rows of every query frame, keyed like the alignment DB
        DBReader qdb;
        qdb.open(p.db[0]);
        TextDB result;
        result.dbtype = DBTYPE_ALIGNMENT_RES;
        result.keys.resize(qdb.size());
        for (size_t i = 0; i < qdb.size(); i++) result.keys[i] = qdb.byKey(i).key;
        result.data.assign(result.keys.size(), std::string());
        TextDB protResult, nuclResult;
        NuclExtraSink extra = [&](int, uint32_t key, const std::string &prot, const std::string &nuc) {
            const size_t i = result.find(key);
            protResult.data[i] = prot;
            nuclResult.data[i] = nuc;
        };
        if (nucl && p.keepIntermediate) {
            protResult.dbtype = nuclResult.dbtype = DBTYPE_ALIGNMENT_RES;
            protResult.keys = nuclResult.keys = result.keys;
            protResult.data.assign(result.keys.size(), std::string());
            nuclResult.data.assign(result.keys.size(), std::string());
        }
        search_dbs(p, p.db[0], tName, [&](int, uint32_t key, const std::string &buf) { result.data[result.find(key)] = buf; }, nucl,
                   nucl && p.keepIntermediate ? &extra : nullptr);
        KeyMap m2s, ssz;
        m2s.load(tName + "_member_to_set");
        ssz.load(tName + "_set_size");
        TextDB aggregate, merged;
        best_hit_per_set(result, m2s, ssz, p.simpleBestHit != 0, p.threads, aggregate);
        merge_results_by_set(qSets, aggregate, p.threads, merged);
        combine_score(merged, m2s, qsz, p.threads, cScore[side]);
        if (p.keepIntermediate) {
            if (nucl) {
                protResult.save(p.db[4] + "/prot_result" + suffix, p.threads);
                nuclResult.save(p.db[4] + "/nucl_result" + suffix, p.threads);
            }
            result.save(p.db[4] + "/result" + suffix, p.threads);
            aggregate.save(p.db[4] + "/aggregate" + suffix, p.threads);
            merged.save(p.db[4] + "/aggregate_merged" + suffix, p.threads);
            cScore[side].save(p.db[4] + "/cScore" + suffix, p.threads);
        }
    }
    TextDB match;
    FdrReport rep;
    filter_match_by_fdr(cScore[0], cScore[1], p.fdr, p.reportFdr, p.threads, match, &rep);
    log_fdr(p, rep);
    match.save(p.db[3], p.threads);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// host-only helpers (CPU tests)
// ---------------------------------------------------------------------------------------------------------
int run_dumpmatrix(int argc, const char **argv) {
    if (argc < 5) throw Fail("usage: spb200 dumpmatrix <matrix> <bitFactor> <scoreBias>");
    SubMat m;
    m.load(argv[2], atof(argv[3]), atof(argv[4]));
    printf("%s\n", m.num2aa);
    for (int i = 0; i < SubMat::N; i++) {
        for (int j = 0; j < SubMat::N; j++) printf("%d%c", m.sub[i * SubMat::N + j], j + 1 == SubMat::N ? '\n' : ' ');
    }
    for (int i = 0; i < SubMat::N; i++) printf("%.17g%c", m.pback[i], i + 1 == SubMat::N ? '\n' : ' ');
    return 0;
}

int run_dbcopy(int argc, const char **argv) {
    if (argc < 4) throw Fail("usage: spb200 dbcopy <inDB> <outDB> [threads] [flushEvery] [compressed]");
    DBReader in;
    in.open(argv[2]);
    const int threads = argc > 4 ? atoi(argv[4]) : 1;
    DBWriter::setCompressedDefault(argc > 6 && atoi(argv[6]) == 1);
    DBWriter out;
    out.open(argv[3], in.dbtype(), threads);
    const size_t n = in.size();
    const size_t flushEvery = argc > 5 ? (size_t)atoll(argv[5]) : 0;      // streaming writer: flush all parts every N entries
    for (size_t i = 0; i < n; i++) {
        const DBEntry &e = in.byKey(i);
        out.write((int)(i * (size_t)threads / (n ? n : 1)), e.key, in.data(e), e.length ? e.length - 1 : 0);
        if (flushEvery && (i + 1) % flushEvery == 0) out.flush(0, threads);
    }
    out.close();
    printf("%zu entries, %llu residues\n", n, (unsigned long long)in.aminoAcidDBSize());
    return 0;
}

// spb200 fmtcheck <n> <seed>: put_e3 / put_int against printf on n random and adversarial values (host self-test, no GPU)
int run_fmtcheck(int argc, const char **argv) {
    const long n = argc > 2 ? atol(argv[2]) : 1000000;
    unsigned long long x = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long bad = 0;
    char a[64], b[64];
    for (long i = 0; i < n; i++) {
        double v;
        const int mode = (int)(i & 3);
        if (mode == 0) {                      // any finite positive double
            unsigned long long bits = next() & 0x7FEFFFFFFFFFFFFFull;
            memcpy(&v, &bits, 8);
        } else if (mode == 1) {               // E-value range
            v = (1.0 + (double)(next() % 9000000) / 1e6) * pow(10.0, (double)((long)(next() % 60) - 45));
        } else {                              // on and next to the rounding boundaries of four digits
            const int d = 1000 + (int)(next() % 9000), e = (int)(next() % 60) - 40;
            v = (d + (mode == 2 ? 0.5 : 0.0)) * pow(10.0, e - 3);
            if (next() & 1) v = nextafter(v, (next() & 1) ? 0.0 : 1e300);
        }
        *put_e3(a, v) = 0;
        snprintf(b, sizeof(b), "%.3E", v);
        if (strcmp(a, b)) { if (bad < 5) fprintf(stderr, "%.17g: %s vs %s\n", v, a, b); bad++; }
        const long long iv = (long long)next();
        *put_int(a, iv) = 0;
        snprintf(b, sizeof(b), "%lld", iv);
        if (strcmp(a, b)) { if (bad < 5) fprintf(stderr, "%lld: %s vs %s\n", iv, a, b); bad++; }
    }
    for (double v : {0.0, 1e-320, 5e-324, 1e308, 9.9995, 9.99949999999, 0.00099995, 1.0, 200.0}) {
        *put_e3(a, v) = 0;
        snprintf(b, sizeof(b), "%.3E", v);
        if (strcmp(a, b)) bad++;
    }
    printf("%ld values, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}

int run_showparams(int argc, const char **argv) {
    if (argc < 3) throw Fail("usage: spb200 showparams prefilter|align <args...>");
    const std::string mod = argv[2];
    Params p;
    parse_args(argc, argv, 3, mod == "align" ? 4 : 3, p);
    p.targetSplitOk = (mod == "prefilter");
    p.altAliOk = (mod == "align");
    p.realignOk = (mod == "align");
    if (mod == "align") check_align_params(p); else check_search_params(p);
    printf("sub_mat=%s seed_sub_mat=%s kmer_thr=%d max_seqs=%d min_ungapped=%d mask=%d comp_bias=%d pattern=%s "
           "evalue=%g gap_open=%s gap_extend=%s threads=%d cov=%g cov_mode=%d min_seq_id=%g min_aln_len=%d seq_id_mode=%d backtrace=%d sw_mode=%d\n",
           multiparam_aa(p.subMat).c_str(), multiparam_aa(p.seedSubMat).c_str(), kmer_threshold(p.sensitivity, p.kmerScore, 6), p.maxSeqs,
           p.minUngapped, p.mask, p.compBias, p.spacedPattern.empty() ? (p.spacedMode ? "11011101" : "111111") : p.spacedPattern.c_str(),
           p.evalue, multiparam_aa(p.gapOpen).c_str(), multiparam_aa(p.gapExtend).c_str(), p.threads, (double)(float)p.covThr, p.covMode,
           (double)(float)p.seqIdThr, p.alnLenThr, p.seqIdMode, p.addBacktrace, sw_mode_of(p));
    return 0;
}

}  // namespace

int main(int argc, const char **argv) {
    pow10_ld(0);      // the table of put_e3 is built before any formatter thread exists
    if (argc < 2) {
        fprintf(stderr, "spb200 (%s): prefilter | align | search | createindex | extractframes | besthitperset | mergeresultsbyset | combinescore | filtermatchbyfdr | predictmatch | proteinaln2nucl | combineprotnuclaln | dumpmatrix | dbcopy | showparams\n", spb_version());
        return 1;
    }
    const std::string cmd = argv[1];
    int rc = 1;
    try {
        if (cmd == "prefilter") rc = run_prefilter(argc, argv);
        else if (cmd == "align") rc = run_align(argc, argv);
        else if (cmd == "search") rc = run_search(argc, argv);
        else if (cmd == "createindex") rc = run_createindex(argc, argv);
        else if (cmd == "extractframes") rc = run_extractframes(argc, argv);
        else if (cmd == "besthitperset") rc = run_besthitperset(argc, argv);
        else if (cmd == "mergeresultsbyset") rc = run_mergeresultsbyset(argc, argv);
        else if (cmd == "combinescore") rc = run_combinescore(argc, argv);
        else if (cmd == "filtermatchbyfdr") rc = run_filtermatchbyfdr(argc, argv);
        else if (cmd == "predictmatch") rc = run_predictmatch(argc, argv);
        else if (cmd == "proteinaln2nucl") rc = run_proteinaln2nucl(argc, argv);
        else if (cmd == "combineprotnuclaln") rc = run_combineprotnuclaln(argc, argv);
        else if (cmd == "dumpmatrix") rc = run_dumpmatrix(argc, argv);
        else if (cmd == "dbcopy") rc = run_dbcopy(argc, argv);
        else if (cmd == "showparams") rc = run_showparams(argc, argv);
        else if (cmd == "fmtcheck") rc = run_fmtcheck(argc, argv);
        else {
            fprintf(stderr, "Invalid Command: %s\n", cmd.c_str());
            return 1;
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return rc;
}
