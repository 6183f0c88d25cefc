/*
 * spb200.h -- C ABI of the B200-native SpacePHARER search hot path (prefilter + align).
 *
 * The reference (soedinglab/spacepharer, MMseqs2 vendored under lib/mmseqs/src = "M/") has no FFI:
 * its seam for this path is the module CLI + on-disk DB format.  The entry points below
 * are the batched equivalents of the reference's internal per-query seams, which a maintainer would
 * bind from M/prefiltering/Prefiltering.cpp and M/alignment/Alignment.cpp:
 *
 *   spb_index_build        <->  IndexBuilder::fillDatabase            M/prefiltering/IndexBuilder.cpp:53-256
 *                               + ExtendedSubstitutionMatrix::calcScoreMatrix  ExtendedSubstitutionMatrix.cpp:20-69
 *   spb_prefilter          <->  QueryMatcher::matchQuery(Sequence*, identityId, isNucl) -> (hit_t*, n)
 *                                                                     M/prefiltering/QueryMatcher.h:64, .cpp:84-186
 *   spb_align_set_targets  <->  tdbr->getData()/Sequence::mapSequence  M/alignment/Alignment.cpp:386-392
 *   spb_align              <->  Matcher::initQuery + Matcher::getSWResult -> result_t
 *                                                                     M/alignment/Matcher.h:152-153, Matcher.cpp:61-187
 *
 * Conventions: plain pointers and sizes only; the caller owns every input buffer (host memory; pinned
 * memory makes the copies asynchronous); the library owns outputs until the matching spb_free_* call;
 * every function returns 0 on success, non-zero on failure with spb_last_error() describing it (the host
 * `main` turns that into the reference's Debug(Debug::ERROR)...EXIT(EXIT_FAILURE) behaviour).  No C++
 * exceptions cross this boundary.  A context is bound to one CUDA device and is not thread-safe (one
 * host thread per context, like one QueryMatcher / Matcher per OpenMP thread in the reference).
 * There is no CPU fallback: without a CUDA device spb_ctx_create fails.
 */
#ifndef SPB200_H
#define SPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_ALPHABET 21   /* ACDEFGHIKLMNPQRSTVWYX, X = 20 (lib/mmseqs/data/VTML40.out:13) */

typedef struct spb_ctx spb_ctx;

/* ---- life cycle ------------------------------------------------------------------------------- */
int spb_ctx_create(int device, spb_ctx **out);
void spb_ctx_destroy(spb_ctx *ctx);
/* A lane of `parent`: a context on the same device with its own stream, query batch, results and scratch that READS the
 * parent's matrices, k-mer index, alignment targets and nucleotide targets in place (nothing is copied).  One host thread per
 * lane, like one QueryMatcher / Matcher per OpenMP thread in the reference (Prefiltering.cpp:776-800, Alignment.cpp:285-321);
 * batches submitted to different lanes overlap on the GPU.  Create lanes after the parent's index / targets are set; do not
 * rebuild those or destroy the parent while a lane exists.  A lane cannot build an index or set targets itself. */
int spb_ctx_create_lane(spb_ctx *parent, spb_ctx **out);
/* number of CUDA devices visible to the process (0 without a driver / device) */
int spb_device_count(void);
const char *spb_last_error(void);
const char *spb_version(void);
/* the CUDA stream all work of this context is enqueued on (cudaStream_t as void*) */
void *spb_ctx_stream(spb_ctx *ctx);
/* number of this library's own CUDA kernels launched on the context so far (cub library kernels not counted) */
uint64_t spb_ctx_launches(spb_ctx *ctx);

/* ---- substitution matrices ---------------------------------------------------------------------
 * int16 scores as the reference's SubstitutionMatrix produces them (row-major [21][21]):
 *   sub_seed     k-mer matrix      VTML80, bit factor 8, bias -0.2   M/prefiltering/Prefiltering.cpp:64
 *   sub_ungapped ungapped matrix   VTML40, bit factor 2, bias -0.2   M/prefiltering/Prefiltering.cpp:65
 *   sub_align    alignment matrix  VTML40, bit factor 2, bias  0     M/alignment/Alignment.cpp:159
 * pback_* = BaseMatrix::pBack of the seed / alignment matrix (composition-bias correction). */
int spb_set_matrices(spb_ctx *ctx, const int16_t *sub_seed, const double *pback_seed,
                     const int16_t *sub_ungapped, const int16_t *sub_align, const double *pback_align);

/* ---- target index ------------------------------------------------------------------------------
 * residues: numeric target residues (0..20) as the index must see them, i.e. already tantan-masked when
 * --mask 1 (host pre-step, M/prefiltering/IndexBuilder.cpp:137-149); offsets[T+1]; ids are positions in
 * this array (DBReader LINEAR_ACCESS order).  pattern: spaced k-mer pattern with six '1' ("11011101").
 * kmer_thr: index-side self-score threshold (M/prefiltering/Prefiltering.cpp:510-513).
 * Builds, resident in HBM: masked residue store (SequenceLookup), k-mer -> sorted (id,pos) postings
 * (IndexTable) and the 3-mer similarity rows (ScoreMatrix). */
int spb_index_build(spb_ctx *ctx, const uint8_t *residues, const uint64_t *offsets, uint32_t T,
                    const char *pattern, int kmer_thr);

typedef struct spb_index_stats {
    uint64_t n_targets, n_residues, n_postings, n_kmers_nonempty;
    uint64_t bytes_device;      /* HBM held by the index */
    double build_ms;            /* device time of the last build */
} spb_index_stats;
int spb_index_stats_get(spb_ctx *ctx, spb_index_stats *out);
/* copy the postings out as sorted (kmer, id, pos) triples; arrays sized n_postings (test hook) */
int spb_index_export(spb_ctx *ctx, uint32_t *kmer, uint32_t *id, uint16_t *pos);
/* similar k-mer list of one 6-mer at threshold thr (test hook for KmerGenerator parity); returns count in *n */
int spb_similar_kmers(spb_ctx *ctx, const uint8_t *kmer6, int thr, uint32_t *out, uint32_t cap, uint32_t *n);

/* ---- prefilter --------------------------------------------------------------------------------- */
typedef struct spb_prefilter_params {
    int32_t kmer_thr;        /* -s 5.7 => 112 (Prefiltering::getKmerThreshold) */
    int32_t max_seqs;        /* --max-seqs (300) */
    int32_t min_diag_score;  /* --min-ungapped-score (15) */
    int32_t bin_size;        /* CacheFriendlyOperations BINSIZE of the emulated host (2..2048, power of 2);
                                M/prefiltering/QueryMatcher.cpp:383-411 -- only affects tie order under truncation */
    int32_t comp_bias;       /* --comp-bias-corr */
    uint32_t id_base;        /* added to local target ids for tie-order purposes (target shards) */
    int32_t emit_trunc_keys; /* 1: keep the max-seqs truncation key of every hit on the device (target shards, see
                                spb_prefilter_device_trunc_keys); requires max_seqs <= 4096 */
    int32_t exact_kmer;      /* --exact-kmer-matching 1: every window looks up only its own k-mer (takeOnlyBestKmer,
                                Prefiltering.cpp:177, QueryMatcher.cpp:221-231); served by the one-CTA-per-query kernel */
} spb_prefilter_params;

typedef struct spb_hit {     /* hit_t, M/prefiltering/QueryMatcher.h:35-52 */
    uint32_t id;             /* target index (position in the array given to spb_index_build) */
    int32_t score;           /* ungapped diagonal score */
    int32_t diag;            /* (short) diagonal */
} spb_hit;

typedef struct spb_prefilter_result {
    uint32_t n_queries;
    uint64_t *offsets;       /* [n_queries+1] */
    spb_hit *hits;           /* sorted per query by (score desc, id asc) like compareHitsByScoreAndId */
    uint64_t n_hits;
    /* counters (Prefiltering.cpp:934-957 prints the same statistics) */
    uint64_t n_kmers, n_db_matches, n_events, n_slow_path;
    double kernel_ms;        /* device time of the match kernel(s) */
    double total_ms;         /* device time of the whole resident stage */
    uint64_t algo_bytes;     /* algorithmic bytes moved by the match kernel */
    double expand_ms, lookup_ms, bookkeep_ms;   /* device time of the three kernel classes inside kernel_ms */
} spb_prefilter_result;

int spb_prefilter(spb_ctx *ctx, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                  const spb_prefilter_params *p, spb_prefilter_result **out);
void spb_free_prefilter_result(spb_prefilter_result *r);

/* ---- alignment --------------------------------------------------------------------------------- */
typedef struct spb_gumbel {  /* ALP parameter set (lib/mmseqs/lib/alp/sls_alignment_evaluer.hpp) */
    double lambda, K, a_I, b_I, alpha_I, beta_I, a_J, b_J, alpha_J, beta_J, sigma, tau;
} spb_gumbel;

typedef struct spb_align_params {
    int32_t gap_open, gap_extend;   /* aa:16 / 2 */
    double eval_thr;                /* -e */
    uint64_t db_res_count;          /* DBReader::getAminoAcidDBSize of the WHOLE target DB (Alignment.cpp:269) */
    int32_t comp_bias;
    int32_t byte_lanes, word_lanes; /* SIMD layout of the emulated reference build: 32/16 (AVX2), 16/8 (SSE) */
    spb_gumbel gumbel;
    /* Acceptance criteria beyond the E-value (all 0 in the predictmatch search): a pair is reported when its lengths can reach
     * the coverage (Util::canBeCovered, M/commons/Util.cpp:533-550, tested before the alignment in Alignment.cpp:378) and its
     * alignment passes Alignment::checkCriteria (Alignment.cpp:555-574; Util::hasCoverage Util.cpp:552-567). */
    float cov_thr;                  /* -c */
    int32_t cov_mode;               /* --cov-mode: 0 query and target, 1 target, 2 query, 3 / 4 / 5 length ratios (Parameters.h:246-251) */
    float seq_id_thr;               /* --min-seq-id */
    int32_t aln_len_thr;            /* --min-aln-len */
    int32_t seq_id_mode;            /* --seq-id-mode: identities / 0 alignment length, 1 shorter, 2 longer sequence (Util.cpp:588-598) */
    int32_t sw_mode;                /* 0: score, coverage, sequence identity, backtrace (-a 1 or --alignment-mode 3: what predictmatch runs);
                                       1: score + coverage (--alignment-mode 2 without -a: no backtrace, identity estimated from the score);
                                       2: score only (--alignment-mode 1: no start positions either) -- Alignment::initSWMode, Alignment.cpp:174-196 */
    int32_t max_accept;             /* --max-accept / --max-rejected (Alignment.cpp:353-404): the pair list of a query is walked in the caller's */
    int32_t max_rejected;           /* order until max_accept pairs passed or max_rejected pairs in a row failed; <= 0 or INT32_MAX: no limit */
} spb_align_params;

typedef struct spb_aln {     /* Matcher::result_t, M/alignment/Matcher.h:24-119 */
    uint32_t query;          /* query index in the batch */
    uint32_t target;         /* target index */
    int32_t score;           /* raw Smith-Waterman score */
    int32_t bits;            /* (int)(bitScore + 0.5) */
    double eval;
    int32_t q_start, q_end, t_start, t_end;   /* 0-based inclusive */
    int32_t n_ident, aln_len;
    float seq_id;
    uint32_t cigar_off, cigar_len;  /* slice of spb_align_result::cigar_ops (one op char per column: M/I/D) */
    uint32_t reserved;              /* 0 (the struct is 64 bytes; rows are bytewise reproducible) */
} spb_aln;

typedef struct spb_align_result {
    uint32_t n_queries;
    uint64_t *offsets;       /* [n_queries+1] accepted rows per query */
    spb_aln *rows;           /* per query sorted like Matcher::compareHits (eval asc, bits desc, tLen asc, id asc) */
    uint64_t n_rows;
    char *cigar_ops;
    uint64_t n_cigar;
    uint64_t n_pairs, n_cells;      /* pairs aligned, forward-pass cells (sum qLen*tLen) */
    double kernel_ms;               /* device time of the forward-pass kernel(s) */
    double total_ms;
} spb_align_result;

/* unmasked numeric target residues (align reads the raw DB, not the masked lookup).  keys (nullable): DB key of
 * every target, used only as the last tie-break of the row order (Matcher::compareHits sorts by dbKey). */
int spb_align_set_targets(spb_ctx *ctx, const uint8_t *residues, const uint64_t *offsets, uint32_t T,
                          const uint32_t *keys);
/* pairs given CSR-style: query i is aligned against pair_targets[pair_offsets[i] .. pair_offsets[i+1]) */
int spb_align(spb_ctx *ctx, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
              const uint64_t *pair_offsets, const uint32_t *pair_targets,
              const spb_align_params *p, spb_align_result **out);
void spb_free_align_result(spb_align_result *r);

/* E-value / bit score of one raw score on the device path's formula, evaluated on the host in double
 * (EvalueComputation::computeEvalue / computeBitScore, M/alignment/EvalueComputation.h:18-40) */
double spb_evalue(const spb_gumbel *g, double score, double q_len, double db_res_count);
double spb_bitscore(const spb_gumbel *g, double score);

/* ---- resident pipeline ----------------------------------------------------------------------------
 * The calls above are compositions of the stages below.  A caller that keeps the batch in HBM (the fused
 * predictmatch search: prefilter -> align without the pref DB round trip of blastp.sh:33,41) uses them directly:
 *   spb_queries_set            upload a query batch once
 *   spb_prefilter_resident     hit lists stay on the device (spb_prefilter_fetch copies them out)
 *   spb_pairs_from_prefilter   the resident hit lists become the alignment work list (no host round trip)
 *   spb_pairs_set              ... or upload an explicit pair list (what `align` reads from the pref DB)
 *   spb_align_resident         rows and CIGARs stay on the device (spb_align_fetch copies them out, sorted)
 * spb_*_stats return the counters of the last resident run (pointer members NULL). */
int spb_queries_set(spb_ctx *ctx, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q);
int spb_prefilter_resident(spb_ctx *ctx, const spb_prefilter_params *p);
int spb_prefilter_stats(spb_ctx *ctx, spb_prefilter_result *out);
int spb_prefilter_fetch(spb_ctx *ctx, spb_prefilter_result **out);
int spb_pairs_set(spb_ctx *ctx, const uint64_t *pair_offsets, const uint32_t *pair_targets);
int spb_pairs_from_prefilter(spb_ctx *ctx);
int spb_align_resident(spb_ctx *ctx, const spb_align_params *p);
int spb_align_stats(spb_ctx *ctx, spb_align_result *out);
int spb_align_fetch(spb_ctx *ctx, spb_align_result **out);
/* pipelined fetch for batch streams (1 M spacers = many batches): _begin snapshots the resident result (device-to-device,
 * on the context's stream) and starts the copy to pinned host memory on a second stream, then returns; the caller may
 * immediately upload and run the next batch.  The result's pointers are valid after spb_align_fetch_wait(result); free it
 * with spb_free_align_result as usual.  (The reference's equivalent is the writer thread behind DBWriter::writeData.) */
int spb_align_fetch_begin(spb_ctx *ctx, spb_align_result **out);
int spb_align_fetch_wait(spb_align_result *result);

/* ---- exhaustive search (BASELINE config 5) ---------------------------------------------------------
 * Every resident query against EVERY alignment target, no k-mer stage.  The reference equivalent is `align`
 * (M/alignment/Alignment.cpp:263-520) fed with a prefilter DB that lists all targets for every query (SURVEY.md 0 #4,
 * 8d; at this MMseqs2 commit --exhaustive-search itself is the sliced, swapped search of M/workflow/Search.cpp:317-351).
 * spb_pairs_exhaustive scores all Q x T pairs with the streaming register kernel and leaves, as the resident pair
 * list, the pairs whose Smith-Waterman score reaches the E-value threshold (queries too long for the register kernel
 * keep every target); spb_align_resident / spb_align_fetch then produce the rows exactly as for an explicit list. */
typedef struct spb_exhaustive_stats {
    uint64_t n_queries, n_targets;
    uint64_t n_scored;       /* pairs scored by the streaming kernel */
    uint64_t n_cells;        /* their forward cells (sum qLen * tLen) */
    uint64_t n_kept;         /* of those, pairs whose score reaches the threshold */
    uint64_t n_unfiltered;   /* pairs of long queries, passed on without filtering */
    uint64_t n_pairs;        /* resident pair list = n_kept + n_unfiltered */
    double kernel_ms;        /* device time of the streaming kernel launches */
    double total_ms;
} spb_exhaustive_stats;
int spb_pairs_exhaustive(spb_ctx *ctx, const spb_align_params *p, spb_exhaustive_stats *stats);
int spb_align_exhaustive(spb_ctx *ctx, const uint8_t *q_residues, const uint64_t *q_offsets, uint32_t Q,
                         const spb_align_params *p, spb_align_result **out, spb_exhaustive_stats *stats);

/* ---- query preparation: nucleotide spacers -> six-frame peptide fragments ---------------------------
 * The spacer query set is prepared by createsetdb --extractorf-spacer 1 (data/createsetdb.sh:88-113) as
 *   extractorfs --min-length 9 --orf-start-mode 1   M/util/extractorfs.cpp:19-160, Orf::findForward M/commons/Orf.cpp:230-352
 *   translatenucs (canonical table)                 M/util/translatenucs.cpp:13-115
 * spb_spacer_frames does both on the device.  nt: numeric nucleotides 0..3 = A C G T, anything else = ambiguous (N);
 * nt_offsets[N+1].  Fragments come out in the reference's order (per sequence: forward strand then reverse complement,
 * by the position of the codon that ends them), i.e. entry i here is DB key i of the reference's <db>_nucl_orf / <db>.
 * residues are numeric amino acids (SPB_ALPHABET order).  make_resident != 0 additionally installs the fragments as
 * the resident query batch (as spb_queries_set would), so spacers go from nucleotides to the search in HBM. */
typedef struct spb_orf {     /* Orf::SequenceLocation, M/commons/Orf.h */
    uint32_t contig;         /* index of the input sequence */
    uint32_t from, to;       /* first / last nucleotide of the fragment on its strand (0-based, inclusive) */
    uint8_t strand;          /* 0 forward, 1 reverse complement */
    uint8_t incomplete_start, incomplete_end, pad;
} spb_orf;
typedef struct spb_frames_result {
    uint32_t n_sequences, n_frames;
    uint64_t n_residues;
    uint64_t *offsets;       /* [n_frames+1] */
    uint8_t *residues;
    spb_orf *orfs;           /* [n_frames] */
    double kernel_ms;
} spb_frames_result;
int spb_spacer_frames(spb_ctx *ctx, const uint8_t *nt, const uint64_t *nt_offsets, uint32_t N, int min_codons,
                      int max_codons, int make_resident, spb_frames_result **out);
void spb_free_frames_result(spb_frames_result *r);

/* ---- nucleotide re-alignment branch of predictmatch (--perform-nucl-aln 1, the reference's default) ------
 * data/predictmatch.sh:27-45,74-92 runs, after each protein search,
 *   proteinaln2nucl     lib/mmseqs/src/util/proteinaln2nucl.cpp:84-175   every alignment row is re-scored on the nucleotide
 *                       ORFs (<db>_nucl_orf, same keys as <db>) along its protein backtrace: coordinates x 3, M columns scored
 *                       with nucleotide.out (+2 / -3) and counted for identity, gap runs cost open + ((len-1)*3)*extend,
 *                       bit score / E-value from the nucleotide ALP set with the nucleotide query length
 *   combineprotnuclaln  src/util/CombineProtNuclAln.cpp:55-95           (text only; host side)
 * Here the nucleotide ORFs stay resident next to the alignment targets and one kernel re-scores all rows of a batch.
 * nt: raw ASCII nucleotides of the entries back to back (identity is a character comparison, as in the reference),
 * offsets[n+1], start_codon (nullable): 1 where the amino-acid entry begins with '*'.  Target index i / query index i is
 * the same index spb_align_set_targets / the query batch uses. */
typedef struct spb_nucl_params {
    int32_t gap_open, gap_extend;   /* nucl:10 / 2 */
    uint64_t db_res_count;          /* DBReader::getAminoAcidDBSize of the target <db>_nucl_orf */
    spb_gumbel gumbel;              /* ALP set of the nucleotide matrix and gap costs */
    uint8_t letter[256];            /* NucleotideMatrix::setupLetterMapping, ASCII -> 0..4 (A C T G X) */
    int16_t sub[25];                /* nucleotide matrix [5][5], bit factor 1, bias 0 */
} spb_nucl_params;
typedef struct spb_nucl_aln {       /* the Matcher::result_t proteinaln2nucl writes */
    int32_t raw_score;              /* nucleotide score along the backtrace (may be negative) */
    int32_t bits;                   /* (int) computeBitScore(raw): truncated */
    double eval;
    int32_t q_start, q_end, t_start, t_end;
    uint32_t q_len, t_len;          /* nucleotide lengths */
    uint32_t n_ident, aln_len;
    float seq_id;
    uint32_t err;                   /* 0 ok; 1 a start codon '*' is aligned (the reference exits: "Alignment contains
                                       unalignable character"); 2 the backtrace leaves the nucleotide entry; 3 bad index */
} spb_nucl_aln;
int spb_nucl_set_targets(spb_ctx *ctx, const uint8_t *nt, const uint64_t *offsets, uint32_t T, const uint8_t *start_codon);
int spb_nucl_set_queries(spb_ctx *ctx, const uint8_t *nt, const uint64_t *offsets, uint32_t Q, const uint8_t *start_codon);
/* re-score explicit rows (what the module reads from an alignment DB): rows[i].query / .target index the resident
 * nucleotide entries, rows[i].cigar_off / cigar_len slice cigar_ops (one op per column).  *out: n_rows results owned by
 * the context until spb_free_nucl_result. */
int spb_nucl_rescore(spb_ctx *ctx, const spb_nucl_params *p, const spb_aln *rows, uint64_t n_rows, const char *cigar_ops,
                     uint64_t n_cigar, spb_nucl_aln **out);
/* re-score the resident alignment result (after spb_align_resident; the rows never leave HBM): out[i] belongs to row i
 * of spb_align_fetch.  kernel_ms (nullable): device time of the kernel. */
int spb_nucl_rescore_resident(spb_ctx *ctx, const spb_nucl_params *p, spb_nucl_aln **out, uint64_t *n_rows, double *kernel_ms);
void spb_free_nucl_result(spb_ctx *ctx, spb_nucl_aln *r);

/* ---- multi-GPU target split ------------------------------------------------------------------------
 * MMseqs2 shards targets with --split (Prefiltering.cpp:258-362) and merges the per-split result files
 * (mergeTargetSplits, Prefiltering.cpp:364-481).  Here every GPU prefilters all queries against its target shard
 * (spb_prefilter_params.id_base = first global id of the shard), the per-query lists are exchanged device to device
 * (NCCL all-to-all by query range, done by the caller) and merged on the owning GPU:
 *   spb_prefilter_device_result      device pointers of the resident result (hits, offsets[Q+1]) for the exchange
 *   spb_prefilter_device_trunc_keys  device pointer of the hits' truncation keys (u64 per hit; run the prefilter with
 *                                    emit_trunc_keys = 1).  The reference cuts a list at max-seqs in the order of its
 *                                    stable counting sorts (QueryMatcher.cpp:131-164, CacheFriendlyOperations.cpp:
 *                                    311-341): capped score desc, id & (BINSIZE-1), stream order of the winning
 *                                    double-hit event, id.  The key packs the first three (shard independent); with
 *                                    the global id it orders the union of the shards' lists like the unsplit list.
 *   spb_prefilter_merge_shards       d_recv = shard-major concatenation of the lists of queries q_lo..q_hi (global
 *                                    target ids), d_recv_keys their truncation keys in the same layout (or NULL),
 *                                    d_counts[shard][q - q_lo] their lengths (device pointers).  The merged lists
 *                                    (score desc, id asc), cut at max_seqs in the reference's order, replace the
 *                                    resident result: empty outside the range.  Returns 2 if a query reaches max_seqs
 *                                    and d_recv_keys is NULL, or if its cut falls among saturated 8-bit counters (the
 *                                    rescoreHits corner, QueryMatcher.cpp:486-505, needs the unsplit candidate set):
 *                                    use query split for that batch. */
int spb_prefilter_device_result(spb_ctx *ctx, spb_hit **d_hits, uint64_t **d_offsets, uint64_t *n_hits);
int spb_prefilter_device_trunc_keys(spb_ctx *ctx, uint64_t **d_keys);
int spb_prefilter_merge_shards(spb_ctx *ctx, uint32_t q_lo, uint32_t q_hi, const spb_hit *d_recv, const uint64_t *d_recv_keys,
                               uint64_t n_recv, const uint32_t *d_counts, uint32_t n_shards, int32_t max_seqs);

/* ---- multi-GPU index split: the k-mer table sharded over the GPUs' HBM, read through NVLink peer memory ------
 * The second way to spread one target DB over the GPUs of a box (MMseqs2 only has the target split above, because a CPU
 * node cannot read its neighbours' memory): the k-mer table -- the 16-byte-per-k-mer direct table plus the posting lists,
 * the bulk of IndexTable (M/prefiltering/IndexTable.h:65-76) -- is cut into `world` contiguous k-mer ranges, GPU r keeps
 * range r, and the lookup kernel of every GPU gathers the entries of the other ranges with plain loads on peer-mapped
 * pointers (NVLink / NVSwitch).  Queries are partitioned (every GPU runs expand -> lookup -> bookkeeping -> align for its own
 * queries), so no hit list is exchanged or merged and the result is the unsplit result by construction; only the 8 MB
 * occupancy bitmap and the residues are replicated.  Works across processes (CUDA IPC handles, one process per GPU) and inside
 * one process (several contexts; peer access is enabled between their devices).
 *   spb_index_shard          after spb_index_build: keep k-mer range `rank` of `world`, free the rest of the table
 *   spb_index_shard_export   CUDA IPC handles of this context's shard (2 x 64 bytes: table, posting lists)
 *   spb_index_shard_attach_ipc   map the shard of rank `peer` from its exported handles (another process)
 *   spb_index_shard_attach_ctx   ... or from a context of the same process
 * Every rank must be attached before the first prefilter call; lanes of a context inherit its shard map. */
int spb_index_shard(spb_ctx *ctx, int rank, int world);
int spb_index_shard_export(spb_ctx *ctx, void *handles128);
int spb_index_shard_attach_ipc(spb_ctx *ctx, int peer, const void *handles128);
int spb_index_shard_attach_ctx(spb_ctx *ctx, int peer, spb_ctx *peer_ctx);

/* ---- host pre-step (CPU, as in the reference) -----------------------------------------------------
 * tantan repeat masking of the target residues before the index build (--mask 1; call site
 * M/prefiltering/IndexBuilder.cpp:137-149, M/commons/tantan.cpp).  prob = joint probabilities P_ab [21][21] and
 * pback[21] of the seed matrix (BaseMatrix::probMatrix / pBack).  Masks in place (residue -> X); OpenMP over
 * targets.  Not a device fallback: the reference keeps this O(50 x residues) double-precision HMM on the host too. */
int spb_tantan_mask(const double *prob, const double *pback, uint8_t *residues, const uint64_t *offsets, uint32_t T,
                    int threads, uint64_t *n_masked);

#ifdef __cplusplus
}
#endif
#endif
