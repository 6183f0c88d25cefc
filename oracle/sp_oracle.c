/*
 * sp_oracle.c -- CPU restatement of the SpacePHARER / MMseqs2 prefilter+align hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sp_oracle.h).  Scalar C, no SIMD; the striped Smith-Waterman of
 * the reference is emulated lane by lane so that layout-dependent results are reproduced.
 * "M/" abbreviates /root/reference/lib/mmseqs/src/.
 *
 * Compile with -ffp-contract=off: the reference build (-mavx2, no -mfma) never fuses a*b+c.
 */
#include "sp_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * substitution matrix
 * ---------------------------------------------------------------------------------------- */

static int split_words(char *line, char **words, int maxWords) {
    int n = 0;
    char *p = line;
    while (*p && n < maxWords) {
        while (*p == ' ' || *p == '\t' || *p == '\r') p++;
        if (*p == '\0' || *p == '\n') break;
        words[n++] = p;
        while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++;
        if (*p) { *p = '\0'; p++; }
    }
    return n;
}

/* SubstitutionMatrix ctor + readProbMatrix + setupLetterMapping + BaseMatrix::generateSubMatrix
 * (M/commons/SubstitutionMatrix.cpp:13-51,257-297,327-407,409-441; M/commons/BaseMatrix.cpp:96-156).
 * Only matrices carrying the precomputed "# Background" and "# Lambda" header lines are handled
 * (all matrices shipped with the reference do). */
int spo_matrix_parse(const char *text, double bitFactor, double scoreBias, spo_matrix *m) {
    memset(m, 0, sizeof(*m));
    memset(m->aa2num, 0xFF, sizeof(m->aa2num));
    size_t n = strlen(text);
    char *buf = (char *)malloc(n + 2);
    double raw[SPO_ALPH][SPO_ALPH];
    memset(raw, 0, sizeof(raw));
    int hasLambda = 0, hasBg = 0, started = 0, alph = 0;
    const char *p = text;
    while (*p) {
        const char *e = strchr(p, '\n');
        size_t len = e ? (size_t)(e - p) : strlen(p);
        memcpy(buf, p, len);
        buf[len] = '\0';
        p += len + (e ? 1 : 0);
        char *words[256];
        if (buf[0] == '#') {
            if (strncmp(buf, "# Background (precomputed optional):", 36) == 0) {
                int w = split_words(buf, words, 256);
                for (int i = 4; i < w && i - 4 < SPO_ALPH; i++) m->pBack[i - 4] = strtod(words[i], NULL);
                hasBg = 1;
            } else if (strncmp(buf, "# Lambda     (precomputed optional):", 36) == 0) {
                int w = split_words(buf, words, 256);
                if (w > 4) { m->lambda = strtod(words[4], NULL); hasLambda = 1; }
            }
            continue;
        }
        int w = split_words(buf, words, 256);
        if (w <= 1) continue;
        if (!started) {
            /* alphabet header (setAaMappingDetectAlphSize) */
            if (w > SPO_ALPH) { free(buf); return -1; }
            for (int i = 0; i < w; i++) {
                int aa = toupper((unsigned char)words[i][0]);
                m->aa2num[aa] = (unsigned char)i;
                m->num2aa[i] = (char)aa;
            }
            alph = w;
            started = 1;
            continue;
        }
        int row = m->aa2num[toupper((unsigned char)words[0][0])];
        if (row >= alph) { free(buf); return -2; }
        for (int i = 0; i < alph && i + 1 < w; i++) raw[row][i] = strtod(words[i + 1], NULL);
    }
    free(buf);
    if (alph != SPO_ALPH || m->aa2num['X'] != SPO_X || !hasLambda || !hasBg) return -3;
    m->alph = alph;
    int xIsPositive = 0;
    for (int j = 0; j < alph; j++)
        if (raw[SPO_X][j] > 0 || raw[j][SPO_X] > 0) { xIsPositive = 1; break; }
    if (!xIsPositive)
        for (int i = 0; i < alph - 1; i++) m->pBack[i] = m->pBack[i] * (1.0 - m->pBack[SPO_X]);
    for (int i = 0; i < alph; i++)
        for (int j = 0; j < alph; j++) m->prob[i][j] = exp(m->lambda * raw[i][j]) * m->pBack[i] * m->pBack[j];
    /* letter mapping (setupLetterMapping) */
    unsigned char base[256];
    memcpy(base, m->aa2num, 256);
    for (int letter = 0; letter < 255; letter++) {
        int up = toupper(letter);
        unsigned char v;
        switch (up) {
            case 'A': case 'T': case 'G': case 'C': case 'D': case 'E': case 'F': case 'H': case 'I':
            case 'K': case 'L': case 'M': case 'N': case 'P': case 'Q': case 'R': case 'S': case 'V':
            case 'W': case 'Y': case 'X': v = base[up]; break;
            case 'J': v = base['L']; break;
            case 'U': case 'O': v = base['X']; break;
            case 'Z': v = base['E']; break;
            case 'B': v = base['D']; break;
            default: v = base['X']; break;
        }
        m->aa2num[letter] = v;
    }
    m->aa2num[255] = base['X'];
    /* generateSubMatrix: background = row sums (X forced to 1e-5), score = log2(P/(pa pb)) */
    double bg[SPO_ALPH];
    for (int i = 0; i < alph; i++) {
        bg[i] = 0;
        for (int j = 0; j < alph; j++) bg[i] += m->prob[i][j];
    }
    bg[alph - 1] = 1E-5;
    for (int i = 0; i < alph; i++)
        for (int j = 0; j < alph; j++) {
            double s = log2(m->prob[i][j] / (bg[i] * bg[j]));
            double v = bitFactor * s + scoreBias;
            m->sub[i][j] = (short)((v < 0.0) ? v - 0.5 : v + 0.5);
        }
    return 0;
}

int spo_map_sequence(const spo_matrix *m, const char *s, int n, unsigned char *out) {
    int l = 0;
    while (l < n && s[l] != '\0' && s[l] != '\n') {
        out[l] = m->aa2num[(unsigned char)s[l]];
        l++;
    }
    return l;
}

/* SubstitutionMatrix::calcLocalAaBiasCorrection, M/commons/SubstitutionMatrix.cpp:85-114.
 * float accumulator; the division and every pBack*score term are evaluated in double. */
void spo_comp_bias(const spo_matrix *m, const unsigned char *seq, int L, float *out) {
    for (int i = 0; i < L; i++) {
        int lo = i - 20 > 0 ? i - 20 : 0;
        int hi = i + 20 < L ? i + 20 : L;
        int W = hi - lo;
        const short *row = m->sub[seq[i]];
        int sum = 0;
        for (int j = lo; j < hi; j++) sum += row[seq[j]];
        sum -= row[seq[i]];
        float d = (float)sum;
        d = (float)((double)d / (-1.0 * (double)(float)W));
        for (int a = 0; a < m->alph; a++) d = (float)((double)d + m->pBack[a] * (double)(float)row[a]);
        out[i] = d;
    }
}

/* Prefiltering::getKmerThreshold, M/prefiltering/Prefiltering.cpp:987-1021 (k = 6, sequence search) */
int spo_kmer_threshold(float sensitivity) {
    float base = 163.2;
    double best = base - (sensitivity * 8.917);
    return (int)best;
}

/* ------------------------------------------------------------------------------------------
 * tantan (gapless path), M/commons/tantan.cpp:221-235,318-351,365-396; call site
 * M/prefiltering/IndexBuilder.cpp:137-149; likelihood ratios M/commons/BaseMatrix.h:80-109
 * ---------------------------------------------------------------------------------------- */
int spo_tantan_mask(const spo_matrix *m, unsigned char *seq, int L) {
    enum { MAXOFF = 50, STEP = 16 };
    if (L <= 0) return 0;
    const double repeatProb = 0.005, repeatEndProb = 0.05, decay = 0.9, minMaskProb = 0.9;
    double lr[SPO_ALPH][SPO_ALPH];
    for (int i = 0; i < m->alph; i++)
        for (int j = 0; j < m->alph; j++) lr[i][j] = m->prob[i][j] / (m->pBack[i] * m->pBack[j]);
    const double b2b = 1 - repeatProb, f2b = repeatEndProb, f2f0 = 1 - repeatEndProb;
    const double b2fFirst = repeatProb * ((1 - decay) / (1 - pow(decay, MAXOFF)));
    double b2f[MAXOFF], fg[MAXOFF];
    double pp = b2fFirst;
    for (int i = 0; i < MAXOFF; i++) { b2f[i] = pp; pp *= decay; }
    float *prob = (float *)malloc(sizeof(float) * (size_t)L);
    double *scale = (double *)calloc((size_t)L / STEP + 1, sizeof(double));
    double B = 1.0;
    for (int i = 0; i < MAXOFF; i++) fg[i] = 0.0;
    for (int p = 0; p < L; p++) {
        const double *row = lr[seq[p]];
        int n = p < MAXOFF ? p : MAXOFF;
        double b = B, from = 0;
        for (int i = 0; i < n; i++) {
            double f = fg[i];
            from += f;
            fg[i] = (b * b2f[i] + f * f2f0) * row[seq[p - i - 1]];
        }
        B = b * b2b + from * f2b;
        if (p % STEP == STEP - 1) {
            double s = 1 / B;
            scale[p / STEP] = s;
            B *= s;
            for (int i = 0; i < MAXOFF; i++) fg[i] *= s;
        }
        prob[p] = (float)B;
    }
    double sum = 0.0;
    for (int i = 0; i < MAXOFF; i++) sum += fg[i];
    double z = B * b2b + sum * f2b;
    B = b2b;
    for (int i = 0; i < MAXOFF; i++) fg[i] = f2b;
    for (int p = L - 1; p >= 0; p--) {
        double nonRep = prob[p] * B / z;
        prob[p] = 1 - (float)nonRep;
        if (p % STEP == STEP - 1) {
            double s = scale[p / STEP];
            B *= s;
            for (int i = 0; i < MAXOFF; i++) fg[i] *= s;
        }
        const double *row = lr[seq[p]];
        int n = p < MAXOFF ? p : MAXOFF;
        double toB = f2b * B, toF = 0;
        for (int i = 0; i < n; i++) {
            double f = fg[i] * row[seq[p - i - 1]];
            toF += b2f[i] * f;
            fg[i] = toB + f2f0 * f;
        }
        B = b2b * B + toF;
    }
    int masked = 0;
    for (int p = 0; p < L; p++)
        if (prob[p] >= minMaskProb) { seq[p] = SPO_X; masked++; }
    free(prob);
    free(scale);
    return masked;
}

/* ------------------------------------------------------------------------------------------
 * prefilter
 * ---------------------------------------------------------------------------------------- */
#define N3 8000        /* 20^3 */
#define NK 64000000u   /* 20^6 */

typedef struct row3 {
    short score[N3];
    unsigned short idx[N3];
} row3;

struct spo_prefilter {
    spo_prefilter_params p;
    int span, pos[SPO_KMER];
    spo_matrix m80, m40u;
    const unsigned char *tres;
    const uint64_t *toff;
    uint32_t T;
    /* postings sorted by (kmer, id, pos) */
    uint64_t nPost;
    uint32_t *pk, *pid;
    uint16_t *ppos;
    uint32_t *offsets; /* direct CSR offsets [NK+1], built only for big indexes */
    row3 **rows;       /* lazily built 3-mer rows */
};

static int parse_pattern(const char *pat, int *pos) {
    int span = (int)strlen(pat), k = 0;
    if (span > SPO_MAXSPAN) return -1;
    for (int i = 0; i < span; i++)
        if (pat[i] == '1') {
            if (k >= SPO_KMER) return -1;
            pos[k++] = i;
        }
    return k == SPO_KMER ? span : -1;
}

typedef struct post_tmp { uint32_t k, id; uint16_t pos; } post_tmp;

static int cmp_kp(const void *a, const void *b) { /* (kmer, pos) within one sequence */
    const post_tmp *x = (const post_tmp *)a, *y = (const post_tmp *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}
static int cmp_kip(const void *a, const void *b) { /* (kmer, id, pos) */
    const post_tmp *x = (const post_tmp *)a, *y = (const post_tmp *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    if (x->id != y->id) return x->id < y->id ? -1 : 1;
    return (x->pos > y->pos) - (x->pos < y->pos);
}

/* IndexBuilder::fillDatabase / IndexTable::addKmerCount / addSequence / sortDBSeqLists
 * (M/prefiltering/IndexBuilder.cpp:53-256, IndexTable.h:130-186,341-394): every spaced window
 * without X whose self score reaches kmerThr; per target each distinct k-mer once at its smallest
 * position; posting lists sorted by (id, pos). */
static void build_index(spo_prefilter *pf) {
    size_t cap = 1024, n = 0;
    post_tmp *all = (post_tmp *)malloc(cap * sizeof(post_tmp));
    post_tmp *tmp = NULL;
    size_t tmpCap = 0;
    for (uint32_t t = 0; t < pf->T; t++) {
        const unsigned char *s = pf->tres + pf->toff[t];
        long len = (long)(pf->toff[t + 1] - pf->toff[t]);
        long nw = len - pf->span + 1;
        if (nw <= 0) continue;
        if ((size_t)nw > tmpCap) { tmpCap = (size_t)nw; tmp = (post_tmp *)realloc(tmp, tmpCap * sizeof(post_tmp)); }
        size_t c = 0;
        for (long w = 0; w < nw; w++) {
            uint32_t k = 0, mul = 1;
            int hasX = 0, self = 0;
            for (int i = 0; i < SPO_KMER; i++) {
                unsigned char a = s[w + pf->pos[i]];
                if (a >= SPO_X) { hasX = 1; break; }
                /* IndexBuilder.cpp:8-19 truncates the diagonal score to char */
                self += (signed char)pf->m80.sub[a][a];
                k += a * mul;
                mul *= 20;
            }
            if (hasX) continue;
            /* an exact-matching search indexes every k-mer: localKmerThr = 0 (Prefiltering.cpp:509-512) */
            if (!pf->p.exactKmer && pf->p.kmerThr > 0 && self < pf->p.kmerThr) continue;
            tmp[c].k = k; tmp[c].id = t; tmp[c].pos = (uint16_t)w; c++;
        }
        if (c > 1) qsort(tmp, c, sizeof(post_tmp), cmp_kp);
        uint32_t prev = 0xFFFFFFFFu;
        for (size_t i = 0; i < c; i++) {
            if (tmp[i].k == prev) continue;
            prev = tmp[i].k;
            if (n == cap) { cap *= 2; all = (post_tmp *)realloc(all, cap * sizeof(post_tmp)); }
            all[n++] = tmp[i];
        }
    }
    free(tmp);
    qsort(all, n, sizeof(post_tmp), cmp_kip);
    pf->nPost = n;
    pf->pk = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    pf->pid = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    pf->ppos = (uint16_t *)malloc((n + 1) * sizeof(uint16_t));
    for (size_t i = 0; i < n; i++) { pf->pk[i] = all[i].k; pf->pid[i] = all[i].id; pf->ppos[i] = all[i].pos; }
    free(all);
    pf->offsets = NULL;
    if (n > (1u << 18)) {
        pf->offsets = (uint32_t *)calloc((size_t)NK + 1, sizeof(uint32_t));
        for (size_t i = 0; i < n; i++) pf->offsets[pf->pk[i] + 1]++;
        for (size_t k = 0; k < NK; k++) pf->offsets[k + 1] += pf->offsets[k];
    }
}

static inline void postings_of(const spo_prefilter *pf, uint32_t k, uint64_t *from, uint64_t *to) {
    if (pf->offsets) { *from = pf->offsets[k]; *to = pf->offsets[k + 1]; return; }
    uint64_t lo = 0, hi = pf->nPost;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (pf->pk[mid] < k) lo = mid + 1; else hi = mid; }
    *from = lo;
    hi = pf->nPost;
    uint64_t l2 = lo;
    while (l2 < hi) { uint64_t mid = (l2 + hi) / 2; if (pf->pk[mid] <= k) l2 = mid + 1; else hi = mid; }
    *to = l2;
}

/* ExtendedSubstitutionMatrix::calcScoreMatrix for k=3 (M/prefiltering/ExtendedSubstitutionMatrix.cpp:20-69):
 * all 8000 3-mers y by score desc, std::stable_sort => ties keep cartesian order (y0 slowest). */
static row3 *get_row3(spo_prefilter *pf, int x) {
    row3 *r;
#ifdef _OPENMP
#pragma omp atomic read
#endif
    r = pf->rows[x];
    if (r) return r;
    r = (row3 *)malloc(sizeof(row3));
    int x0 = x % 20, x1 = (x / 20) % 20, x2 = x / 400;
    short sc[N3];
    int cnt[512];
    memset(cnt, 0, sizeof(cnt));
    for (int c = 0; c < N3; c++) {
        int y0 = c / 400, y1 = (c / 20) % 20, y2 = c % 20;
        short s = (short)(pf->m80.sub[x0][y0] + pf->m80.sub[x1][y1] + pf->m80.sub[x2][y2]);
        sc[c] = s;
        cnt[255 - s]++; /* |s| <= 3*44 */
    }
    int start[512], acc = 0;
    for (int b = 0; b < 512; b++) { start[b] = acc; acc += cnt[b]; }
    for (int c = 0; c < N3; c++) {
        int y0 = c / 400, y1 = (c / 20) % 20, y2 = c % 20;
        int o = start[255 - sc[c]]++;
        r->score[o] = sc[c];
        r->idx[o] = (unsigned short)(y0 + 20 * y1 + 400 * y2);
    }
    row3 *old;
#ifdef _OPENMP
#pragma omp critical(spo_row3)
#endif
    {
        old = pf->rows[x];
        if (!old) pf->rows[x] = r;
    }
    if (old) { free(r); return old; }
    return r;
}

/* KmerGenerator::generateKmerList + calculateArrayProduct, M/prefiltering/KmerGenerator.cpp:107-216 */
int spo_similar_kmers(spo_prefilter *pf, const unsigned char *k6, int thr, uint32_t *out, int maxOut) {
    int x1 = k6[0] + 20 * k6[1] + 400 * k6[2];
    int x2 = k6[3] + 20 * k6[4] + 400 * k6[5];
    const row3 *r1 = get_row3(pf, x1), *r2 = get_row3(pf, x2);
    short best2 = r2->score[0];
    short cutoff1 = (short)(thr - best2);
    int n = 0;
    for (int i = 0; i < N3 && r1->score[i] >= cutoff1; i++) {
        short s1 = r1->score[i];
        short cutoff2 = (short)(thr - s1 - 0);
        for (int j = 0; j < N3 && r2->score[j] >= cutoff2; j++) {
            if (n >= maxOut) return n;
            out[n++] = (uint32_t)r1->idx[i] + (uint32_t)r2->idx[j] * 8000u;
        }
    }
    return n;
}

spo_prefilter *spo_prefilter_create(const unsigned char *tres, const uint64_t *toff, uint32_t T,
                                    const spo_matrix *m80, const spo_matrix *m40u,
                                    const spo_prefilter_params *p) {
    spo_prefilter *pf = (spo_prefilter *)calloc(1, sizeof(spo_prefilter));
    pf->p = *p;
    pf->span = parse_pattern(p->pattern, pf->pos);
    if (pf->span < 0) { free(pf); return NULL; }
    pf->m80 = *m80;
    pf->m40u = *m40u;
    pf->tres = tres;
    pf->toff = toff;
    pf->T = T;
    pf->rows = (row3 **)calloc(N3, sizeof(row3 *));
    build_index(pf);
    return pf;
}

void spo_prefilter_destroy(spo_prefilter *pf) {
    if (!pf) return;
    for (int i = 0; i < N3; i++) free(pf->rows[i]);
    free(pf->rows);
    free(pf->pk); free(pf->pid); free(pf->ppos); free(pf->offsets);
    free(pf);
}

uint64_t spo_prefilter_num_postings(const spo_prefilter *pf) { return pf->nPost; }
void spo_prefilter_postings(const spo_prefilter *pf, uint32_t *kmer, uint32_t *id, uint16_t *pos) {
    memcpy(kmer, pf->pk, pf->nPost * sizeof(uint32_t));
    memcpy(id, pf->pid, pf->nPost * sizeof(uint32_t));
    memcpy(pos, pf->ppos, pf->nPost * sizeof(uint16_t));
}

typedef struct hit_s { uint32_t id; uint16_t diag; } hit_s;
typedef struct ev_s { uint32_t id; uint16_t diag; uint8_t count; } ev_s;


/* UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring
 * (M/prefiltering/UngappedAlignment.cpp:37-51,368-382) with prof[i][a] - bias == sub + corr */
static int ungapped_score(const spo_prefilter *pf, const unsigned char *q, int L, const signed char *corr,
                          uint32_t id, int diagonal) {
    const unsigned char *t = pf->tres + pf->toff[id];
    unsigned int tLen = (unsigned int)(pf->toff[id + 1] - pf->toff[id]);
    unsigned int minDist = (unsigned int)abs(diagonal);
    const unsigned char *qq, *tt;
    unsigned int n;
    int qi;
    if (diagonal >= 0 && minDist < (unsigned int)L) {
        n = tLen < (unsigned int)L - minDist ? tLen : (unsigned int)L - minDist;
        qi = (int)minDist; tt = t;
    } else if (diagonal < 0 && minDist < tLen) {
        n = tLen - minDist < (unsigned int)L ? tLen - minDist : (unsigned int)L;
        qi = 0; tt = t + minDist;
    } else {
        return 0;
    }
    qq = q + qi;
    int max = 0, score = 0;
    for (unsigned int k = 0; k < n; k++) {
        unsigned char a = tt[k];
        int curr = (a < SPO_ALPH) ? pf->m40u.sub[qq[k]][a] + corr[qi + (int)k] : 0;
        score += curr;
        if (score < 0) score = 0;
        if (score > max) max = score;
    }
    return max;
}

/* UngappedAlignment::scoreSingleSequence incl. computeLongScore (UngappedAlignment.cpp:216-233,395-403) */
static int ungapped_score_u16(const spo_prefilter *pf, const unsigned char *q, int L, const signed char *corr,
                              uint32_t id, uint16_t diag) {
    unsigned int tLen = (unsigned int)(pf->toff[id + 1] - pf->toff[id]);
    if (L >= 32768 || tLen >= 32768) {
        int totalMax = 0;
        for (unsigned int d = 1; d <= 1 + tLen / 32768; d++) {
            int real = (int)(-(long)d * 65536 + diag);
            int m = ungapped_score(pf, q, L, corr, id, real);
            if (m > totalMax) totalMax = m;
        }
        for (unsigned int d = 0; d <= (unsigned int)L / 65536; d++) {
            int real = (int)(d * 65536 + diag);
            int m = ungapped_score(pf, q, L, corr, id, real);
            if (m > totalMax) totalMax = m;
        }
        return totalMax;
    }
    return ungapped_score(pf, q, L, corr, id, (short)diag);
}

/* stable counting sort by count descending of the elements with count >= thr
 * (QueryMatcher::radixSortByScoreSize, M/prefiltering/QueryMatcher.cpp:459-484) */
static size_t sort_by_score(const ev_s *in, size_t n, unsigned thr, ev_s *out) {
    size_t cnt[256];
    memset(cnt, 0, sizeof(cnt));
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) if (in[i].count >= thr) { cnt[in[i].count]++; kept++; }
    size_t start[256], acc = 0;
    for (int s = 255; s >= 0; s--) { start[s] = acc; acc += cnt[s]; }
    for (size_t i = 0; i < n; i++) if (in[i].count >= thr) out[start[in[i].count]++] = in[i];
    return kept;
}

static int cmp_hit(const void *a, const void *b) { /* hit_t::compareHitsByScoreAndId, QueryMatcher.h:40-50 */
    const spo_hit *x = (const spo_hit *)a, *y = (const spo_hit *)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* id -> small dense slot for the per-target byte arrays of CacheFriendlyOperations */
typedef struct idmap { uint32_t *keys; uint32_t *vals; uint32_t cap; } idmap;
static void idmap_init(idmap *m, size_t n) {
    uint32_t cap = 16;
    while (cap < 2 * n + 1) cap <<= 1;
    m->cap = cap;
    m->keys = (uint32_t *)malloc(cap * sizeof(uint32_t));
    m->vals = (uint32_t *)calloc(cap, sizeof(uint32_t));
    memset(m->keys, 0xFF, cap * sizeof(uint32_t));
}
static uint32_t *idmap_at(idmap *m, uint32_t id) {
    uint32_t h = (id * 2654435761u) & (m->cap - 1);
    while (m->keys[h] != 0xFFFFFFFFu && m->keys[h] != id) h = (h + 1) & (m->cap - 1);
    if (m->keys[h] == 0xFFFFFFFFu) { m->keys[h] = id; m->vals[h] = 0; }
    return &m->vals[h];
}
static void idmap_free(idmap *m) { free(m->keys); free(m->vals); }

int spo_prefilter_query(spo_prefilter *pf, const unsigned char *q, int L, spo_hit *out, spo_query_stats *st) {
    if (st) memset(st, 0, sizeof(*st));
    const int P = L - pf->span + 1;
    if (P < 1) return 0;
    const unsigned BIN = (unsigned)pf->p.binSize;
    float *cb = (float *)malloc(sizeof(float) * (size_t)L);
    if (pf->p.compBias) spo_comp_bias(&pf->m80, q, L, cb);
    else memset(cb, 0, sizeof(float) * (size_t)L);

    /* ---- QueryMatcher::match, QueryMatcher.cpp:188-307: hit stream in (window, k-mer rank, posting) order.
     * Only windows i < indexTo (= last window index) are counted (quirk A, CacheFriendlyOperations.cpp:43). */
    size_t hcap = 4096, nh = 0;
    hit_s *hits = (hit_s *)malloc(hcap * sizeof(hit_s));
    int kcap = 262144;
    uint32_t *klist = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)kcap);
    const int wLast = P - 1;
    for (int w = 0; w < P; w++) {
        unsigned char k6[SPO_KMER];
        int hasX = 0;
        for (int i = 0; i < SPO_KMER; i++) { k6[i] = q[w + pf->pos[i]]; if (k6[i] >= SPO_X) hasX = 1; }
        if (hasX) continue;
        float b = 0;
        for (int i = 0; i < SPO_KMER; i++) b += cb[w + pf->pos[i]];
        short bias = (short)((b < 0.0) ? b - 0.5 : b + 0.5);
        int thr = pf->p.kmerThr - bias;
        if (thr < 0) thr = 0;
        thr = (short)thr;
        int nk;
        if (pf->p.exactKmer) { /* takeOnlyBestKmer: the window's own k-mer, whatever the threshold (QueryMatcher.cpp:221-227) */
            uint32_t k = 0, mul = 1;
            for (int i = 0; i < SPO_KMER; i++) { k += k6[i] * mul; mul *= 20; }
            klist[0] = k;
            nk = 1;
        } else nk = spo_similar_kmers(pf, k6, thr, klist, kcap);
        while (!pf->p.exactKmer && nk == kcap && kcap < 262144 * 32) { /* KmerGenerator::MAX_KMER_RESULT_SIZE */
            kcap *= 2;
            klist = (uint32_t *)realloc(klist, sizeof(uint32_t) * (size_t)kcap);
            nk = spo_similar_kmers(pf, k6, thr, klist, kcap);
        }
        if (st) st->kmers += (uint64_t)nk;
        for (int j = 0; j < nk; j++) {
            uint64_t from, to;
            postings_of(pf, klist[j], &from, &to);
            if (st) st->dbMatches += to - from;
            if (w >= wLast) continue;
            for (uint64_t z = from; z < to; z++) {
                if (nh == hcap) { hcap *= 2; hits = (hit_s *)realloc(hits, hcap * sizeof(hit_s)); }
                hits[nh].id = pf->pid[z];
                hits[nh].diag = (uint16_t)(w - (int)pf->ppos[z]);
                nh++;
            }
        }
    }
    free(klist);

    /* ---- CacheFriendlyOperations::findDuplicates, CacheFriendlyOperations.cpp:146-235 (computeTotalScore=false) */
    ev_s *ev = (ev_s *)malloc((nh + 1) * sizeof(ev_s));
    ev_s *evbin = (ev_s *)malloc((nh + 1) * sizeof(ev_s));
    size_t ne = 0;
    {
        idmap last;
        idmap_init(&last, nh);
        for (unsigned bin = 0; bin < BIN; bin++) {
            size_t nb = 0;
            for (size_t i = 0; i < nh; i++) {
                if ((hits[i].id & (BIN - 1)) != bin) continue;
                uint32_t *slot = idmap_at(&last, hits[i].id);
                unsigned char cur = (unsigned char)hits[i].diag;
                if (cur == (unsigned char)*slot) { evbin[nb].id = hits[i].id; evbin[nb].diag = hits[i].diag; evbin[nb].count = 0; nb++; }
                *slot = cur;
            }
            for (size_t i = nb; i-- > 0;) *idmap_at(&last, evbin[i].id) = (unsigned char)((unsigned char)evbin[i].diag + 1);
            for (size_t i = 0; i < nb; i++) {
                uint32_t *slot = idmap_at(&last, evbin[i].id);
                if ((unsigned char)*slot != (unsigned char)evbin[i].diag) ev[ne++] = evbin[i];
                *slot = (unsigned char)evbin[i].diag;
            }
            /* the reference zeroes the byte array between bins; ids of different bins never collide */
        }
        idmap_free(&last);
    }
    free(evbin);
    free(hits);
    if (st) st->events = ne;

    /* ---- UngappedAlignment::processQuery / createProfile, UngappedAlignment.cpp:27-35,295-355 */
    signed char *corr = (signed char *)malloc((size_t)L);
    int minSub = 0, aaBias = 0;
    for (int i = 0; i < pf->m40u.alph; i++)
        for (int j = 0; j < pf->m40u.alph; j++) if (pf->m40u.sub[i][j] < minSub) minSub = pf->m40u.sub[i][j];
    for (int i = 0; i < L; i++) {
        float a = cb[i];
        a = (float)((a < 0.0) ? a / 4 - 0.5 : a / 4 + 0.5);
        corr[i] = (signed char)a;
        if (corr[i] < aaBias) aaBias = corr[i];
    }
    if (aaBias > 0) aaBias = 0;
    const int bias = abs(minSub) + abs(aaBias);
    const int cap = 255 - bias;
    for (size_t i = 0; i < ne; i++) {
        int s = ungapped_score_u16(pf, q, L, corr, ev[i].id, ev[i].diag);
        ev[i].count = (uint8_t)(s < cap ? s : cap);
    }

    /* ---- selection, QueryMatcher.cpp:102-164 */
    ev_s *A = (ev_s *)malloc((ne + 1) * sizeof(ev_s));
    ev_s *B = (ev_s *)malloc((ne + 1) * sizeof(ev_s));
    size_t nA = sort_by_score(ev, ne, (unsigned)pf->p.minDiagScore, A);
    size_t nB = 0;
    {   /* keepMaxScoreElementOnly, CacheFriendlyOperations.cpp:311-341 */
        idmap mx;
        idmap_init(&mx, nA);
        for (unsigned bin = 0; bin < BIN; bin++) {
            for (size_t i = 0; i < nA; i++) {
                if ((A[i].id & (BIN - 1)) != bin) continue;
                uint32_t *slot = idmap_at(&mx, A[i].id);
                if (A[i].count > *slot) *slot = A[i].count;
            }
            for (size_t i = 0; i < nA; i++) {
                if ((A[i].id & (BIN - 1)) != bin) continue;
                uint32_t *slot = idmap_at(&mx, A[i].id);
                if (*slot == A[i].count) { B[nB++] = A[i]; *slot = 0; }
            }
        }
        idmap_free(&mx);
    }
    size_t maxHits = (size_t)pf->p.maxSeqs < pf->T ? (size_t)pf->p.maxSeqs : pf->T;
    /* computeScoreThreshold, QueryMatcher.h:199-209 */
    size_t hist[256];
    memset(hist, 0, sizeof(hist));
    for (size_t i = 0; i < nB; i++) hist[B[i].count]++;
    unsigned thrD = 0;
    {
        size_t found = 0;
        for (int s = 255; s >= 1; s--) {
            found += hist[s];
            if (found >= maxHits) { thrD = (unsigned)s; break; }
        }
    }
    if (thrD < (unsigned)pf->p.minDiagScore) thrD = (unsigned)pf->p.minDiagScore;
    ev_s *C = A; /* reuse */
    size_t nC = sort_by_score(B, nB, thrD, C);
    int nres = 0;
    if ((int)thrD >= cap) {
        /* rescoreHits path, QueryMatcher.cpp:149-155,486-505 */
        int maxSelf = 0;
        {   /* align->scoreSingleSequence(query itself, diagonal 0) */
            int s = 0;
            for (int k = 0; k < L; k++) {
                int curr = pf->m40u.sub[q[k]][q[k]] + corr[k];
                s += curr; if (s < 0) s = 0; if (s > maxSelf) maxSelf = s;
            }
        }
        maxSelf -= cap;
        if (maxSelf < 1) maxSelf = 1;
        if (maxSelf > 65535) maxSelf = 65535;
        float fltMax = (float)maxSelf;
        size_t nr = 0;
        for (size_t i = 0; i < nC && C[i].count >= cap; i++) {
            unsigned int ns = (unsigned int)ungapped_score_u16(pf, q, L, corr, C[i].id, C[i].diag);
            ns -= (unsigned int)cap;
            float sc = (float)(ns < 65535u ? ns : 65535u);
            C[i].count = (unsigned char)((sc / fltMax) * (float)255 + 0.5);
            nr++;
        }
        size_t nD = sort_by_score(C, nr, 0, B);
        for (size_t i = 0; i < nD && (size_t)nres < maxHits; i++) {
            unsigned int ns = (unsigned int)cap + (B[i].count * (unsigned)maxSelf / 255);
            out[nres].id = B[i].id; out[nres].score = (int)ns; out[nres].diag = (short)B[i].diag; nres++;
        }
    } else {
        /* getResult<UNGAPPED_DIAGONAL_SCORE>, QueryMatcher.cpp:324-381 */
        for (size_t i = 0; i < nC && (size_t)nres < maxHits; i++) {
            if (C[i].count < thrD) continue;
            int sc = C[i].count;
            if (sc >= cap) sc = ungapped_score_u16(pf, q, L, corr, C[i].id, C[i].diag);
            out[nres].id = C[i].id; out[nres].score = sc; out[nres].diag = (short)C[i].diag; nres++;
        }
    }
    if (nres > 1) qsort(out, (size_t)nres, sizeof(spo_hit), cmp_hit);
    free(A); free(B); free(ev); free(corr); free(cb);
    return nres;
}

void spo_prefilter_batch(spo_prefilter *pf, const unsigned char *qres, const uint64_t *qoff, uint32_t Q,
                         uint64_t *outOff, spo_hit *out, int threads) {
    int ms = pf->p.maxSeqs;
    int *cnt = (int *)calloc(Q + 1, sizeof(int));
    spo_hit *tmp = (spo_hit *)malloc(sizeof(spo_hit) * (size_t)Q * (size_t)ms);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (long i = 0; i < (long)Q; i++)
        cnt[i] = spo_prefilter_query(pf, qres + qoff[i], (int)(qoff[i + 1] - qoff[i]), tmp + (size_t)i * ms, NULL);
    outOff[0] = 0;
    for (uint32_t i = 0; i < Q; i++) {
        memcpy(out + outOff[i], tmp + (size_t)i * ms, sizeof(spo_hit) * (size_t)cnt[i]);
        outOff[i + 1] = outOff[i] + (uint64_t)cnt[i];
    }
    free(tmp);
    free(cnt);
}

/* ------------------------------------------------------------------------------------------
 * E-value (M/alignment/EvalueComputation.h:18-40, lib/alp/sls_pvalues.cpp:366-541)
 * ---------------------------------------------------------------------------------------- */
double spo_evalue(const spo_gumbel *g, double y, double qLen, double dbRes) {
    const double const_val = 1.0 / sqrt(2.0 * 3.1415926535897932384626433832795);
    double vi_thr = 2.0 * g->alpha_I / g->lambda, vj_thr = 2.0 * g->alpha_J / g->lambda;
    double c_thr = 2.0 * g->sigma / g->lambda;
    double m_li_y = dbRes - (g->a_I * y + g->b_I);
    double vi_y = g->alpha_I * y + g->beta_I; if (vi_y < vi_thr) vi_y = vi_thr;
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = (sqrt_vi_y == 0.0) ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = 0.5 * erfc(-sqrt(0.5) * m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = qLen - (g->a_J * y + g->b_J);
    double vj_y = g->alpha_J * y + g->beta_J; if (vj_y < vj_thr) vj_y = vj_thr;
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = (sqrt_vj_y == 0.0) ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = 0.5 * erfc(-sqrt(0.5) * n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    double c_y = g->sigma * y + g->tau; if (c_y < c_thr) c_y = c_thr;
    double area = p1 * p2 + c_y * (P_m_F * P_n_F);
    double epa = g->K * exp(-g->lambda * y);
    return epa * area;
}

double spo_bitscore(const spo_gumbel *g, double score) { return (g->lambda * score - log(g->K)) / log(2.0); }

/* ------------------------------------------------------------------------------------------
 * striped Smith-Waterman, emulated lane by lane
 * ---------------------------------------------------------------------------------------- */
typedef struct sw_end { int score, ref, read; } sw_end;

/* sw_sse2_byte, M/alignment/StripedSmithWaterman.cpp:301-516.  sc[a*qLen + i] = unbiased profile
 * score of query row i against residue a; layout row = lane*segLen + seg. */
static sw_end sw_byte(const unsigned char *db, int dir, int dbLen, int qLen, int go, int ge,
                      const int *sc, int terminate, int bias, int LANES) {
    int segLen = (qLen + LANES - 1) / LANES;
    int n = segLen * LANES;
    int *Hs = (int *)calloc((size_t)n * 4, sizeof(int));
    int *Hl = Hs + n, *E = Hl + n, *Hmax = E + n;
    int *vF = (int *)malloc(sizeof(int) * (size_t)LANES * 3);
    int *vH = vF + LANES, *vT = vH + LANES;
    int max = 0, end_query = qLen - 1, end_db = -1;
    int *vMaxScore = (int *)calloc((size_t)LANES * 2, sizeof(int));
    int *vMaxMark = vMaxScore + LANES;
    int begin = 0, end = dbLen, step = 1;
    if (dir == 1) { begin = dbLen - 1; end = -1; step = -1; }
#define IDX(seg, lane) ((seg) * LANES + (lane))
#define PROF(a, seg, lane) (((lane) * segLen + (seg)) < qLen ? sc[(a) * qLen + (lane) * segLen + (seg)] + bias : bias)
    for (int i = begin; i != end; i += step) {
        int maxCol = 0;
        for (int l = 0; l < LANES; l++) vF[l] = 0;
        /* vH = pvHStore[segLen-1] shifted by one lane */
        vH[0] = 0;
        for (int l = 1; l < LANES; l++) vH[l] = Hs[IDX(segLen - 1, l - 1)];
        int *t = Hl; Hl = Hs; Hs = t;
        int a = db[i];
        for (int j = 0; j < segLen; j++) {
            for (int l = 0; l < LANES; l++) {
                int h = vH[l] + PROF(a, j, l); if (h > 255) h = 255;
                h -= bias; if (h < 0) h = 0;
                int e = E[IDX(j, l)];
                if (e > h) h = e;
                if (vF[l] > h) h = vF[l];
                if (h > maxCol) maxCol = h;
                Hs[IDX(j, l)] = h;
                h -= go; if (h < 0) h = 0;
                e -= ge; if (e < 0) e = 0;
                if (h > e) e = h;
                E[IDX(j, l)] = e;
                int f = vF[l] - ge; if (f < 0) f = 0;
                if (h > f) f = h;
                vF[l] = f;
                vH[l] = Hl[IDX(j, l)];
            }
        }
        /* lazy F */
        {
            int j = 0;
            for (int l = LANES - 1; l > 0; l--) vF[l] = vF[l - 1];
            vF[0] = 0;
            for (;;) {
                int any = 0;
                for (int l = 0; l < LANES; l++) {
                    int h = Hs[IDX(j, l)] - go; if (h < 0) h = 0;
                    int d = vF[l] - h; if (d < 0) d = 0;
                    if (d != 0) any = 1;
                }
                if (!any) break;
                for (int l = 0; l < LANES; l++) {
                    int h = Hs[IDX(j, l)];
                    if (vF[l] > h) h = vF[l];
                    if (h > maxCol) maxCol = h;
                    Hs[IDX(j, l)] = h;
                    int f = vF[l] - ge; if (f < 0) f = 0;
                    vF[l] = f;
                }
                j++;
                if (j >= segLen) {
                    j = 0;
                    for (int l = LANES - 1; l > 0; l--) vF[l] = vF[l - 1];
                    vF[0] = 0;
                }
            }
        }
        /* the per-lane running maximum only matters through its horizontal max */
        (void)vT; (void)vMaxScore; (void)vMaxMark;
        if (maxCol > max) {
            max = maxCol;
            if (max + bias >= 255) break;
            end_db = i;
            memcpy(Hmax, Hs, sizeof(int) * (size_t)n);
        }
        if (terminate > 0 && maxCol == terminate) break;
    }
    for (int i = 0; i < n; i++) {
        if (Hmax[i] == max) {
            int temp = i / LANES + i % LANES * segLen;
            if (temp < end_query) end_query = temp;
        }
    }
#undef IDX
#undef PROF
    sw_end r;
    r.score = (max + bias >= 255) ? 255 : max;
    r.ref = end_db;
    r.read = end_query;
    free(vMaxScore);
    free(vF);
    free(Hs < Hl ? Hs : Hl);
    return r;
}

/* sw_sse2_word, M/alignment/StripedSmithWaterman.cpp:519-687 (signed 16-bit lanes, no bias) */
static sw_end sw_word(const unsigned char *db, int dir, int dbLen, int qLen, int go, int ge,
                      const int *sc, int terminate, int LANES) {
    int segLen = (qLen + LANES - 1) / LANES;
    int n = segLen * LANES;
    int *base = (int *)calloc((size_t)n * 4, sizeof(int));
    int *Hs = base, *Hl = Hs + n, *E = Hl + n, *Hmax = E + n;
    int *vF = (int *)malloc(sizeof(int) * (size_t)LANES * 2);
    int *vH = vF + LANES;
    int max = 0, end_read = qLen - 1, end_ref = 0;
    int begin = 0, end = dbLen, step = 1;
    if (dir == 1) { begin = dbLen - 1; end = -1; step = -1; }
#define IDX(seg, lane) ((seg) * LANES + (lane))
#define PROF(a, seg, lane) (((lane) * segLen + (seg)) < qLen ? sc[(a) * qLen + (lane) * segLen + (seg)] : 0)
    for (int i = begin; i != end; i += step) {
        int maxCol = 0;
        for (int l = 0; l < LANES; l++) vF[l] = 0;
        vH[0] = 0;
        for (int l = 1; l < LANES; l++) vH[l] = Hs[IDX(segLen - 1, l - 1)];
        int *t = Hl; Hl = Hs; Hs = t;
        int a = db[i];
        for (int j = 0; j < segLen; j++) {
            for (int l = 0; l < LANES; l++) {
                int h = vH[l] + PROF(a, j, l);
                if (h > 32767) h = 32767;
                if (h < -32768) h = -32768;
                int e = E[IDX(j, l)];
                if (e > h) h = e;
                if (vF[l] > h) h = vF[l];
                if (h > maxCol) maxCol = h;
                Hs[IDX(j, l)] = h;
                h -= go; if (h < 0) h = 0;
                e -= ge; if (e < 0) e = 0;
                if (h > e) e = h;
                E[IDX(j, l)] = e;
                int f = vF[l] - ge; if (f < 0) f = 0;
                if (h > f) f = h;
                vF[l] = f;
                vH[l] = Hl[IDX(j, l)];
            }
        }
        for (int k = 0; k < LANES; k++) {
            for (int l = LANES - 1; l > 0; l--) vF[l] = vF[l - 1];
            vF[0] = 0;
            int done = 0;
            for (int j = 0; j < segLen; j++) {
                int any = 0;
                for (int l = 0; l < LANES; l++) {
                    int h = Hs[IDX(j, l)];
                    if (vF[l] > h) h = vF[l];
                    if (h > maxCol) maxCol = h;
                    Hs[IDX(j, l)] = h;
                    h -= go; if (h < 0) h = 0;
                    int f = vF[l] - ge; if (f < 0) f = 0;
                    vF[l] = f;
                    if (f > h) any = 1;
                }
                if (!any) { done = 1; break; }
            }
            if (done) break;
        }
        if (maxCol > max) {
            max = maxCol;
            end_ref = i;
            memcpy(Hmax, Hs, sizeof(int) * (size_t)n);
        }
        if (terminate > 0 && maxCol == terminate) break;
    }
    for (int i = 0; i < n; i++) {
        if (Hmax[i] == max) {
            int temp = i / LANES + i % LANES * segLen;
            if (temp < end_read) end_read = temp;
        }
    }
#undef IDX
#undef PROF
    sw_end r;
    r.score = max; r.ref = end_ref; r.read = end_read;
    free(vF);
    free(base);
    return r;
}

/* banded_sw, M/alignment/StripedSmithWaterman.cpp:781-984.  Returns the per-column op string
 * (M/I/D) in forward order, or NULL on the reference's "Trace back error" path. */
static char *banded_trace(const unsigned char *db, const unsigned char *qs, const signed char *cb,
                          int dbLen, int qLen, int score, int go, int ge, int bw, const spo_matrix *m) {
    int64_t s1 = 8, s2 = 1024;
    int32_t *h_b = (int32_t *)calloc((size_t)s1, sizeof(int32_t));
    int32_t *e_b = (int32_t *)calloc((size_t)s1, sizeof(int32_t));
    int32_t *h_c = (int32_t *)calloc((size_t)s1, sizeof(int32_t));
    int8_t *direction = (int8_t *)calloc((size_t)s2, 1);
    int64_t width, width_d;
    int max = 0;
    int8_t *dl = direction;
#define SET_U(u, w, i, j) { int x = (i) - (w); x = x > 0 ? x : 0; (u) = (j) - x + 1; }
#define SET_D(u, w, i, j, p) { int x = (i) - (w); x = x > 0 ? x : 0; x = (j) - x; (u) = x * 3 + (p); }
    do {
        width = (int64_t)bw * 2 + 3; width_d = (int64_t)bw * 2 + 1;
        if (width >= s1) {
            int64_t ns = s1;
            while (width >= ns) ns *= 2;
            h_b = (int32_t *)realloc(h_b, (size_t)ns * sizeof(int32_t));
            e_b = (int32_t *)realloc(e_b, (size_t)ns * sizeof(int32_t));
            h_c = (int32_t *)realloc(h_c, (size_t)ns * sizeof(int32_t));
            memset(h_b + s1, 0, (size_t)(ns - s1) * sizeof(int32_t));
            memset(e_b + s1, 0, (size_t)(ns - s1) * sizeof(int32_t));
            memset(h_c + s1, 0, (size_t)(ns - s1) * sizeof(int32_t));
            s1 = ns;
        }
        int64_t target = width_d * qLen * 3;
        if (target >= s2) {
            int64_t ns = s2;
            while (target >= ns) ns *= 2;
            direction = (int8_t *)realloc(direction, (size_t)ns);
            memset(direction + s2, 0, (size_t)(ns - s2));
            s2 = ns;
        }
        for (int j = 1; j < width - 1; j++) h_b[j] = 0;
        for (int i = 0; i < qLen; i++) {
            int beg = 0, end = dbLen - 1, u = 0, edge, f, j;
            j = i - bw; beg = beg > j ? beg : j;
            j = i + bw; end = end < j ? end : j;
            edge = end + 1 < width - 1 ? end + 1 : (int)(width - 1);
            f = h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
            dl = direction + width_d * i * 3;
            for (j = beg; j <= end; j++) {
                int b, e1, f1, d, de, df, dh, e, t1, t2;
                SET_U(u, bw, i, j); SET_U(e, bw, i - 1, j);
                SET_U(b, bw, i, j - 1); SET_U(d, bw, i - 1, j - 1);
                SET_D(de, bw, i, j, 0); SET_D(df, bw, i, j, 1); SET_D(dh, bw, i, j, 2);
                t1 = i == 0 ? -go : h_b[e] - go;
                t2 = i == 0 ? -ge : e_b[e] - ge;
                e_b[u] = t1 > t2 ? t1 : t2;
                dl[de] = t1 > t2 ? 3 : 2;
                t1 = h_c[b] - go;
                t2 = f - ge;
                f = t1 > t2 ? t1 : t2;
                dl[df] = t1 > t2 ? 5 : 4;
                e1 = e_b[u] > 0 ? e_b[u] : 0;
                f1 = f > 0 ? f : 0;
                t1 = e1 > f1 ? e1 : f1;
                t2 = h_b[d] + m->sub[qs[i]][db[j]] + cb[i];
                h_c[u] = t1 > t2 ? t1 : t2;
                if (h_c[u] > max) max = h_c[u];
                if (t1 <= t2) dl[dh] = 1;
                else dl[dh] = e1 > f1 ? dl[de] : dl[df];
            }
            for (j = 1; j <= u; j++) h_b[j] = h_c[j];
        }
        bw *= 2;
    } while (max < score);
    bw /= 2;
    int i = qLen - 1, j = dbLen - 1, st = 2;
    size_t cap = (size_t)qLen + (size_t)dbLen + 4, n = 0;
    char *ops = (char *)malloc(cap + 1);
    int ok = 1;
    while (i > 0 || j > 0) {
        int t1;
        SET_D(t1, bw, i, j, st);
        char op;
        switch (dl[t1]) {
            case 1: --i; --j; st = 2; dl -= width_d * 3; op = 'M'; break;
            case 2: --i; st = 0; dl -= width_d * 3; op = 'I'; break;
            case 3: --i; st = 2; dl -= width_d * 3; op = 'I'; break;
            case 4: --j; st = 1; op = 'D'; break;
            case 5: --j; st = 2; op = 'D'; break;
            default: ok = 0; op = 0; break;
        }
        if (!ok || n >= cap) { ok = 0; break; }
        ops[n++] = op;
    }
#undef SET_U
#undef SET_D
    free(h_b); free(e_b); free(h_c); free(direction);
    if (!ok) { free(ops); return NULL; }
    ops[n++] = 'M'; /* final forced M (StripedSmithWaterman.cpp:943-960) */
    for (size_t a = 0, b = n - 1; a < b; a++, b--) { char t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
    ops[n] = '\0';
    return ops;
}

/* profile scores and SW composition bias, SmithWaterman::ssw_init StripedSmithWaterman.cpp:689-779 */
static int build_profile(const spo_matrix *m, int compBias, const unsigned char *q, int L, signed char *cb, int *sc) {
    int cbMin = 0;
    if (compBias) {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)L);
        spo_comp_bias(m, q, L, tmp);
        for (int i = 0; i < L; i++) {
            cb[i] = (signed char)((tmp[i] < 0.0) ? tmp[i] - 0.5 : tmp[i] + 0.5);
            if (cb[i] < cbMin) cbMin = cb[i];
        }
        free(tmp);
    } else {
        memset(cb, 0, (size_t)L);
    }
    int bias = 0;
    for (int i = 0; i < m->alph; i++)
        for (int j = 0; j < m->alph; j++) if ((signed char)m->sub[i][j] < bias) bias = (signed char)m->sub[i][j];
    bias = abs(bias) + abs(cbMin);
    for (int a = 0; a < m->alph; a++)
        for (int i = 0; i < L; i++) sc[a * L + i] = (signed char)m->sub[a][q[i]] + cb[i];
    return bias;
}

void spo_sw_forward(const spo_matrix *m, const spo_align_params *p, const unsigned char *q, int L,
                    const unsigned char *t, int tLen, int *score, int *qEnd, int *tEnd, int *usedWord) {
    signed char *cb = (signed char *)malloc((size_t)L);
    int *sc = (int *)malloc(sizeof(int) * (size_t)L * SPO_ALPH);
    int bias = build_profile(m, p->compBias, q, L, cb, sc);
    sw_end b = sw_byte(t, 0, tLen, L, p->gapOpen, p->gapExtend, sc, -1, bias, p->byteLanes);
    int word = 0;
    if (b.score == 255) { b = sw_word(t, 0, tLen, L, p->gapOpen, p->gapExtend, sc, -1, p->wordLanes); word = 1; }
    *score = b.score; *qEnd = b.read; *tEnd = b.ref; *usedWord = word;
    free(cb); free(sc);
}

static void rle_cigar(const char *bt, char *out, size_t outCap) { /* Matcher::compressAlignment, Matcher.cpp:211-230 */
    size_t o = 0;
    char state = 'M';
    size_t counter = 0;
    for (size_t i = 0; bt[i]; i++) {
        if (bt[i] != state) {
            o += (size_t)snprintf(out + o, outCap - o, "%zu%c", counter, state);
            state = bt[i];
            counter = 1;
        } else counter++;
        if (o + 24 > outCap) break;
    }
    snprintf(out + o, outCap - o, "%zu%c", counter, state);
}

/* Util::canBeCovered, M/commons/Util.cpp:533-550 (float arithmetic, as there); case labels by value (Parameters.h:246-251:
 * BIDIRECTIONAL 0, TARGET 1, QUERY 2, LENGTH_QUERY 3, LENGTH_TARGET 4, LENGTH_SHORTER 5) */
int spo_can_be_covered(float covThr, int covMode, float qLen, float tLen) {
    switch (covMode) {
    case 0: return (qLen / tLen >= covThr) && (tLen / qLen >= covThr);
    case 2: return (tLen / qLen) >= covThr;
    case 1: return (qLen / tLen) >= covThr;
    case 3: return ((tLen / qLen) >= covThr) && (tLen / qLen) <= 1.0;
    case 4: return ((qLen / tLen) >= covThr) && (qLen / tLen) <= 1.0;
    case 5: return ((tLen < qLen ? tLen : qLen) / (tLen > qLen ? tLen : qLen)) >= covThr;
    default: return 1;
    }
}

/* Util::hasCoverage, M/commons/Util.cpp:552-567 */
static int spo_has_coverage(float covThr, int covMode, float qCov, float tCov) {
    switch (covMode) {
    case 0: return (qCov >= covThr) && (tCov >= covThr);
    case 2: return qCov >= covThr;
    case 1: return tCov >= covThr;
    default: return 1;
    }
}

/* SmithWaterman::computeCov, StripedSmithWaterman.cpp:1055-1057 */
static float spo_compute_cov(unsigned start, unsigned end, unsigned len) {
    unsigned hi = start > end ? start : end, lo = start < end ? start : end;
    return ((len < hi ? len : hi) - lo + 1) / (float)len;
}

void spo_align_pair(const spo_matrix *m, const spo_align_params *p, const unsigned char *q, int L,
                    const unsigned char *t, int tLen, spo_aln *out) {
    memset(out, 0, sizeof(*out));
    out->qStart = out->tStart = -1;
    /* the pair is skipped when its lengths cannot reach the coverage (Alignment.cpp:378-381) */
    if (!spo_can_be_covered(p->covThr, p->covMode, (float)L, (float)tLen)) return;
    signed char *cb = (signed char *)malloc((size_t)L);
    int *sc = (int *)malloc(sizeof(int) * (size_t)L * SPO_ALPH);
    int bias = build_profile(m, p->compBias, q, L, cb, sc);
    const int go = p->gapOpen, ge = p->gapExtend;
    sw_end b = sw_byte(t, 0, tLen, L, go, ge, sc, -1, bias, p->byteLanes);
    int word = 0;
    if (b.score == 255) { b = sw_word(t, 0, tLen, L, go, ge, sc, -1, p->wordLanes); word = 1; }
    out->score = b.score; out->qEnd = b.read; out->tEnd = b.ref;
    if (b.ref == -1) { free(cb); free(sc); return; } /* no residue could be aligned */
    out->eval = spo_evalue(&p->g, (double)b.score, (double)L, (double)p->dbResCount);
    out->bits = (int)(spo_bitscore(&p->g, (double)b.score) + 0.5);
    if (p->swMode == 2) {
        /* SCORE_ONLY: ssw_align returns after the forward pass (:195-197); no start positions (-1), coverage stays 0
         * (Matcher.cpp:140-147), identity estimated from the score per column over the END positions (:164-169, :205-209) */
        unsigned qa = (unsigned)(b.read > 1 ? b.read : 1), ta = (unsigned)(b.ref > 1 ? b.ref : 1);
        float est = (float)(((unsigned short)b.score / (float)(qa > ta ? qa : ta)) * 0.1656 + 0.1141);
        if (est > 1.0f) est = 1.0f;
        if (est < 0.0f) est = 0.0f;
        out->seqId = est;
        int a = abs(b.read - (-1)), d = abs(b.ref - (-1));   /* computeAlnLength(qStart = -1, qEnd, dbStart = -1, dbEnd) */
        out->alnLen = (a > d ? a : d) + 1;
        out->ok = (out->eval <= p->evalThr && out->seqId >= p->seqIdThr && spo_has_coverage(p->covThr, p->covMode, 0.0f, 0.0f) &&
                   out->alnLen >= p->alnLenThr) ? 1 : 0;
        free(cb); free(sc);
        return;
    }
    if (out->eval > p->evalThr) { free(cb); free(sc); return; } /* ssw_align early return, :191-200 */
    /* reverse pass on q[qEnd..0] against t[tEnd..0] (:202-231) */
    int rl = b.read + 1;
    unsigned char *qr = (unsigned char *)malloc((size_t)rl);
    int *scr = (int *)malloc(sizeof(int) * (size_t)rl * SPO_ALPH);
    for (int i = 0; i < rl; i++) qr[i] = q[b.read - i];
    for (int a = 0; a < m->alph; a++)
        for (int i = 0; i < rl; i++) scr[a * rl + i] = (signed char)m->sub[a][qr[i]] + cb[b.read - i];
    sw_end r = word ? sw_word(t, 1, b.ref + 1, rl, go, ge, scr, b.score, p->wordLanes)
                    : sw_byte(t, 1, b.ref + 1, rl, go, ge, scr, b.score, bias, p->byteLanes);
    free(qr); free(scr);
    if (r.score != b.score) { out->ok = -1; free(cb); free(sc); return; } /* reference aborts here */
    out->tStart = r.ref;
    out->qStart = b.read - r.read;
    int dl = out->tEnd - out->tStart + 1, ql = out->qEnd - out->qStart + 1;
    int bw = abs(dl - ql) + 1;
    /* coverage of the final coordinates (:226-232): below the threshold ssw_align returns without a backtrace and
     * checkCriteria rejects the pair on the same test */
    const float qCov = spo_compute_cov((unsigned)out->qStart, (unsigned)out->qEnd, (unsigned)L);
    const float tCov = spo_compute_cov((unsigned)out->tStart, (unsigned)out->tEnd, (unsigned)tLen);
    if (!spo_has_coverage(p->covThr, p->covMode, qCov, tCov)) { free(cb); free(sc); return; }
    if (p->swMode == 1) {
        /* SCORE_COV: start and end positions, no backtrace (:233-235); identity estimated over the aligned spans */
        unsigned qa = (unsigned)(out->qEnd - out->qStart > 1 ? out->qEnd - out->qStart : 1);
        unsigned ta = (unsigned)(out->tEnd - out->tStart > 1 ? out->tEnd - out->tStart : 1);
        float est = (float)(((unsigned short)b.score / (float)(qa > ta ? qa : ta)) * 0.1656 + 0.1141);
        if (est > 1.0f) est = 1.0f;
        if (est < 0.0f) est = 0.0f;
        out->seqId = est;
        int a = abs(out->qEnd - out->qStart), d = abs(out->tEnd - out->tStart);
        out->alnLen = (a > d ? a : d) + 1;
        out->ok = (out->eval <= p->evalThr && out->seqId >= p->seqIdThr && out->alnLen >= p->alnLenThr) ? 1 : 0;
        free(cb); free(sc);
        return;
    }
    char *bt = banded_trace(t + out->tStart, q + out->qStart, cb + out->qStart, dl, ql, b.score, go, ge, bw, m);
    int ids = 0;
    if (bt) {
        int qi = out->qStart, ti = out->tStart;
        size_t n = strlen(bt);
        for (size_t k = 0; k < n; k++) {
            if (bt[k] == 'M') { if (q[qi] == t[ti]) ids++; qi++; ti++; }
            else if (bt[k] == 'I') qi++;
            else ti++;
        }
        out->alnLen = (int)n;
        rle_cigar(bt, out->cigar, sizeof(out->cigar));
        free(bt);
    } else {
        int a = abs(out->qEnd - out->qStart), d = abs(out->tEnd - out->tStart);
        out->alnLen = (a > d ? a : d) + 1;
    }
    out->nIdent = ids;
    /* Util::computeSeqId, M/commons/Util.cpp:588-598 */
    const int idDen = p->seqIdMode == 1 ? (L < tLen ? L : tLen) : p->seqIdMode == 2 ? (L > tLen ? L : tLen) : out->alnLen;
    out->seqId = (float)ids / (float)idDen;
    /* Alignment::checkCriteria, Alignment.cpp:555-574 (coverage was tested above) */
    out->ok = (out->eval <= p->evalThr && out->seqId >= p->seqIdThr && out->alnLen >= p->alnLenThr) ? 1 : 0;
    free(cb); free(sc);
}

/* batch driver (OpenMP over queries like Alignment.cpp:285-321, schedule(dynamic,5)); rows[k] for pair k of the CSR
 * list pairOff[Q+1] / pairT[]; used by the parity tests at scale and by bench.py's CPU baseline leg */
void spo_align_batch(const spo_matrix *m, const spo_align_params *p, const unsigned char *qres, const uint64_t *qoff,
                     uint32_t Q, const unsigned char *tres, const uint64_t *toff, const uint64_t *pairOff,
                     const uint32_t *pairT, spo_aln_row *rows, int threads) {
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 5) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (long i = 0; i < (long)Q; i++) {
        spo_aln a;
        for (uint64_t k = pairOff[i]; k < pairOff[i + 1]; k++) {
            uint32_t t = pairT[k];
            spo_align_pair(m, p, qres + qoff[i], (int)(qoff[i + 1] - qoff[i]), tres + toff[t], (int)(toff[t + 1] - toff[t]), &a);
            spo_aln_row *r = rows + k;
            r->ok = a.ok; r->score = a.score; r->bits = a.bits; r->eval = a.eval;
            r->qStart = a.qStart; r->qEnd = a.qEnd; r->tStart = a.tStart; r->tEnd = a.tEnd;
            r->nIdent = a.nIdent; r->alnLen = a.alnLen;
        }
    }
}

/* Matcher::resultToBuffer + Util::fastSeqIdToBuffer (M/alignment/Matcher.cpp:323-370, M/commons/Util.cpp:278-307) */
int spo_format_row(char *buf, uint32_t tKey, const spo_aln *a, int qLen, int tLen) {
    char sid[16];
    float s = a->seqId;
    if (s == 1.0) {
        strcpy(sid, "1.00");
    } else {
        char *w = sid;
        *w++ = '0'; *w++ = '.';
        if (s < 0.10) *w++ = '0';
        if (s < 0.01) *w++ = '0';
        sprintf(w, "%d", (int)(s * 1000));
    }
    return sprintf(buf, "%u\t%d\t%s\t%.3E\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n", tKey, a->bits, sid, a->eval,
                   a->qStart, a->qEnd, qLen, a->tStart, a->tEnd, tLen, a->cigar);
}

/* the same row without the backtrace column: `--alignment-mode 3` without `-a` (addBacktrace == false, Matcher.cpp:355-367) */
int spo_format_row_nobt(char *buf, uint32_t tKey, const spo_aln *a, int qLen, int tLen) {
    int n = spo_format_row(buf, tKey, a, qLen, tLen);
    while (n > 0 && buf[n - 1] != '\t') n--;
    buf[n - 1] = '\n';
    buf[n] = 0;
    return n;
}

/* =========================================================================================================
 * Query preparation: extractorfs --orf-start-mode 1 (ANY_TO_STOP) + translatenucs, canonical table.
 * Follows Orf::setSequence / Orf::findAll / Orf::findForward (M/commons/Orf.cpp:130-159,172-191,230-352),
 * the caller's filter and coordinate flip (M/util/extractorfs.cpp:76-99) and TranslateNucl::translate
 * (M/commons/TranslateNucl.h:488-503) on ASCII nucleotides.
 * ========================================================================================================= */
#define SPO_CHAR_MAX 127

static char spo_complement(char c) {
    /* Orf::iupacReverseComplementTable, M/commons/Orf.cpp:48-52 */
    static const char *tab =
        "................................................................"
        ".TVGH..CD..M.KN...YSAABW.R.......tvgh..cd..m.kn...ysaabw.r......"
        "................................................................"
        "................................................................";
    return tab[(unsigned char)c];
}

static int spo_is_stop(const char *codon) {
    /* TranslateNucl::getStopCodons of the canonical table: TAA TAG TGA; exact match as isInCodons does */
    return codon[0] == 'T' && ((codon[1] == 'A' && (codon[2] == 'A' || codon[2] == 'G')) || (codon[1] == 'G' && codon[2] == 'A'));
}

static int spo_nt_index(char c) {   /* T C A G -> 0..3 (ncbieaa order), -1 ambiguous */
    switch (c) {
        case 'T': case 't': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'A': case 'a': return 2;
        case 'G': case 'g': return 3;
        default: return -1;
    }
}

char spo_translate_codon(const char *codon) {
    /* the NCBI automaton of TranslateNucl returns the residue shared by all expansions of an ambiguous codon, else X
     * (restricted here to the ambiguity N: any base) */
    static const char *ncbieaa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        const int x = spo_nt_index(codon[k]);
        lo[k] = x < 0 ? 0 : x;
        hi[k] = x < 0 ? 3 : x;
    }
    char res = 0;
    for (int a = lo[0]; a <= hi[0]; a++)
        for (int b = lo[1]; b <= hi[1]; b++)
            for (int c = lo[2]; c <= hi[2]; c++) {
                const char r = ncbieaa[16 * a + 4 * b + c];
                if (res == 0) res = r;
                else if (res != r) return 'X';
            }
    return res;
}

static int spo_find_forward(const char *sequence, size_t sequenceLength, size_t minLength, size_t maxLength, int strand,
                            spo_orf *out, int cap, int n) {
    int isInsideOrf[3] = {1, 1, 1}, hasStartCodon[3] = {0, 0, 0};
    size_t countLength[3] = {0, 0, 0}, from[3] = {0, 1, 2};
    char codon[3];
    for (size_t i = 0; i < sequenceLength - 2; i += 3) {
        for (size_t position = i; position < i + 3; position++) {
            for (int k = 0; k < 3; k++)
                codon[k] = sequence[position + k] == SPO_CHAR_MAX ? SPO_CHAR_MAX : (char)(sequence[position + k] & ~0x20);
            const size_t frame = position % 3;
            const int thisIncomplete = codon[0] == SPO_CHAR_MAX || codon[1] == SPO_CHAR_MAX || codon[2] == SPO_CHAR_MAX;
            const char *nx = sequence + position + 3;
            const int isLast = !thisIncomplete && (nx[0] == SPO_CHAR_MAX || nx[1] == SPO_CHAR_MAX || nx[2] == SPO_CHAR_MAX);
            if (!isInsideOrf[frame]) {   /* ANY_TO_STOP: shouldStart */
                isInsideOrf[frame] = 1;
                hasStartCodon[frame] = 1;
                from[frame] = position;
                countLength[frame] = 0;
            }
            const int stop = spo_is_stop(codon);
            if (isInsideOrf[frame] && !stop) countLength[frame]++;
            if (isInsideOrf[frame] && (stop || isLast)) {
                isInsideOrf[frame] = 0;
                if (countLength[frame] == 0 && stop) continue;
                const size_t to = position + ((isLast && !stop) ? 2 : -1);
                if (countLength[frame] > maxLength || countLength[frame] < minLength) continue;
                if (n < cap) {
                    out[n].from = (uint32_t)from[frame];
                    out[n].to = (uint32_t)to;
                    out[n].strand = strand;
                    out[n].incomplete_start = !hasStartCodon[frame];
                    out[n].incomplete_end = !stop;
                }
                n++;
            }
        }
    }
    return n;
}

int spo_extract_frames(const char *seq, int length, int minCodons, int maxCodons, spo_orf *out, int cap, char *aa, int aaCap,
                       int *aaOff) {
    if (length < 3) return 0;
    char *fwd = (char *)malloc((size_t)length + 40), *rev = (char *)malloc((size_t)length + 40);
    for (int i = 0; i < length; i++) fwd[i] = (seq[i] == 'u') ? 't' : seq[i];      /* Orf.cpp:141-144: only 'u' survives the two assignments */
    for (int i = 0; i < length; i++) {
        rev[i] = spo_complement(fwd[length - i - 1]);
        if (rev[i] == '.') rev[i] = 'N';
    }
    for (int i = length; i < length + 40; i++) fwd[i] = rev[i] = SPO_CHAR_MAX;
    int n = spo_find_forward(fwd, (size_t)length, (size_t)minCodons, (size_t)maxCodons, 0, out, cap, 0);
    n = spo_find_forward(rev, (size_t)length, (size_t)minCodons, (size_t)maxCodons, 1, out, cap, n);
    int w = 0;
    for (int k = 0; k < n && k < cap; k++) {
        const char *s = out[k].strand ? rev : fwd;
        aaOff[k] = w;
        for (uint32_t p = out[k].from; p + 2 <= out[k].to; p += 3) {
            if (w < aaCap) aa[w] = spo_translate_codon(s + p);
            w++;
        }
    }
    if (n <= cap) aaOff[n < cap ? n : cap] = w;
    free(fwd);
    free(rev);
    return n;
}

/* =========================================================================================================
 * Nucleotide re-alignment branch of predictmatch (--perform-nucl-aln 1).
 * ========================================================================================================= */
/* NucleotideMatrix::setupLetterMapping (M/commons/NucleotideMatrix.cpp:18-63) over the alphabet order of
 * lib/mmseqs/data/nucleotide.out (A C T G X); scores of that file at bit factor 1, bias 0; ALP parameters of
 * nucleotide.out with gap open 10 / extend 2 as EvalueComputation::init estimates them (EvalueComputation.h:48-154,
 * seed 42; probed from the reference build, identical over runs -- an input like the VTML40 set) */
void spo_nucl_default_params(spo_nucl_params *p, uint64_t dbResCount) {
    memset(p, 0, sizeof(*p));
    p->gapOpen = 10; p->gapExtend = 2; p->dbResCount = dbResCount;
    for (int c = 0; c < 256; c++) {
        int u = toupper(c), n;
        switch (u) {
            case 'A': n = 0; break;
            case 'C': case 'M': case 'Y': case 'H': n = 1; break;
            case 'T': case 'U': case 'W': n = 2; break;
            case 'G': case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': n = 3; break;
            default: n = 4; break;
        }
        p->letter[c] = (unsigned char)n;
    }
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) p->sub[i][j] = (short)((i == j && i < 4) ? 2 : -3);
    spo_gumbel g = {0.63361984397325744, 0.42150719938226583, 0.68666204026925159, 0.18923575026627315,
                    0.71225621255786731, 2.9058945364387716, 0.68666204026925159, 0.18923575026627315,
                    0.71225621255786731, 2.9058945364387716, 0.7125673897037631, 2.8984262849372726};
    p->g = g;
}

/* proteinaln2nucl, per alignment row (lib/mmseqs/src/util/proteinaln2nucl.cpp:113-172): coordinates times three (minus
 * one codon when the amino-acid entry starts with '*'), the protein backtrace walked codon by codon over the raw
 * nucleotide characters: M = three columns scored with the ASCII nucleotide matrix and compared for identity,
 * I / D = gapOpen + ((cnt - 1) * 3) * gapExtend; bit score truncated to int, E-value with the NUCLEOTIDE query length */
void spo_nucl_rescore(const spo_nucl_params *p, const char *q, int qLen, int qStartCodon, const char *t, int tLen, int tStartCodon,
                      int qStart, int qEnd, int tStart, int tEnd, const char *cigar, spo_nucl_row *out) {
    memset(out, 0, sizeof(*out));
    if ((qStartCodon && qStart == 0) || (tStartCodon && tStart == 0)) { out->err = 1; return; }
    out->tStart = tStart * 3 + (tStartCodon ? -3 : 0);
    out->tEnd = tEnd * 3 + 2 + (tStartCodon ? -3 : 0);
    out->tLen = tLen;
    out->qStart = qStart * 3 + (qStartCodon ? -3 : 0);
    out->qEnd = qEnd * 3 + 2 + (qStartCodon ? -3 : 0);
    out->qLen = qLen;
    int qPos = out->qStart, tPos = out->tStart, score = 0;
    unsigned idCnt = 0, alnLen = 0;
    char *w = out->cigar;
    for (const char *c = cigar; *c;) {
        int cnt = 0;
        while (isdigit((unsigned char)*c)) cnt = cnt * 10 + (*c++ - '0');
        char op = *c++;
        if (op == 'M') {
            for (int bt = 0; bt < cnt * 3; bt++) {
                idCnt += (q[qPos] == t[tPos]);
                score += p->sub[p->letter[(unsigned char)q[qPos]]][p->letter[(unsigned char)t[tPos]]];
                tPos++; qPos++;
            }
        } else if (op == 'D') {
            tPos += cnt * 3;
            score -= p->gapOpen + ((cnt - 1) * 3) * p->gapExtend;
        } else if (op == 'I') {
            qPos += cnt * 3;
            score -= p->gapOpen + ((cnt - 1) * 3) * p->gapExtend;
        } else {
            continue;
        }
        alnLen += (unsigned)cnt * 3;
        w += sprintf(w, "%d%c", cnt * 3, op);
    }
    out->rawScore = score;
    out->bits = (int)spo_bitscore(&p->g, score);
    out->eval = spo_evalue(&p->g, score, qLen, (double)p->dbResCount);
    out->idCnt = idCnt; out->alnLen = alnLen;
    out->seqId = (float)idCnt / (float)alnLen;
}

int spo_format_nucl_row(char *buf, uint32_t tKey, const spo_nucl_row *r) {
    spo_aln a;
    memset(&a, 0, sizeof(a));
    a.bits = r->bits; a.eval = r->eval; a.seqId = r->seqId;
    a.qStart = r->qStart; a.qEnd = r->qEnd; a.tStart = r->tStart; a.tEnd = r->tEnd;
    strcpy(a.cigar, r->cigar);
    return spo_format_row(buf, tKey, &a, r->qLen, r->tLen);
}

/* src/util/CombineProtNuclAln.cpp:85 */
double spo_combine_evalue(double protEval, double nuclEval) {
    return ((log(protEval) * 0.5 + log(nuclEval) * 0.5) < log(nuclEval)) ? exp(log(protEval) * 0.5 + log(nuclEval) * 0.5) : nuclEval;
}
