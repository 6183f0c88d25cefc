// capi.cu -- context life cycle, error plumbing, matrices, host-side E-value helpers.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace spb {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace spb

// pinned staging pool: results point into these blocks until spb_free_*_result hands them back
void *spb_ctx::pin_acquire(size_t bytes) {
    if (bytes == 0) bytes = 1;
    int best = -1;
    for (size_t i = 0; i < pins.size(); i++)
        if (!pins[i].used && pins[i].cap >= bytes && (best < 0 || pins[i].cap < pins[(size_t)best].cap)) best = (int)i;
    if (best < 0) {
        // replace the smallest free block instead of accumulating small ones
        for (size_t i = 0; i < pins.size(); i++)
            if (!pins[i].used && (best < 0 || pins[i].cap < pins[(size_t)best].cap)) best = (int)i;
        if (best >= 0) {
            cudaFreeHost(pins[(size_t)best].p);
            pins[(size_t)best] = spb::PinBlock();
        } else {
            pins.push_back(spb::PinBlock());
            best = (int)pins.size() - 1;
        }
        const size_t want = bytes + (bytes >> 2) + 4096;
        if (cudaHostAlloc(&pins[(size_t)best].p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            pins[(size_t)best].p = nullptr;
            spb::set_error("cudaHostAlloc(%zu) failed", want);
            return nullptr;
        }
        pins[(size_t)best].cap = want;
    }
    pins[(size_t)best].used = true;
    return pins[(size_t)best].p;
}

void spb_ctx::pin_release(void *p) {
    if (!p) return;
    for (spb::PinBlock &b : pins)
        if (b.p == p) b.used = false;
}

extern "C" {

const char *spb_last_error(void) { return spb::g_err; }
const char *spb_version(void) { return "spb200 0.2 (sm_100a)"; }
int spb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int spb_ctx_create(int device, spb_ctx **out) {
    SPB_CHECK(out != nullptr, "spb_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        spb::set_error("spb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return 1;
    }
    SPB_CHECK(device >= 0 && device < n, "spb_ctx_create: device %d out of range (0..%d)", device, n - 1);
    SPB_CUDA(cudaSetDevice(device));
    spb_ctx *c = new spb_ctx();
    c->device = device;
    {
        int sm = 0;
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sm > 0) c->n_sm = sm;
    }
    if (const char *e = getenv("SPB_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));   // tuning aid
    SPB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SPB_CUDA(cudaEventCreate(&c->ev0));
    SPB_CUDA(cudaEventCreate(&c->ev1));
    SPB_CUDA(cudaEventCreate(&c->ev2));
    SPB_CUDA(cudaEventCreate(&c->ev3));
    *out = c;
    return 0;
}

// A lane: a second (third, ...) context on the parent's device with its own stream, query batch, results and scratch, reading
// the parent's matrices, k-mer index, alignment targets and nucleotide targets in place.  Batches submitted to different
// lanes from different host threads run concurrently on the GPU: the kernels of one lane fill the tails and the host-sync
// bubbles of the other (10k-spacer batches: 13.1 -> 15.7 -> 17.0 M frames/s with 1 / 2 / 3 lanes, end to end 11.2 -> 16.3).
int spb_ctx_create_lane(spb_ctx *parent, spb_ctx **out) {
    SPB_CHECK(parent && out, "spb_ctx_create_lane: NULL argument");
    *out = nullptr;
    spb_ctx *c = nullptr;
    if (spb_ctx_create(parent->device, &c)) return 1;
    c->is_lane = true;
    c->have_matrices = parent->have_matrices;
    c->hm = parent->hm;
    c->d_mat = parent->d_mat.view();
    c->have_index = parent->have_index;
    c->T = parent->T; c->n_res = parent->n_res; c->span = parent->span;
    memcpy(c->pat, parent->pat, sizeof(c->pat));
    c->d_tres = parent->d_tres.view(); c->d_toff = parent->d_toff.view(); c->d_koff = parent->d_koff.view();
    c->d_post = parent->d_post.view(); c->d_direct = parent->d_direct.view(); c->d_kbits = parent->d_kbits.view();
    c->d_row3 = parent->d_row3.view(); c->d_cnt3 = parent->d_cnt3.view();
    c->n_post = parent->n_post; c->n_kmers_nonempty = parent->n_kmers_nonempty; c->build_ms = parent->build_ms;
    c->n_shards = parent->n_shards; c->shard_rank = parent->shard_rank; c->shard_R = parent->shard_R;
    for (int i = 0; i < SPB_MAX_SHARDS; i++) { c->shard_direct[i] = parent->shard_direct[i]; c->shard_post[i] = parent->shard_post[i]; }
    c->have_aln_targets = parent->have_aln_targets;
    c->aT = parent->aT; c->a_res = parent->a_res; c->a_maxlen = parent->a_maxlen;
    c->d_ares = parent->d_ares.view(); c->d_aoff = parent->d_aoff.view(); c->d_akeys = parent->d_akeys.view();
    c->h_aoff = parent->h_aoff; c->h_akeys = parent->h_akeys;
    c->have_nt_targets = parent->have_nt_targets;
    c->ntT = parent->ntT;
    c->d_tnt = parent->d_tnt.view(); c->d_tntoff = parent->d_tntoff.view(); c->d_tntsc = parent->d_tntsc.view();
    // everything the parent uploaded must have landed before the lane's own stream reads it
    SPB_CUDA(cudaStreamSynchronize(parent->stream));
    *out = c;
    return 0;
}

void spb_ctx_destroy(spb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    spb::DevBuf *bufs[] = {&c->d_mat, &c->d_tres, &c->d_toff, &c->d_koff, &c->d_post, &c->d_direct, &c->d_kbits, &c->d_row3, &c->d_cnt3,
                           &c->d_ares, &c->d_aoff, &c->d_akeys, &c->d_q, &c->d_qoff, &c->d_qoff64, &c->d_qorder, &c->d_hits, &c->d_hits_alt, &c->d_hitoff, &c->d_hit_tkeys, &c->s_tk,
                           &c->d_pairoff, &c->d_pairT, &c->d_pairRank, &c->d_pairQ, &c->d_pairout, &c->d_acc, &c->d_cig, &c->d_rows, &c->d_rowoff, &c->d_st_rows, &c->d_st_cig, &c->d_st_off, &c->d_tnt, &c->d_tntoff, &c->d_tntsc, &c->d_qnt, &c->d_qntoff, &c->d_qntsc, &c->d_nrows,
                           &c->s_a, &c->s_b, &c->s_c, &c->s_d, &c->s_e, &c->s_f, &c->s_g, &c->s_h, &c->s_i, &c->s_j,
                           &c->s_k, &c->s_l, &c->s_m, &c->s_n, &c->s_o, &c->s_p, &c->s_q, &c->s_cub,
                           &c->p_recs, &c->p_ctab, &c->p_ccnt, &c->p_harena, &c->p_meta, &c->p_heavy, &c->p_retry, &c->p_team2};
    for (spb::DevBuf *b : bufs) b->release();
    for (void *m : c->ipc_mapped) cudaIpcCloseMemHandle(m);
    c->ipc_mapped.clear();
    for (spb::PinBlock &b : c->pins)
        if (b.p) cudaFreeHost(b.p);
    c->pins.clear();
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        cudaEventDestroy(c->ev_staged);
        cudaEventDestroy(c->ev_stage_free);
    }
    for (cudaEvent_t ev : c->ev_phase) cudaEventDestroy(ev);
    if (c->aux_stream) {
        cudaStreamSynchronize(c->aux_stream);
        cudaStreamDestroy(c->aux_stream);
        cudaEventDestroy(c->ev_fork);
        cudaEventDestroy(c->ev_join);
    }
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev2);
    cudaEventDestroy(c->ev3);
    cudaStreamDestroy(c->stream);
    delete c;
}

void *spb_ctx_stream(spb_ctx *c) { return c ? (void *)c->stream : nullptr; }
uint64_t spb_ctx_launches(spb_ctx *c) { return c ? c->n_launches : 0; }

int spb_set_matrices(spb_ctx *c, const int16_t *sub_seed, const double *pback_seed, const int16_t *sub_ungapped,
                     const int16_t *sub_align, const double *pback_align) {
    SPB_CHECK(c && sub_seed && pback_seed && sub_ungapped && sub_align && pback_align, "spb_set_matrices: NULL argument");
    SPB_CHECK(!c->is_lane, "spb_set_matrices: a lane context borrows its parent's matrices");
    SPB_CUDA(cudaSetDevice(c->device));
    const int n = SPB_ALPHABET * SPB_ALPHABET;
    memcpy(c->hm.seed, sub_seed, n * sizeof(int16_t));
    memcpy(c->hm.ungapped, sub_ungapped, n * sizeof(int16_t));
    memcpy(c->hm.align, sub_align, n * sizeof(int16_t));
    memcpy(c->hm.pback_seed, pback_seed, SPB_ALPHABET * sizeof(double));
    memcpy(c->hm.pback_align, pback_align, SPB_ALPHABET * sizeof(double));
    for (int i = 0; i < n; i++) {
        SPB_CHECK(sub_seed[i] >= -85 && sub_seed[i] <= 85, "spb_set_matrices: seed score %d out of the supported range", sub_seed[i]);
        SPB_CHECK(sub_ungapped[i] >= -64 && sub_ungapped[i] <= 63 && sub_align[i] >= -64 && sub_align[i] <= 63,
                  "spb_set_matrices: alignment score out of the supported int8 range");
    }
    if (c->d_mat.ensure(sizeof(spb::Matrices))) return 1;
    SPB_CUDA(cudaMemcpyAsync(c->d_mat.p, &c->hm, sizeof(spb::Matrices), cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_matrices = true;
    c->have_index = false;  // 3-mer rows depend on the seed matrix
    return 0;
}

// EvalueComputation::computeEvalue (M/alignment/EvalueComputation.h:36-40) with
// pvalues::get_appr_tail_prob_with_cov_without_errors (lib/alp/sls_pvalues.cpp:366-541)
double spb_evalue(const spb_gumbel *g, double y, double n, double m) {
    const double const_val = 1.0 / sqrt(2.0 * 3.1415926535897932384626433832795);
    double m_li_y = m - (g->a_I * y + g->b_I);
    double vi_y = fmax(2.0 * g->alpha_I / g->lambda, g->alpha_I * y + g->beta_I);
    double sqrt_vi_y = sqrt(vi_y);
    double m_F = sqrt_vi_y == 0.0 ? 1e100 : m_li_y / sqrt_vi_y;
    double P_m_F = 0.5 * erfc(-sqrt(0.5) * m_F);
    double E_m_F = -const_val * exp(-0.5 * m_F * m_F);
    double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
    double n_lj_y = n - (g->a_J * y + g->b_J);
    double vj_y = fmax(2.0 * g->alpha_J / g->lambda, g->alpha_J * y + g->beta_J);
    double sqrt_vj_y = sqrt(vj_y);
    double n_F = sqrt_vj_y == 0.0 ? 1e100 : n_lj_y / sqrt_vj_y;
    double P_n_F = 0.5 * erfc(-sqrt(0.5) * n_F);
    double E_n_F = -const_val * exp(-0.5 * n_F * n_F);
    double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
    double c_y = fmax(2.0 * g->sigma / g->lambda, g->sigma * y + g->tau);
    double area = p1 * p2 + c_y * (P_m_F * P_n_F);
    return g->K * exp(-g->lambda * y) * area;
}

double spb_bitscore(const spb_gumbel *g, double score) { return (g->lambda * score - log(g->K)) / log(2.0); }

}  // extern "C"
