// nuclaln.cu -- nucleotide re-alignment branch of predictmatch (--perform-nucl-aln 1): proteinaln2nucl on the device.
//
// Reference: lib/mmseqs/src/util/proteinaln2nucl.cpp:84-175 (per alignment row; OpenMP over query entries).  Here the
// nucleotide ORFs of targets and queries are resident (ASCII, same indices as the amino-acid entries) and one thread
// re-scores one alignment row along its protein backtrace.  HBM-bound and tiny next to the search: per row one spb_aln
// (56 B), its backtrace (<= ~20 B) and 2 x 3 nucleotides per aligned column are read, one spb_nucl_aln (56 B) written.
#include "common.cuh"

namespace spb {

struct NuclArgs {
    const spb_aln *rows;
    uint64_t n;
    const char *cig;
    uint64_t n_cig;
    const uint8_t *qnt, *tnt;
    const uint64_t *qoff, *toff;
    const uint8_t *qsc, *tsc;   // nullable
    uint32_t Q, T;
    spb_nucl_aln *out;
    spb_nucl_params p;
};

__global__ void __launch_bounds__(128) nucl_rescore_kernel(const NuclArgs A) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const spb_aln r = A.rows[i];
    spb_nucl_aln o;
    memset(&o, 0, sizeof(o));
    if (r.query >= A.Q || r.target >= A.T || (uint64_t)r.cigar_off + r.cigar_len > A.n_cig) {
        o.err = 3;
        A.out[i] = o;
        return;
    }
    const bool qsc = A.qsc && A.qsc[r.query], tsc = A.tsc && A.tsc[r.target];
    if ((qsc && r.q_start == 0) || (tsc && r.t_start == 0)) {   // proteinaln2nucl.cpp:97-100,119-122
        o.err = 1;
        A.out[i] = o;
        return;
    }
    const uint64_t q0 = A.qoff[r.query], t0 = A.toff[r.target];
    const int qLen = (int)(A.qoff[r.query + 1] - q0), tLen = (int)(A.toff[r.target + 1] - t0);
    o.t_start = r.t_start * 3 + (tsc ? -3 : 0);
    o.t_end = r.t_end * 3 + 2 + (tsc ? -3 : 0);
    o.q_start = r.q_start * 3 + (qsc ? -3 : 0);
    o.q_end = r.q_end * 3 + 2 + (qsc ? -3 : 0);
    o.q_len = (uint32_t)qLen;
    o.t_len = (uint32_t)tLen;
    const uint8_t *q = A.qnt + q0, *t = A.tnt + t0;
    const char *ops = A.cig + r.cigar_off;
    int qPos = o.q_start, tPos = o.t_start, score = 0;
    uint32_t ident = 0;
    char prev = 0;
    bool bad = false;
    for (uint32_t k = 0; k < r.cigar_len; k++) {
        const char op = ops[k];
        if (op == 'M') {
            if (qPos + 3 > qLen || tPos + 3 > tLen) { bad = true; break; }
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const uint8_t a = q[qPos + b], c = t[tPos + b];
                ident += (a == c);
                score += A.p.sub[A.p.letter[a] * 5 + A.p.letter[c]];
            }
            qPos += 3;
            tPos += 3;
        } else {
            // a run of cnt gap codons costs gapOpen + ((cnt - 1) * 3) * gapExtend (proteinaln2nucl.cpp:150-159)
            score -= (op == prev) ? 3 * A.p.gap_extend : A.p.gap_open;
            if (op == 'D') tPos += 3; else qPos += 3;
        }
        prev = op;
    }
    if (bad || qPos > qLen || tPos > tLen) {
        o.err = 2;
        A.out[i] = o;
        return;
    }
    o.raw_score = score;
    o.n_ident = ident;
    o.aln_len = r.cigar_len * 3;
    o.seq_id = (float)ident / (float)o.aln_len;
    // EvalueComputation::computeBitScore / computeEvalue with the nucleotide query length (proteinaln2nucl.cpp:166-167)
    o.bits = (int)((A.p.gumbel.lambda * (double)score - log(A.p.gumbel.K)) / log(2.0));
    o.eval = dev_evalue(A.p.gumbel, (double)score, (double)qLen, (double)A.p.db_res_count);
    A.out[i] = o;
}

static int upload_nt(spb_ctx *c, const uint8_t *nt, const uint64_t *off, uint32_t n, const uint8_t *sc, DevBuf &d_nt, DevBuf &d_off,
                     DevBuf &d_sc, const char *who) {
    SPB_CHECK(c && off && (nt || off[n] == 0), "%s: NULL argument", who);
    SPB_CUDA(cudaSetDevice(c->device));
    const uint64_t bytes = off[n];
    if (d_nt.ensure(bytes + 16) || d_off.ensure(((size_t)n + 1) * 8) || d_sc.ensure((size_t)n + 1)) return 1;
    if (bytes) SPB_CUDA(cudaMemcpyAsync(d_nt.p, nt, bytes, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(d_off.p, off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (sc && n) SPB_CUDA(cudaMemcpyAsync(d_sc.p, sc, n, cudaMemcpyHostToDevice, c->stream));
    else SPB_CUDA(cudaMemsetAsync(d_sc.p, 0, (size_t)n + 1, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int rescore(spb_ctx *c, const spb_nucl_params *p, const spb_aln *d_rows, uint64_t n, const char *d_cig, uint64_t n_cig,
                   spb_nucl_aln **out, double *kernel_ms, const char *who) {
    SPB_CHECK(c->have_nt_targets && c->have_nt_queries, "%s: nucleotide ORFs not set (spb_nucl_set_targets / spb_nucl_set_queries)", who);
    *out = nullptr;
    if (c->d_nrows.ensure(std::max<uint64_t>(n, 1) * sizeof(spb_nucl_aln))) return 1;
    spb_nucl_aln *h = (spb_nucl_aln *)c->pin_acquire(std::max<uint64_t>(n, 1) * sizeof(spb_nucl_aln));
    if (!h) return 1;
    cudaStream_t st = c->stream;
    cudaError_t e = cudaEventRecord(c->ev0, st);
    if (n && e == cudaSuccess) {
        NuclArgs A;
        A.rows = d_rows; A.n = n; A.cig = d_cig; A.n_cig = n_cig;
        A.qnt = c->d_qnt.as<uint8_t>(); A.tnt = c->d_tnt.as<uint8_t>();
        A.qoff = c->d_qntoff.as<uint64_t>(); A.toff = c->d_tntoff.as<uint64_t>();
        A.qsc = c->d_qntsc.as<uint8_t>(); A.tsc = c->d_tntsc.as<uint8_t>();
        A.Q = c->ntQ; A.T = c->ntT;
        A.out = c->d_nrows.as<spb_nucl_aln>();
        A.p = *p;
        nucl_rescore_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A);
        e = cudaGetLastError();
        c->n_launches++;
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->ev1, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(h, c->d_nrows.p, n * sizeof(spb_nucl_aln), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        c->pin_release(h);
        set_error("%s: %s", who, cudaGetErrorString(e));
        return 1;
    }
    if (kernel_ms) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        *kernel_ms = ms;
    }
    *out = h;
    return 0;
}

}  // namespace spb

extern "C" {

int spb_nucl_set_targets(spb_ctx *c, const uint8_t *nt, const uint64_t *offsets, uint32_t T, const uint8_t *start_codon) {
    SPB_CHECK(!c || !c->is_lane, "spb_nucl_set_targets: a lane context borrows its parent's index / targets");
    if (spb::upload_nt(c, nt, offsets, T, start_codon, c->d_tnt, c->d_tntoff, c->d_tntsc, "spb_nucl_set_targets")) return 1;
    c->ntT = T;
    c->have_nt_targets = true;
    return 0;
}

int spb_nucl_set_queries(spb_ctx *c, const uint8_t *nt, const uint64_t *offsets, uint32_t Q, const uint8_t *start_codon) {
    if (spb::upload_nt(c, nt, offsets, Q, start_codon, c->d_qnt, c->d_qntoff, c->d_qntsc, "spb_nucl_set_queries")) return 1;
    c->ntQ = Q;
    c->have_nt_queries = true;
    return 0;
}

int spb_nucl_rescore(spb_ctx *c, const spb_nucl_params *p, const spb_aln *rows, uint64_t n_rows, const char *cigar_ops,
                     uint64_t n_cigar, spb_nucl_aln **out) {
    SPB_CHECK(c && p && out && (rows || n_rows == 0) && (cigar_ops || n_cigar == 0), "spb_nucl_rescore: NULL argument");
    SPB_CUDA(cudaSetDevice(c->device));
    if (c->s_a.ensure(std::max<uint64_t>(n_rows, 1) * sizeof(spb_aln)) || c->s_b.ensure(n_cigar + 16)) return 1;
    if (n_rows) SPB_CUDA(cudaMemcpyAsync(c->s_a.p, rows, n_rows * sizeof(spb_aln), cudaMemcpyHostToDevice, c->stream));
    if (n_cigar) SPB_CUDA(cudaMemcpyAsync(c->s_b.p, cigar_ops, n_cigar, cudaMemcpyHostToDevice, c->stream));
    return spb::rescore(c, p, c->s_a.as<spb_aln>(), n_rows, c->s_b.as<char>(), n_cigar, out, nullptr, "spb_nucl_rescore");
}

int spb_nucl_rescore_resident(spb_ctx *c, const spb_nucl_params *p, spb_nucl_aln **out, uint64_t *n_rows, double *kernel_ms) {
    SPB_CHECK(c && p && out && n_rows, "spb_nucl_rescore_resident: NULL argument");
    SPB_CHECK(c->have_aln, "spb_nucl_rescore_resident: no resident alignment result");
    SPB_CHECK(c->ntQ == c->Q, "spb_nucl_rescore_resident: %u nucleotide query entries for a batch of %u queries", c->ntQ, c->Q);
    SPB_CUDA(cudaSetDevice(c->device));
    *n_rows = c->n_acc;
    return spb::rescore(c, p, c->d_rows.as<spb_aln>(), c->n_acc, c->d_cig.as<char>(), c->n_cigar, out, kernel_ms,
                        "spb_nucl_rescore_resident");
}

void spb_free_nucl_result(spb_ctx *c, spb_nucl_aln *r) {
    if (c && r) c->pin_release(r);
}

}  // extern "C"
