// frames.cu -- query preparation on the device: nucleotide spacers -> six-frame peptide fragments.
//
// Replaces, for the spacer query set of createsetdb --extractorf-spacer 1 (data/createsetdb.sh:88-113),
//     extractorfs  --min-length 9 --orf-start-mode 1   (M/util/extractorfs.cpp:19-160, Orf::findForward  M/commons/Orf.cpp:230-352)
//     translatenucs                                     (M/util/translatenucs.cpp:13-115, TranslateNucl::translate, canonical table)
// so that nucleotide spacers go straight to the resident query batch of the search without a DB round trip.
//
// Semantics reproduced (ANY_TO_STOP): on each strand (forward, then reverse complement) the three frames are walked
// position by position; a fragment runs from the start of the frame / the codon after a stop to the codon before the
// next stop or the last complete codon; it is kept when it has min..max non-stop codons.  Fragments are emitted in the
// order in which they END (position of the terminating codon, frames interleaved) -- the order of the reference's
// per-contig output, which createRenumberedDB (SORT_BY_ID_OFFSET) turns into the DB keys of the frames.
// Stop codons are the exact strings TAA / TAG / TGA; a codon with an ambiguous base is never a stop and translates
// to the residue all its expansions share, else X.
#include <cub/cub.cuh>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace spb {

#define NT_INC 5u   // beyond the end of the sequence (the reference pads with CHAR_MAX)

// numeric residue (ACDEFGHIKLMNPQRSTVWYX order, 255 = stop) of codon 16 a + 4 b + c with A C G T = 0 1 2 3
__constant__ uint8_t c_codon_aa[64];

struct FrameArgs {
    const uint8_t *nt;
    const uint64_t *ntoff;
    uint32_t N;
    int minCodons, maxCodons;
    // pass 1 out
    uint32_t *cntFrames;       // [2N + 1]
    uint64_t *cntRes;          // [2N + 1]
    // pass 2 in / out
    const uint32_t *frameBase; // exclusive scans of the above
    const uint64_t *resBase;
    spb_orf *orfs;
    uint64_t *qoff;            // [F + 1]
    uint8_t *qres;
};

__device__ __forceinline__ uint32_t nt_at(const uint8_t *__restrict__ s, int L, int strand, int p) {
    if (p >= L) return NT_INC;
    if (strand == 0) return min((uint32_t)__ldg(s + p), 4u);
    const uint32_t c = min((uint32_t)__ldg(s + (L - 1 - p)), 4u);
    return c < 4u ? 3u - c : 4u;
}

__device__ __forceinline__ uint8_t translate_codon(uint32_t a, uint32_t b, uint32_t c) {
    if (a < 4u && b < 4u && c < 4u) return c_codon_aa[16 * a + 4 * b + c];
    // ambiguous base: the residue shared by every expansion, else X (NCBI translation automaton)
    int res = -1;
    for (uint32_t x = (a < 4u ? a : 0u); x <= (a < 4u ? a : 3u); x++)
        for (uint32_t y = (b < 4u ? b : 0u); y <= (b < 4u ? b : 3u); y++)
            for (uint32_t z = (c < 4u ? c : 0u); z <= (c < 4u ? c : 3u); z++) {
                const int r = c_codon_aa[16 * x + 4 * y + z];
                if (res == -1) res = r;
                else if (res != r) return (uint8_t)SPB_X;
            }
    return res == 255 ? (uint8_t)SPB_X : (uint8_t)res;
}

// one thread per (sequence, strand); EMIT = false counts, EMIT = true writes records and residues
template <bool EMIT>
__global__ void __launch_bounds__(128) frames_kernel(FrameArgs A) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2u * A.N) return;
    const uint32_t s = idx >> 1;
    const int strand = (int)(idx & 1u);
    const uint64_t o = A.ntoff[s];
    const int L = (int)(A.ntoff[s + 1] - o);
    const uint8_t *seq = A.nt + o;
    uint32_t nFrames = 0;
    uint64_t nRes = 0;
    uint32_t fbase = 0;
    uint64_t rbase = 0;
    if (EMIT) { fbase = A.frameBase[idx]; rbase = A.resBase[idx]; }
    if (L >= 3) {
        bool inside[3] = {true, true, true}, hasStart[3] = {false, false, false};
        int cnt[3] = {0, 0, 0}, from[3] = {0, 1, 2};
        // sliding window of six bases: the codon at `position` and the one after it
        uint32_t w0 = nt_at(seq, L, strand, 0), w1 = nt_at(seq, L, strand, 1), w2 = nt_at(seq, L, strand, 2);
        uint32_t w3 = nt_at(seq, L, strand, 3), w4 = nt_at(seq, L, strand, 4);
        const int nPos = ((L - 2 + 2) / 3) * 3;     // positions visited: i < L - 2 in steps of 3, three positions each
        for (int position = 0; position < nPos; position++) {
            const uint32_t w5 = nt_at(seq, L, strand, position + 5);
            const int f = position % 3;
            const bool incomplete = (w0 == NT_INC) | (w1 == NT_INC) | (w2 == NT_INC);
            const bool isLast = !incomplete && ((w3 == NT_INC) | (w4 == NT_INC) | (w5 == NT_INC));
            if (!inside[f]) {
                inside[f] = true;
                hasStart[f] = true;
                from[f] = position;
                cnt[f] = 0;
            }
            // TAA TAG TGA
            const bool stop = (w0 == 3u) && ((w1 == 0u && (w2 == 0u || w2 == 2u)) || (w1 == 2u && w2 == 0u));
            if (!stop) cnt[f]++;
            if (stop || isLast) {
                inside[f] = false;
                const int n = cnt[f];
                if (!(n == 0 && stop) && n >= A.minCodons && n <= A.maxCodons) {
                    const int to = position + ((isLast && !stop) ? 2 : -1);
                    if (EMIT) {
                        const uint32_t k = fbase + nFrames;
                        spb_orf r;
                        r.contig = s;
                        r.from = (uint32_t)from[f];
                        r.to = (uint32_t)to;
                        r.strand = (uint8_t)strand;
                        r.incomplete_start = (uint8_t)!hasStart[f];
                        r.incomplete_end = (uint8_t)!stop;
                        r.pad = 0;
                        A.orfs[k] = r;
                        A.qoff[k] = rbase + nRes;
                        uint8_t *dst = A.qres + rbase + nRes;
                        for (int c = 0; c < n; c++) {
                            const int p = from[f] + 3 * c;
                            dst[c] = translate_codon(nt_at(seq, L, strand, p), nt_at(seq, L, strand, p + 1), nt_at(seq, L, strand, p + 2));
                        }
                    }
                    nFrames++;
                    nRes += (uint64_t)n;
                }
            }
            w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5;
        }
    }
    if (!EMIT) {
        A.cntFrames[idx] = nFrames;
        A.cntRes[idx] = nRes;
    }
}

}  // namespace spb

using namespace spb;

struct FramesResultBox {
    spb_frames_result pub;
    spb_ctx *owner;
};

extern "C" void spb_free_frames_result(spb_frames_result *r) {
    if (!r) return;
    FramesResultBox *box = reinterpret_cast<FramesResultBox *>(r);
    if (box->owner) {
        box->owner->pin_release(r->offsets);
        box->owner->pin_release(r->residues);
        box->owner->pin_release(r->orfs);
    }
    free(box);
}

static int upload_codon_table() {
    // TranslateNucl CANONICAL (M/commons/TranslateNucl.h:114-118): ncbieaa over T C A G
    static const char *ncbieaa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const char *aaOrder = "ACDEFGHIKLMNPQRSTVWYX";
    const int tcag[4] = {2, 1, 3, 0};   // index of A C G T in "TCAG"
    uint8_t tab[64];
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++)
            for (int c = 0; c < 4; c++) {
                const char r = ncbieaa[16 * tcag[a] + 4 * tcag[b] + tcag[c]];
                const char *p = strchr(aaOrder, r);
                tab[16 * a + 4 * b + c] = (r == '*' || !p) ? 255 : (uint8_t)(p - aaOrder);
            }
    SPB_CUDA(cudaMemcpyToSymbol(c_codon_aa, tab, sizeof(tab)));
    return 0;
}

extern "C" int spb_spacer_frames(spb_ctx *c, const uint8_t *nt, const uint64_t *nt_offsets, uint32_t N, int min_codons,
                                 int max_codons, int make_resident, spb_frames_result **out) {
    SPB_CHECK(c && nt_offsets && out, "spb_spacer_frames: NULL argument");
    SPB_CHECK(N == 0 || nt, "spb_spacer_frames: NULL nucleotides");
    SPB_CHECK(min_codons >= 1 && max_codons >= min_codons, "spb_spacer_frames: need 1 <= min_codons <= max_codons");
    SPB_CHECK(N < 0x7FFFFFF0u, "spb_spacer_frames: too many sequences in one call; split the batch");
    *out = nullptr;
    SPB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (upload_codon_table()) return 1;
    const uint64_t NT = nt_offsets[N];
    for (uint32_t i = 0; i < N; i++)
        SPB_CHECK(nt_offsets[i + 1] >= nt_offsets[i] && nt_offsets[i + 1] - nt_offsets[i] < 0x7FFFFFF0ull,
                  "spb_spacer_frames: offsets not monotone at %u", i);
    const size_t n2 = (size_t)2 * N + 1;
    // scratch: s_a nucleotides, s_b offsets, s_c frame counts + bases (u32), s_d residue counts + bases (u64)
    if (c->s_a.ensure(NT + 64) || c->s_b.ensure(((size_t)N + 1) * 8) || c->s_c.ensure(n2 * 8 + 64) || c->s_d.ensure(n2 * 16 + 64)) return 1;
    if (NT) SPB_CUDA(cudaMemcpyAsync(c->s_a.p, nt, NT, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaMemcpyAsync(c->s_b.p, nt_offsets, ((size_t)N + 1) * 8, cudaMemcpyHostToDevice, st));
    uint32_t *cntF = c->s_c.as<uint32_t>(), *baseF = cntF + n2;
    uint64_t *cntR = c->s_d.as<uint64_t>(), *baseR = cntR + n2;
    SPB_CUDA(cudaMemsetAsync(cntF + 2 * (size_t)N, 0, 4, st));
    SPB_CUDA(cudaMemsetAsync(cntR + 2 * (size_t)N, 0, 8, st));
    FrameArgs A;
    memset(&A, 0, sizeof(A));
    A.nt = c->s_a.as<uint8_t>();
    A.ntoff = c->s_b.as<uint64_t>();
    A.N = N;
    A.minCodons = min_codons;
    A.maxCodons = max_codons;
    A.cntFrames = cntF;
    A.cntRes = cntR;
    SPB_CUDA(cudaEventRecord(c->ev0, st));
    const unsigned blocks = (unsigned)((2 * (size_t)N + 127) / 128);
    if (N) {
        frames_kernel<false><<<blocks, 128, 0, st>>>(A);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    size_t tmpBytes = 0, tmp2 = 0;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cntF, baseF, (int)n2, st));
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, cntR, baseR, (int)n2, st));
    if (c->s_cub.ensure(std::max(tmpBytes, tmp2))) return 1;
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmpBytes, cntF, baseF, (int)n2, st));
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(c->s_cub.p, tmp2, cntR, baseR, (int)n2, st));
    uint32_t F = 0;
    uint64_t R = 0;
    SPB_CUDA(cudaMemcpyAsync(&F, baseF + 2 * (size_t)N, 4, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaMemcpyAsync(&R, baseR + 2 * (size_t)N, 8, cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    // s_e frame records, s_g offsets, s_h residues
    if (c->s_e.ensure(std::max<size_t>(F, 1) * sizeof(spb_orf)) || c->s_g.ensure(((size_t)F + 1) * 8) || c->s_h.ensure(R + 64)) return 1;
    A.frameBase = baseF;
    A.resBase = baseR;
    A.orfs = c->s_e.as<spb_orf>();
    A.qoff = c->s_g.as<uint64_t>();
    A.qres = c->s_h.as<uint8_t>();
    if (N) {
        frames_kernel<true><<<blocks, 128, 0, st>>>(A);
        SPB_CUDA(cudaGetLastError()); c->n_launches++;
    }
    SPB_CUDA(cudaMemcpyAsync(A.qoff + F, &R, 8, cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaEventRecord(c->ev1, st));

    FramesResultBox *box = (FramesResultBox *)calloc(1, sizeof(FramesResultBox));
    SPB_CHECK(box != nullptr, "spb_spacer_frames: out of memory");
    box->owner = c;
    spb_frames_result *res = &box->pub;
    res->n_sequences = N;
    res->n_frames = F;
    res->n_residues = R;
    res->offsets = (uint64_t *)c->pin_acquire(((size_t)F + 1) * 8);
    res->residues = (uint8_t *)c->pin_acquire(R + 1);
    res->orfs = (spb_orf *)c->pin_acquire(std::max<size_t>(F, 1) * sizeof(spb_orf));
    if (!res->offsets || !res->residues || !res->orfs) { spb_free_frames_result(res); return 1; }
    cudaError_t e = cudaMemcpyAsync(res->offsets, A.qoff, ((size_t)F + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && R) e = cudaMemcpyAsync(res->residues, A.qres, R, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && F) e = cudaMemcpyAsync(res->orfs, A.orfs, (size_t)F * sizeof(spb_orf), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        spb_free_frames_result(res);
        spb::set_error("spb_spacer_frames: %s", cudaGetErrorString(e));
        return 1;
    }
    float ms = 0;
    SPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    res->kernel_ms = ms;
    if (make_resident) {
        // the frames become the resident query batch: residues device to device, offsets from the copy just fetched
        if (spb::queries_install(c, A.qres, true, res->offsets, F)) { spb_free_frames_result(res); return 1; }
    }
    *out = res;
    return 0;
}
