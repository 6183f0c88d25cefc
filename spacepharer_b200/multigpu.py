"""One-process-per-GPU driver of the search step (target split like MMseqs2 --split, or query replicas).

Mirrors how the reference distributes `prefilter`/`align` (M/prefiltering/Prefiltering.cpp:258-362,559-643 target /
query split, M/alignment/Alignment.cpp:226-249): every rank owns one `Searcher` (one CUDA context object) and

  split_mode "target": rank r indexes targets [from_r, to_r) chosen like DBReader::decomposeDomainByAminoAcid
      (M/commons/DBReader.cpp:1171-1213); queries are replicated; every rank prefilters all queries against its shard
      and emits its candidates with their truncation keys; candidates of query range r are sent to rank r over NCCL
      (all_to_all_single, device buffers), merged there to exactly the unsplit top-max-seqs list, and aligned there
      against the replicated (unmasked) target residues.  Unlike the reference's lossy --split heuristic
      (Prefiltering.cpp:343-346) the merged result equals the unsplit run.
  split_mode "query": every rank holds the full index and processes its own slice of the query batch; no collective
      on the data path.

torch / torch.distributed are plumbing only (process group, NCCL); all arithmetic is in libspb200's kernels.
"""
import ctypes as C

import numpy as np

from . import api, capi


def decompose_by_amino_acid(toff, world):
    """DBReader::decomposeDomainByAminoAcid (M/commons/DBReader.cpp:1171-1213): greedy cuts at
    ceil(total / world) residues (+2 per entry like the on-disk entry length); returns [(from, size)] * world."""
    T = len(toff) - 1
    lens = (np.diff(np.asarray(toff, dtype=np.uint64)).astype(np.int64) + 2)
    total = int(lens.sum())
    per = (total + world - 1) // world if world > 0 else total
    out = []
    start = 0
    acc = 0
    for i in range(T):
        acc += int(lens[i])
        if acc >= per and len(out) < world - 1:
            out.append((start, i + 1 - start))
            start = i + 1
            acc = 0
    out.append((start, T - start))
    while len(out) < world:
        out.append((T, 0))
    return out


def query_ranges(Q, world):
    """contiguous query ranges owned by each rank: [(lo, hi)]"""
    return [((Q * r) // world, (Q * (r + 1)) // world) for r in range(world)]


class ShardedSearcher:
    def __init__(self, device=0, rank=0, world=1, split_mode="target", **kw):
        self.rank, self.world, self.split_mode = rank, world, split_mode
        self.searcher = api.Searcher(device=device, **kw)
        self.device = device
        self.shard = None
        self.Q = 0

    def close(self):
        self.searcher.close()

    def launches(self):
        return self.searcher.launches()

    # ---- targets -----------------------------------------------------------------------------
    def load_targets(self, masked, tres, toff, keys=None):
        """masked: residues the index sees (tantan-masked); tres: raw residues (alignment); toff: uint64[T+1]."""
        toff = np.ascontiguousarray(toff, dtype=np.uint64)
        T = len(toff) - 1
        s = self.searcher
        if self.world > 1 and self.split_mode == "target":
            frm, size = decompose_by_amino_acid(toff, self.world)[self.rank]
            lo, hi = int(toff[frm]), int(toff[frm + size])
            self.shard = (frm, size)
            st = s.build_index(masked[lo:hi], toff[frm:frm + size + 1] - toff[frm])
        else:
            self.shard = (0, T)
            st = s.build_index(masked, toff)
        s.set_align_targets(tres, toff, keys)
        self.T = T
        return st

    # ---- queries -----------------------------------------------------------------------------
    def queries_per_step(self, Q):
        return Q

    def _my_queries(self, qres, qoff):
        if self.world > 1 and self.split_mode == "query":
            lo, hi = query_ranges(len(qoff) - 1, self.world)[self.rank]
            qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
            sub = qoff[lo:hi + 1] - qoff[lo]
            return qres[int(qoff[lo]):int(qoff[hi])], sub, lo
        return qres, qoff, 0

    def set_queries(self, qres, qoff):
        q, o, self.q_lo = self._my_queries(qres, qoff)
        self.Q = len(qoff) - 1
        self.searcher.set_queries(q, o)

    def step_resident(self, db_res_count, **kw):
        """prefilter -> (merge) -> align on the resident query batch; nothing leaves HBM."""
        s = self.searcher
        pk = {k: kw[k] for k in ("kmer_thr", "max_seqs", "min_ungapped_score", "bin_size", "comp_bias") if k in kw}
        ak = {k: kw[k] for k in ("evalue", "gap_open", "gap_extend", "comp_bias", "gumbel", "byte_lanes", "word_lanes") if k in kw}
        if self.world > 1 and self.split_mode == "target":
            pst = self._prefilter_target_split(pk)
        else:
            pst = s.prefilter_resident(**pk)
        s.pairs_from_prefilter()
        ast = s.align_resident(db_res_count, **ak)
        return dict(prefilter=pst, align=ast)

    def search_host(self, qres, qoff, db_res_count, **kw):
        """Public end-to-end call: host query buffers in, sorted alignment rows + CIGARs out (host)."""
        self.set_queries(qres, qoff)
        st = self.step_resident(db_res_count, **kw)
        off, rows, cig, _ = self.searcher.align_fetch()
        if self.q_lo:
            rows = rows.copy()
            rows["query"] += self.q_lo
        return dict(offsets=off, rows=rows, cigar=cig, stats=st, q_lo=self.q_lo,
                    d2h_bytes=off.nbytes + rows.nbytes + len(cig))

    # ---- target split: candidate exchange + merge ---------------------------------------------
    def _prefilter_target_split(self, pk):
        import torch
        import torch.distributed as dist
        s = self.searcher
        frm, _ = self.shard
        Q, G = self.Q, self.world
        pst = s.prefilter_candidates_resident(id_base=frm, global_targets=self.T, **pk)
        cands, coff = s.candidates_view()            # torch views of library-owned device memory
        ranges = query_ranges(Q, G)
        bounds = torch.tensor([lo for lo, _ in ranges] + [Q], device=cands.device, dtype=torch.int64)
        cut = coff[bounds].cpu().tolist()            # candidate offsets at the range boundaries (one small D2H)
        send_counts = [cut[r + 1] - cut[r] for r in range(G)]
        sc = torch.tensor(send_counts, dtype=torch.int64, device=cands.device)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv_counts = rc.cpu().tolist()
        lo, hi = ranges[self.rank]
        nq = hi - lo
        # per-query counts of my range from every rank
        cnt_all = (coff[1:] - coff[:-1]).to(torch.int32)
        cnt_recv = torch.empty(G * nq, dtype=torch.int32, device=cands.device)
        dist.all_to_all_single(cnt_recv, cnt_all, output_split_sizes=[nq] * G,
                               input_split_sizes=[h - l for l, h in ranges])
        recv = torch.empty((sum(recv_counts), cands.shape[1]), dtype=cands.dtype, device=cands.device)
        dist.all_to_all_single(recv, cands, output_split_sizes=recv_counts, input_split_sizes=send_counts)
        mst = s.merge_candidates(lo, hi, recv, cnt_recv, G, **pk)
        pst.update(merge_ms=mst["merge_ms"], n_hits=mst["n_hits"])
        return pst
