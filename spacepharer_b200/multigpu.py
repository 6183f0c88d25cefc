"""One-process-per-GPU driver of the search step (target split like MMseqs2 --split, or query split).

Mirrors how the reference distributes `prefilter`/`align` (M/prefiltering/Prefiltering.cpp:258-362,559-643 target /
query split, M/alignment/Alignment.cpp:226-249): every rank owns one `Searcher` (one CUDA context object) and

  split_mode "target": rank r indexes targets [from_r, to_r) chosen like DBReader::decomposeDomainByAminoAcid
      (M/commons/DBReader.cpp:1171-1213); queries are replicated; every rank prefilters all queries against its shard
      (tie-break bins use global ids: id_base = from_r); the per-query hit lists of query range r are sent to rank r
      device to device (NCCL all_to_all_single over NVLink), merged there (spb_prefilter_merge_shards) and aligned
      there against the replicated unmasked target residues with the whole-DB residue count in the E-value.  The
      per-target rules of the prefilter are shard-local, so the merged list equals the UNSPLIT list (unlike the
      reference's lossy --split heuristic, Prefiltering.cpp:343-346).  Lists reaching max-seqs are cut in the
      reference's order as well: every hit travels with its truncation key (capped score, id & (BINSIZE-1), stream
      order of the winning event -- all shard independent), and the first max-seqs of the union in (key, global id)
      order are the reference's selection (QueryMatcher.cpp:131-164).
  split_mode "query": every rank holds the full index and processes its own slice of the query batch; no collective
      on the data path.
  split_mode "index": the k-mer table (direct table + posting lists) is cut into `world` contiguous k-mer ranges, rank r
      keeps range r in its HBM, and the lookup kernel of every rank gathers the entries of the other ranges with plain
      loads on peer-mapped pointers over NVLink (CUDA IPC handles exchanged once at load time).  Queries are partitioned
      like in "query": no hit list is exchanged or merged, the result is the unsplit one by construction, and both the
      table capacity and the throughput scale with the number of GPUs.

torch / torch.distributed are plumbing only (process group, NCCL, device buffers of the exchange); all arithmetic is in
libspb200's kernels.
"""
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


def exchange_hit_lists(hits, offsets, ranges, rank, group=None, keys=None):
    """All-to-all of per-query hit lists by query range (backend agnostic: NCCL on device tensors, gloo on CPU tensors).

    hits: int32 tensor [n, 3] (global id, score, diag), all queries of the batch in query order;
    offsets: int64 tensor [Q+1]; keys: optional int64 tensor [n] of truncation keys travelling with the hits.
    Returns (recv int32 [m, 3] shard-major, counts int32 [world * nq] shard-major[, recv_keys int64 [m]])
    for this rank's query range."""
    import torch
    import torch.distributed as dist
    G = len(ranges)
    bounds = torch.tensor([lo for lo, _ in ranges] + [ranges[-1][1]], dtype=torch.int64, device=offsets.device)
    cut = offsets[bounds].cpu().tolist()                      # hit offsets at the range boundaries (one small D2H)
    send_rows = [cut[r + 1] - cut[r] for r in range(G)]
    lo, hi = ranges[rank]
    nq = hi - lo
    cnt_all = (offsets[1:] - offsets[:-1]).to(torch.int32).contiguous()
    cnt_recv = torch.empty(G * nq, dtype=torch.int32, device=hits.device)
    dist.all_to_all_single(cnt_recv, cnt_all, output_split_sizes=[nq] * G, input_split_sizes=[h - l for l, h in ranges],
                           group=group)
    # rows to expect from every shard = the sum of its per-query counts (no separate size exchange)
    recv_rows = cnt_recv.view(G, nq).sum(dim=1, dtype=torch.int64).cpu().tolist() if nq else [0] * G
    recv = torch.empty((sum(recv_rows), 3), dtype=torch.int32, device=hits.device)
    dist.all_to_all_single(recv, hits.contiguous(), output_split_sizes=recv_rows, input_split_sizes=send_rows, group=group)
    if keys is None:
        return recv, cnt_recv
    recv_keys = torch.empty(sum(recv_rows), dtype=torch.int64, device=hits.device)
    dist.all_to_all_single(recv_keys, keys.contiguous(), output_split_sizes=recv_rows, input_split_sizes=send_rows, group=group)
    return recv, cnt_recv, recv_keys


def merge_hit_lists_reference(recv, cnt_recv, world, nq, keys=None, max_seqs=None):
    """numpy restatement of spb_prefilter_merge_shards (tests): per query concatenate the shard segments and sort by
    (score desc, id asc); with truncation keys, lists longer than max_seqs keep the first max_seqs in (key, id) order.
    Returns (offsets int64[nq+1], hits int32[m, 3])."""
    recv = np.asarray(recv).reshape(-1, 3)
    if keys is not None:
        keys = np.asarray(keys).astype(np.uint64)
    cnt = np.asarray(cnt_recv).reshape(world, nq).astype(np.int64)
    seg = np.zeros(world * nq + 1, dtype=np.int64)
    seg[1:] = np.cumsum(cnt.reshape(-1))
    out, off = [], [0]
    for q in range(nq):
        rows = np.concatenate([recv[seg[s * nq + q]:seg[s * nq + q + 1]] for s in range(world)]) if world else recv[:0]
        if keys is not None and max_seqs is not None and len(rows) > max_seqs:
            kk = np.concatenate([keys[seg[s * nq + q]:seg[s * nq + q + 1]] for s in range(world)])
            rows = rows[np.lexsort((rows[:, 0], kk))[:max_seqs]]
        order = np.lexsort((rows[:, 0], -rows[:, 1]))
        out.append(rows[order])
        off.append(off[-1] + len(rows))
    return np.array(off, dtype=np.int64), (np.concatenate(out) if out else recv[:0])


class _DevView:
    """zero-copy torch view of library-owned device memory (__cuda_array_interface__)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedSearcher:
    def __init__(self, device=0, rank=0, world=1, split_mode="query", lanes=1, shard_group=0, **kw):
        """lanes > 1: that many contexts per GPU sharing one index (api.Searcher.lane), used by run_stream for a stream of
        batches -- the kernels of one lane fill the tails and host-sync bubbles of the others."""
        self.rank, self.world, self.split_mode = rank, world, split_mode
        # index split: GPUs that share one copy of the k-mer table; world must be a multiple of it.  0 = pairs (all GPUs when the
        # world is odd): measured on 8 x B200, groups of 2 / 4 / 8 reach 0.78 / 0.48 / 0.18 of the replica throughput -- remote
        # 16-byte reads through NVSwitch do not scale with the number of peers, so larger groups are for capacity only
        self.shard_group = (2 if world % 2 == 0 else world) if not shard_group else int(shard_group)
        assert self.shard_group >= 1 and world % self.shard_group == 0, "shard_group must divide the number of GPUs"
        self.n_lanes = max(1, int(lanes))
        self._lanes = None
        self.searcher = api.Searcher(device=device, **kw)
        self.device = device
        self.shard = None
        self.Q = 0
        self.q_lo = 0

    def close(self):
        for ln in (self._lanes or [])[1:]:
            ln.close()
        self._lanes = None
        self.searcher.close()

    def launches(self):
        return sum(ln.launches() for ln in (self._lanes or [self.searcher]))

    def lanes(self):
        """the lane contexts (lane 0 is the searcher itself); created on first use, after the targets are loaded"""
        if self._lanes is None:
            self._lanes = [self.searcher] + [self.searcher.lane() for _ in range(self.n_lanes - 1)]
        return self._lanes

    def run_stream(self, batches, db_res_count, steps=None, resident=False, consume=None, before_step=None, **kw):
        """A stream of query batches over the lanes, one host thread per lane (ctypes releases the GIL inside the library).
        batches: list of (qres, qoff) host buffers; step i searches batches[i % len(batches)] (steps defaults to one pass).
        resident=False: every step uploads its batch, searches it and reads rows + CIGARs back into pinned host buffers;
        consume(step, q_lo, view), if given, sees the zero-copy result view before it is handed back.
        resident=True: the batches are uploaded once per lane (lane l owns batches[l]), the steps only run the kernels.
        before_step(lane, searcher), if given, runs in the lane's thread ahead of every step (bench.py: L2 flush).
        Returns a list with one stats dict per step (step order).  Query split / single GPU only."""
        import threading
        assert not (self.world > 1 and self.split_mode == "target"), "run_stream: lanes are for query split / one GPU"
        lanes = self.lanes()
        steps = len(batches) if steps is None else steps
        pk = {k: kw[k] for k in ("kmer_thr", "max_seqs", "min_ungapped_score", "bin_size", "comp_bias") if k in kw}
        ak = {k: kw[k] for k in ("evalue", "gap_open", "gap_extend", "comp_bias", "gumbel", "byte_lanes", "word_lanes") if k in kw}
        mine = [self._my_queries(q, o) for q, o in batches]
        if resident:
            for l, ln in enumerate(lanes):
                q, o, _ = mine[l % len(mine)]
                ln.set_queries(q, o)
        out, errs = [None] * steps, []
        nxt, lock = [0], threading.Lock()

        def work(l):
            ln = lanes[l]
            try:
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= steps:
                        return
                    q, o, q_lo = mine[(l if resident else i) % len(mine)]
                    if before_step is not None:
                        before_step(l, ln)
                    if not resident:
                        ln.set_queries(q, o)
                    pst = ln.prefilter_resident(**pk)
                    ln.pairs_from_prefilter()
                    ast = ln.align_resident(db_res_count, **ak)
                    st = dict(prefilter=pst, align=ast, lane=l)
                    if not resident:
                        v = ln.align_fetch_view()
                        st["d2h_bytes"], st["rows"] = v.nbytes, len(v.rows)
                        if consume is not None:
                            consume(i, q_lo, v)
                        v.close()
                    out[i] = st
            except Exception as e:  # noqa: BLE001 -- re-raised in the caller's thread
                errs.append(e)

        th = [threading.Thread(target=work, args=(l,)) for l in range(len(lanes))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return out

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
        elif self.world > 1 and self.split_mode == "index":
            import torch.distributed as dist
            self.shard = (0, T)
            st = s.build_index(masked, toff)
            G = self.shard_group
            g0 = (self.rank // G) * G                   # first rank of my shard group
            if G > 1:
                s.shard_index(self.rank - g0, G)        # keep k-mer range (rank within the group), free the rest of the table
            handles = [None] * self.world
            dist.all_gather_object(handles, s.shard_export() if G > 1 else b"")
            for r in range(g0, g0 + G):
                if r != self.rank:
                    s.shard_attach_ipc(r - g0, handles[r])
            dist.barrier()
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
        if self.world > 1 and self.split_mode in ("query", "index"):
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

    def search_host(self, qres, qoff, db_res_count, copy=True, **kw):
        """Public end-to-end call: host query buffers in, sorted alignment rows + CIGARs out (host).
        Every rank returns the rows of the queries it owns; row["query"] + q_lo is the batch-global query index.
        copy=False returns views of the library's pinned result buffers plus `view` (close() it when done)."""
        self.set_queries(qres, qoff)
        st = self.step_resident(db_res_count, **kw)
        if not copy:
            v = self.searcher.align_fetch_view()
            return dict(offsets=v.offsets, rows=v.rows, cigar=v.cigar, stats=st, q_lo=self.q_lo, d2h_bytes=v.nbytes, view=v)
        off, rows, cig, _ = self.searcher.align_fetch()
        if self.q_lo:
            rows = rows.copy()
            rows["query"] += self.q_lo
            self_q_lo = 0
        else:
            self_q_lo = self.q_lo
        return dict(offsets=off, rows=rows, cigar=cig, stats=st, q_lo=self_q_lo,
                    d2h_bytes=off.nbytes + rows.nbytes + len(cig))

    def search_host_begin(self, qres, qoff, db_res_count, **kw):
        """Pipelined form of search_host for a stream of batches: uploads and searches the batch, starts the copy of its
        rows + CIGARs to pinned host memory on a second stream and returns; result["view"].wait() gives the views (offsets,
        rows, cigar) once they have landed -- typically after the NEXT batch has been submitted."""
        self.set_queries(qres, qoff)
        st = self.step_resident(db_res_count, **kw)
        return dict(view=self.searcher.align_fetch_begin(), stats=st, q_lo=self.q_lo)

    # ---- target split: hit-list exchange + merge -----------------------------------------------
    def _prefilter_target_split(self, pk):
        import torch
        s = self.searcher
        frm, _ = self.shard
        Q, G = self.Q, self.world
        pst = s.prefilter_resident(id_base=frm, emit_trunc_keys=1, **pk)
        hp, op, n = s.prefilter_device_result()
        kp = s.prefilter_device_trunc_keys()
        dev = torch.device("cuda", self.device)
        stream = torch.cuda.ExternalStream(s.stream_ptr(), device=dev)
        with torch.cuda.stream(stream):
            offsets = torch.as_tensor(_DevView(op, (Q + 1,), "<i8"), device=dev)
            hits = (torch.as_tensor(_DevView(hp, (n, 3), "<i4"), device=dev) if n
                    else torch.empty((0, 3), dtype=torch.int32, device=dev)).clone()
            hits[:, 0] += frm                         # shard-local -> global target ids
            keys = (torch.as_tensor(_DevView(kp, (n,), "<i8"), device=dev) if n
                    else torch.empty((0,), dtype=torch.int64, device=dev))
            ranges = query_ranges(Q, G)
            recv, cnt_recv, recv_keys = exchange_hit_lists(hits, offsets, ranges, self.rank, keys=keys)
            stream.synchronize()
        lo, hi = ranges[self.rank]
        s.merge_shards(lo, hi, recv.data_ptr(), recv.shape[0], cnt_recv.data_ptr(), G, min(pk.get("max_seqs", 300), self.T),
                       recv_keys_ptr=recv_keys.data_ptr())
        pst["n_hits_shard"] = pst["n_hits"]
        pst["n_hits"] = int(recv.shape[0])
        return pst
