"""Multi-GPU target split: the merged per-query lists equal the unsplit prefilter result, and the aligned rows equal the
single-GPU rows.  The merge kernel is exercised on one GPU (shards run one after the other); the NCCL path needs two."""
import os
import socket

import numpy as np
import pytest

from oracle import pyoracle as O  # noqa: F401  (keeps the import order of the other GPU tests)

pytestmark = pytest.mark.gpu


def _workload():
    from spacepharer_b200 import synth
    tres, toff, _ = synth.make_targets(n_genomes=80, seed=11)
    qres, qoff, _ = synth.make_query_frames(2500, tres, toff, seed=12)
    return tres, toff, qres, qoff


def _family_workload():
    """families of near-identical targets: most queries see far more candidates than max-seqs, with many score ties"""
    rng = np.random.default_rng(99)
    T, Q = [], []
    for f in range(30):
        base = rng.integers(0, 20, 90).astype(np.uint8)
        for _ in range(50):
            t = base.copy()
            for _ in range(int(rng.integers(0, 4))):
                t[int(rng.integers(0, 90))] = rng.integers(0, 20)
            T.append(np.concatenate([rng.integers(0, 20, int(rng.integers(0, 40))).astype(np.uint8), t]))
        for _ in range(12):
            L = int(rng.integers(10, 22))
            p = int(rng.integers(0, 90 - L))
            q = base[p:p + L].copy()
            if rng.random() < 0.5:
                q[int(rng.integers(0, L))] = rng.integers(0, 20)
            Q.append(q)
    order = rng.permutation(len(T))
    tres, toff = O.concat([T[i] for i in order])
    qres, qoff = O.concat(Q)
    return tres, toff, qres, qoff


def _shard_lists_on_one_gpu(s, tres, toff, qres, qoff, G, max_seqs, bin_size=2):
    """every target shard prefiltered in turn on this GPU (as its own rank would), lists and truncation keys kept on the
    device -> (recv int32[m,3] shard-major with global ids, keys int64[m], counts int32[G*Q])"""
    import torch
    from spacepharer_b200 import multigpu
    Q = len(qoff) - 1
    parts = multigpu.decompose_by_amino_acid(toff, G)
    hs, ks, cs = [], [], []
    dev = torch.device("cuda", 0)
    for frm, size in parts:
        lo, hi = int(toff[frm]), int(toff[frm + size])
        s.build_index(tres[lo:hi], toff[frm:frm + size + 1] - toff[frm])
        s.set_queries(qres, qoff)
        s.prefilter_resident(id_base=frm, emit_trunc_keys=1, max_seqs=max_seqs, bin_size=bin_size)
        hp, op, n = s.prefilter_device_result()
        kp = s.prefilter_device_trunc_keys()
        torch.cuda.synchronize()
        off = torch.as_tensor(multigpu._DevView(op, (Q + 1,), "<i8"), device=dev).clone()
        h = (torch.as_tensor(multigpu._DevView(hp, (n, 3), "<i4"), device=dev).clone() if n
             else torch.empty((0, 3), dtype=torch.int32, device=dev))
        k = (torch.as_tensor(multigpu._DevView(kp, (n,), "<i8"), device=dev).clone() if n
             else torch.empty((0,), dtype=torch.int64, device=dev))
        h[:, 0] += frm
        hs.append(h)
        ks.append(k)
        cs.append((off[1:] - off[:-1]).to(torch.int32))
    return torch.cat(hs).contiguous(), torch.cat(ks).contiguous(), torch.cat(cs).contiguous()


def test_merge_of_target_shards_equals_unsplit():
    from spacepharer_b200 import api, multigpu
    tres, toff, qres, qoff = _workload()
    Q = len(qoff) - 1
    s = api.Searcher(device=0)
    try:
        s.build_index(tres, toff)
        off0, hits0, _ = s.prefilter(qres, qoff)
        G = 3
        recv, keys, cnt = _shard_lists_on_one_gpu(s, tres, toff, qres, qoff, G, 300)
        s.set_queries(qres, qoff)
        s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), G)
        off1, hits1, _ = s.prefilter_fetch()
        assert (off1 == off0).all()
        assert (hits1["id"] == hits0["id"]).all() and (hits1["score"] == hits0["score"]).all() and (hits1["diag"] == hits0["diag"]).all()
        # numpy restatement of the merge agrees too
        moff, mh = multigpu.merge_hit_lists_reference(recv.cpu().numpy(), cnt.cpu().numpy(), G, Q)
        assert (moff == off0.astype(np.int64)).all() and (mh[:, 0] == hits0["id"]).all()
        # without the shards' truncation keys, reaching max-seqs across shards is reported, not silently truncated
        with pytest.raises(Exception, match="max-seqs"):
            s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), G, max_seqs=2)
    finally:
        s.close()


@pytest.mark.parametrize("G,max_seqs,bin_size", [(3, 4, 2), (2, 10, 8), (5, 2, 2), (4, 7, 2048), (8, 40, 2)])
def test_merge_cuts_at_max_seqs_like_the_unsplit_list(G, max_seqs, bin_size):
    """lists reaching max-seqs across the shards are cut in the reference's order (capped score, bin, stream order, id):
    the merged result equals the unsplit prefilter, which the parity tests pin against the oracle and the golden lists"""
    from spacepharer_b200 import api, multigpu
    tres, toff, qres, qoff = _family_workload()
    Q = len(qoff) - 1
    s = api.Searcher(device=0)
    try:
        s.build_index(tres, toff)
        off0, hits0, _ = s.prefilter(qres, qoff, max_seqs=max_seqs, bin_size=bin_size)
        assert (np.diff(off0.astype(np.int64)) == max_seqs).sum() > Q // 2        # the cut is exercised
        recv, keys, cnt = _shard_lists_on_one_gpu(s, tres, toff, qres, qoff, G, max_seqs, bin_size)
        s.set_queries(qres, qoff)
        s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), G, max_seqs=max_seqs, recv_keys_ptr=keys.data_ptr())
        off1, hits1, _ = s.prefilter_fetch()
        assert (off1 == off0).all()
        assert (hits1["id"] == hits0["id"]).all() and (hits1["score"] == hits0["score"]).all() and (hits1["diag"] == hits0["diag"]).all()
        moff, mh = multigpu.merge_hit_lists_reference(recv.cpu().numpy(), cnt.cpu().numpy(), G, Q, keys=keys.cpu().numpy(),
                                                      max_seqs=max_seqs)
        assert (moff == off0.astype(np.int64)).all() and (mh[:, 0] == hits0["id"]).all() and (mh[:, 1] == hits0["score"]).all()
    finally:
        s.close()


def test_merge_family_truncation_golden():
    """the reference's own truncated lists (tests/golden/family: hundreds of near-identical targets per query) are
    reproduced from three target shards"""
    import common as Cm
    from spacepharer_b200 import api
    T, Qs = Cm.load_fasta_set("family")
    tres, toff = O.concat(T)
    qres, qoff = O.concat(Qs)
    Q = len(Qs)
    s = api.Searcher(device=0)
    try:
        tres, _ = api.tantan_mask(s.seed, tres, toff)        # --mask 1 like the reference run that wrote the fixtures
        for fname, max_seqs in (("pref_max300.txt.gz", 300), ("pref_max10.txt.gz", 10)):
            recv, keys, cnt = _shard_lists_on_one_gpu(s, tres, toff, qres, qoff, 3, max_seqs)
            s.set_queries(qres, qoff)
            s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), 3, max_seqs=max_seqs, recv_keys_ptr=keys.data_ptr())
            off, hits, _ = s.prefilter_fetch()
            ref = Cm.dump("family", fname)
            for i in range(Q):
                got = api.format_prefilter_rows(hits[int(off[i]):int(off[i + 1])])
                assert got == ref.get(i, []), (fname, i)
    finally:
        s.close()


def _rank_main(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    from spacepharer_b200 import multigpu
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    tres, toff, qres, qoff = _workload()
    eng = multigpu.ShardedSearcher(device=rank, rank=rank, world=world, split_mode=mode)
    eng.load_targets(tres, tres, toff)
    res = eng.search_host(qres, qoff, int(toff[-1]))
    np.savez(os.path.join(out_dir, "%s_r%d.npz" % (mode, rank)), rows=res["rows"], cigar=np.frombuffer(res["cigar"], dtype=np.uint8))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_index_split_on_one_gpu(world):
    """Index split (k-mer ranges of the table held by different contexts, gathered through mapped pointers): `world` contexts
    on this GPU each keep one range and attach the others; every one of them returns the unsplit prefilter lists and rows.
    (Same code path as across GPUs -- there the mapped pointers are NVLink peer memory.)"""
    from spacepharer_b200 import api
    tres, toff, qres, qoff = _workload()
    ref = api.Searcher(device=0)
    ref.build_index(tres, toff)
    ref.set_align_targets(tres, toff)
    off0, hits0, st0 = ref.prefilter(qres, qoff)
    _, rows0, cig0, _, _ = ref.search(qres, qoff, int(toff[-1]))
    ref.close()
    ctxs = [api.Searcher(device=0) for _ in range(world)]
    try:
        for r, s in enumerate(ctxs):
            s.build_index(tres, toff)
            s.shard_index(r, world)
        for r, s in enumerate(ctxs):
            for p, o in enumerate(ctxs):
                if p != r:
                    s.shard_attach(p, o)
        with pytest.raises(Exception):
            ctxs[0].export_index()          # the CSR view is gone once the table is sharded
        for r, s in enumerate(ctxs):
            off, hits, st = s.prefilter(qres, qoff)
            assert (off == off0).all() and (hits["id"] == hits0["id"]).all() and (hits["score"] == hits0["score"]).all()
            assert (hits["diag"] == hits0["diag"]).all() and st["n_db_matches"] == st0["n_db_matches"]
        # a lane of a sharded context inherits the shard map; max-seqs small enough to reach the fallback kernels as well
        ln = ctxs[1].lane()
        ctxs[1].set_align_targets(tres, toff)
        _, rows, cig, _, _ = ctxs[1].search(qres, qoff, int(toff[-1]))
        assert len(rows) == len(rows0) and (rows["target"] == rows0["target"]).all() and (rows["score"] == rows0["score"]).all()
        o5, h5, _ = ln.prefilter(qres, qoff, max_seqs=5)
        ln.close()
        ref = api.Searcher(device=0)
        ref.build_index(tres, toff)
        o5r, h5r, _ = ref.prefilter(qres, qoff, max_seqs=5)
        ref.close()
        assert (o5 == o5r).all() and (h5["id"] == h5r["id"]).all() and (h5["score"] == h5r["score"]).all()
    finally:
        for s in ctxs:
            s.close()


@pytest.mark.parametrize("mode", ["target", "query", "index"])
def test_two_gpus_equal_one(tmp_path, mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from spacepharer_b200 import api
    tres, toff, qres, qoff = _workload()
    s = api.Searcher(device=0)
    s.build_index(tres, toff)
    s.set_align_targets(tres, toff)
    _, rows0, cig0, _, _ = s.search(qres, qoff, int(toff[-1]))
    s.close()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_rank_main, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(str(tmp_path / ("%s_r%d.npz" % (mode, r)))) for r in range(2)]
    rows = np.concatenate([p["rows"] for p in parts])
    assert len(rows) == len(rows0)
    for f in ("query", "target", "score", "bits", "q_start", "q_end", "t_start", "t_end", "n_ident", "aln_len", "cigar_len"):
        assert (rows[f] == rows0[f]).all(), f
    assert np.allclose(rows["eval"], rows0["eval"], rtol=1e-12)


def _rank_main_config2(rank, world, port, out_dir):
    import pickle
    import torch
    import torch.distributed as dist
    from spacepharer_b200 import api, multigpu, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    tres, toff, _ = synth.make_targets(n_genomes=8000)
    qres, qoff, _ = synth.make_query_frames(10000, tres, toff)
    eng = multigpu.ShardedSearcher(device=rank, rank=rank, world=world, split_mode="target")
    masked, _ = api.tantan_mask(eng.searcher.seed, tres, toff)
    eng.load_targets(masked, tres, toff)
    res = eng.search_host(qres, qoff, int(toff[-1]))
    rows = res["rows"]
    lines = api.format_align_rows(rows, res["cigar"], np.diff(qoff).astype(np.int64), np.diff(toff).astype(np.int64))
    with open(os.path.join(out_dir, "cfg2_r%d.pkl" % rank), "wb") as f:
        pickle.dump((rows["query"].astype(np.int64), lines), f)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def test_target_split_config2_equals_the_reference_binary(tmp_path):
    """BASELINE config 2 with the targets sharded over two GPUs (index per shard, hit lists exchanged over NCCL, merged on the
    owning GPU, aligned there): the union of the ranks' alignment rows is, byte for byte, the UNSPLIT reference binary's
    alignment DB (tests/golden/config2_digest.json)"""
    import hashlib
    import json
    import pickle
    import torch
    import torch.multiprocessing as mp
    from conftest import GOLDEN
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    want = json.load(open(os.path.join(GOLDEN, "config2_digest.json")))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_rank_main_config2, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    entries = {}
    n_rows = 0
    for r in range(2):
        qs, lines = pickle.load(open(str(tmp_path / ("cfg2_r%d.pkl" % r)), "rb"))
        for q, line in zip(qs, lines):
            entries.setdefault(int(q), []).append(line)
        n_rows += len(lines)
    h = hashlib.sha256()
    for k in range(want["n_queries"]):
        h.update(b"%d\n" % k)
        h.update("".join(x + "\n" for x in entries.get(k, [])).encode())
    assert n_rows == want["aln_rows"]
    assert h.hexdigest() == want["aln_sha256"]
