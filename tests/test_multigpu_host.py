"""Host logic of the multi-GPU driver on CPU (gloo, world_size 2): target-shard decomposition, query ranges, the
all-to-all of per-query hit lists and the merge rule.  The device merge (spb_prefilter_merge_shards) is checked
against the same numpy restatement in the GPU tests."""
import os
import socket

import numpy as np
import pytest

from spacepharer_b200 import multigpu


def test_decompose_covers_all_targets():
    rng = np.random.default_rng(1)
    lens = rng.integers(30, 400, 1000)
    toff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 8):
        parts = multigpu.decompose_by_amino_acid(toff, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and sum(s for _, s in parts) == 1000
        for (f0, s0), (f1, _) in zip(parts, parts[1:]):
            assert f0 + s0 == f1
        res = [int(toff[f + s] - toff[f]) for f, s in parts]
        assert max(res) <= (int(toff[-1]) + 2 * 1000) / world + 402      # greedy cut at ceil(total / world)


def test_query_ranges_partition():
    for Q in (0, 1, 7, 38024):
        for world in (1, 2, 8):
            r = multigpu.query_ranges(Q, world)
            assert r[0][0] == 0 and r[-1][1] == Q and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _shard_lists(rank, world, Q, seed=7):
    """deterministic fake per-shard prefilter output: (offsets, hits[id, score, diag]) with global ids ≡ rank mod world"""
    rng = np.random.default_rng(seed * 100 + rank)
    cnt = rng.integers(0, 5, Q)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    n = int(off[-1])
    hits = np.stack([rng.integers(0, 5000, n) * world + rank, rng.integers(15, 120, n), rng.integers(-300, 10, n)], axis=1).astype(np.int32)
    for q in range(Q):     # per query sorted like the prefilter output
        seg = hits[off[q]:off[q + 1]]
        hits[off[q]:off[q + 1]] = seg[np.lexsort((seg[:, 0], -seg[:, 1]))]
    return off, hits


def _worker(rank, world, port, Q, out_dir):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    off, hits = _shard_lists(rank, world, Q)
    ranges = multigpu.query_ranges(Q, world)
    recv, cnt = multigpu.exchange_hit_lists(torch.from_numpy(hits), torch.from_numpy(off), ranges, rank)
    lo, hi = ranges[rank]
    moff, mhits = multigpu.merge_hit_lists_reference(recv.numpy(), cnt.numpy(), world, hi - lo)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), off=moff, hits=mhits, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_merge_world2(tmp_path):
    import torch.multiprocessing as mp
    world, Q = 2, 101
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, Q, str(tmp_path)), nprocs=world, join=True)
    # expected: union of both shards per query, (score desc, id asc)
    shards = [_shard_lists(r, world, Q) for r in range(world)]
    for r in range(world):
        z = np.load(str(tmp_path / ("r%d.npz" % r)))
        lo, hi = int(z["lo"]), int(z["hi"])
        assert (lo, hi) == multigpu.query_ranges(Q, world)[r]
        for q in range(lo, hi):
            rows = np.concatenate([h[o[q]:o[q + 1]] for o, h in shards])
            rows = rows[np.lexsort((rows[:, 0], -rows[:, 1]))]
            got = z["hits"][z["off"][q - lo]:z["off"][q - lo + 1]]
            assert (got == rows).all(), q


def _fake_keys(hits):
    """a shard-independent truncation key per hit: higher score first, then a pseudo stream order derived from the id"""
    return ((255 - np.minimum(hits[:, 1], 255)).astype(np.int64) << 55) | ((hits[:, 0].astype(np.int64) * 7919 % 1000) << 16) | (1 << 62)


def _worker_keys(rank, world, port, Q, max_seqs, out_dir):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    off, hits = _shard_lists(rank, world, Q)
    ranges = multigpu.query_ranges(Q, world)
    recv, cnt, rkeys = multigpu.exchange_hit_lists(torch.from_numpy(hits), torch.from_numpy(off), ranges, rank,
                                                   keys=torch.from_numpy(_fake_keys(hits)))
    lo, hi = ranges[rank]
    assert (rkeys.numpy() == _fake_keys(recv.numpy())).all()        # keys stay aligned with their hits through the exchange
    moff, mhits = multigpu.merge_hit_lists_reference(recv.numpy(), cnt.numpy(), world, hi - lo, keys=rkeys.numpy(), max_seqs=max_seqs)
    np.savez(os.path.join(out_dir, "k%d.npz" % rank), off=moff, hits=mhits, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_with_truncation_keys_world2(tmp_path):
    """the per-hit truncation keys travel with the hits; lists longer than max-seqs keep the first max-seqs in
    (key, id) order and are then put into the final (score desc, id asc) order"""
    import torch.multiprocessing as mp
    world, Q, max_seqs = 2, 101, 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_keys, args=(world, port, Q, max_seqs, str(tmp_path)), nprocs=world, join=True)
    shards = [_shard_lists(r, world, Q) for r in range(world)]
    cut = 0
    for r in range(world):
        z = np.load(str(tmp_path / ("k%d.npz" % r)))
        lo, hi = int(z["lo"]), int(z["hi"])
        for q in range(lo, hi):
            rows = np.concatenate([h[o[q]:o[q + 1]] for o, h in shards])
            if len(rows) > max_seqs:
                rows = rows[np.lexsort((rows[:, 0], _fake_keys(rows)))[:max_seqs]]
                cut += 1
            rows = rows[np.lexsort((rows[:, 0], -rows[:, 1]))]
            got = z["hits"][z["off"][q - lo]:z["off"][q - lo + 1]]
            assert (got == rows).all(), q
    assert cut > 10
