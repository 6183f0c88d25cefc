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


def test_merge_of_target_shards_equals_unsplit():
    import torch
    from spacepharer_b200 import api, multigpu
    tres, toff, qres, qoff = _workload()
    Q = len(qoff) - 1
    s = api.Searcher(device=0)
    try:
        s.build_index(tres, toff)
        off0, hits0, _ = s.prefilter(qres, qoff)
        G = 3
        parts = multigpu.decompose_by_amino_acid(toff, G)
        lists = []
        for frm, size in parts:
            lo, hi = int(toff[frm]), int(toff[frm + size])
            s.build_index(tres[lo:hi], toff[frm:frm + size + 1] - toff[frm])
            off, hits, _ = s.prefilter(qres, qoff, id_base=frm)
            h = np.stack([hits["id"].astype(np.int64) + frm, hits["score"], hits["diag"]], axis=1).astype(np.int32)
            lists.append((off.astype(np.int64), h))
        recv = torch.from_numpy(np.concatenate([h for _, h in lists])).cuda()
        cnt = torch.from_numpy(np.concatenate([np.diff(o) for o, _ in lists]).astype(np.int32)).cuda()
        s.set_queries(qres, qoff)
        s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), G)
        off1, hits1, _ = s.prefilter_fetch()
        assert (off1 == off0).all()
        assert (hits1["id"] == hits0["id"]).all() and (hits1["score"] == hits0["score"]).all() and (hits1["diag"] == hits0["diag"]).all()
        # numpy restatement of the merge agrees too
        moff, mh = multigpu.merge_hit_lists_reference(recv.cpu().numpy(), cnt.cpu().numpy(), G, Q)
        assert (moff == off0.astype(np.int64)).all() and (mh[:, 0] == hits0["id"]).all()
        # reaching max-seqs across shards is reported, not silently truncated
        with pytest.raises(Exception, match="max-seqs"):
            s.merge_shards(0, Q, recv.data_ptr(), recv.shape[0], cnt.data_ptr(), G, max_seqs=2)
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


@pytest.mark.parametrize("mode", ["target", "query"])
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
