"""N > 1 host logic on CPU: two gloo ranks each answer their shard of one batch (CPU oracle
standing in for the per-rank engine), rank 0 gathers; result must equal the single-process
answer, shards must be disjoint and complete, timings reduce with max."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harness import dist as hd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rle, queries, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    assert hd.env_rank_world() == (rank, world, rank)
    b = O.RleBWT()
    b.load_vector(rle)
    lo, hi = hd.shard_bounds(len(queries), rank, world)
    local = b.count_kmers_fixed(queries[lo:hi], k).astype(np.int64)
    dist.barrier()
    t = hd.max_over_ranks(1.0 + rank)
    total = hd.sum_over_ranks(hi - lo)
    full = hd.gather_slices(torch.from_numpy(local), len(queries))
    if rank == 0:
        ret["counts"] = full.numpy()
        ret["t"] = t
        ret["total"] = total
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from harness import bwt_build, synth
    from oracle import oracle as O
    reads = synth.np_make_reads(400, 60, 12.0, 0.01, seed=3)
    rle, _ = bwt_build.build_rle_bwt(torch.from_numpy(reads))
    rle = rle.numpy()
    k = 17
    q = synth.np_make_queries(reads, k, 501, 500, seed=4)
    b = O.RleBWT()
    b.load_vector(rle)
    want = b.count_kmers_fixed(q, k).astype(np.int64)
    for world in (2, 3):
        bounds = [hd.shard_bounds(len(q), r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == len(q)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), rle, q, k, ret), nprocs=2, join=True)
    assert (ret["counts"] == want).all()
    assert ret["t"] == 2.0 and ret["total"] == len(q)
