"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path. Each rank factorises its
block of the mats axis (the CPU oracle stands in for the single-GPU launcher, which is identical on every
rank), results are all-gathered and must be bit-identical to the unsharded run because shards are independent."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, n, out_dir):
    sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_np as oracle
    from gputils_b200 import sharding
    A = oracle.fill_spd_batched(n, k, float(n), 123)                 # every rank can regenerate any shard
    b = oracle.fill_uniform(k * n, -1.0, 1.0, 124).reshape(k, n, 1)
    lo, hi = sharding.shard_range(k, world, rank)
    L_loc, info_loc = oracle.potrf_batched(A[lo:hi]) if hi > lo else (np.zeros((0, n, n)), np.zeros(0, np.int32))
    x_loc = oracle.potrs_batched(L_loc, b[lo:hi]) if hi > lo else np.zeros((0, n, 1))
    L_all = sharding.all_gather_shards(torch.from_numpy(L_loc), k)
    x_all = sharding.all_gather_shards(torch.from_numpy(x_loc), k)
    info_all = sharding.all_gather_shards(torch.from_numpy(info_loc), k)
    nrm = sharding.global_norm_f(float((x_loc ** 2).sum()))
    mx = sharding.global_max_abs(float(np.abs(x_loc).max()) if hi > lo else 0.0)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), L=L_all.numpy(), x=x_all.numpy(), info=info_all.numpy(), nrm=nrm, mx=mx)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [7, 8, 1])
def test_sharded_run_equals_single_run(tmp_path, oracle, k):
    n, world = 6, 2
    mp.spawn(_worker, args=(world, _free_port(), k, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    A = oracle.fill_spd_batched(n, k, float(n), 123)
    b = oracle.fill_uniform(k * n, -1.0, 1.0, 124).reshape(k, n, 1)
    L, info = oracle.potrf_batched(A)
    x = oracle.potrs_batched(L, b)
    assert np.array_equal(got["L"], L) and np.array_equal(got["x"], x) and np.array_equal(got["info"], info)
    assert abs(float(got["nrm"]) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
    assert float(got["mx"]) == np.abs(x).max()


def test_shard_ranges_cover_the_batch_exactly():
    from gputils_b200 import sharding
    for k in (0, 1, 5, 8, 1_000_000, 1_000_003):
        for world in (1, 2, 4, 8):
            rs = [sharding.shard_range(k, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == k
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert sum(sharding.shard_sizes(k, world)) == k
            assert max(sharding.shard_sizes(k, world)) - min(s for s in sharding.shard_sizes(k, world) if s or k < world) <= max(1, -(-k // world))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
