"""Host-side logic of the multi-GPU sweep: sharding, key packing, and the N>1 reduction over gloo
(world_size 2, CPU) exactly as bench.py / the device path performs it over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam_toolbox_b200 import sweep


def test_shards_cover_everything_once():
    for n in (0, 1, 7, 1000, 50001):
        for w in (1, 2, 3, 8):
            rs = [sweep.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_key_order_prefers_sum_then_lowest_id():
    k = sweep.pack_keys(np.array([100, 100, 99, 108100]), np.array([5, 3, 0, 49999]))
    assert k[3] == k.max()
    assert k[1] > k[0] > k[2]
    s, g = sweep.unpack_keys(k)
    assert list(s) == [100, 100, 99, 108100] and list(g) == [5, 3, 0, 49999]
    assert (k > 0).all()


def test_local_best_keys():
    best = np.array([10, 50, 50, 7])
    pq = np.array([0, 1, 1, 0])
    gid = np.array([4, 9, 2, 1])
    keys = sweep.local_best_keys(best, pq, gid, 3)
    s, g = sweep.unpack_keys(keys[:2])
    assert list(s) == [10, 50] and list(g) == [4, 2]
    assert keys[2] == 0


def _worker(rank, world, port, n_queries, n_cand, seed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)                       # same global table on every rank
    sums = rng.integers(0, 60, size=(n_queries, n_cand))    # small range -> plenty of ties
    lo, hi = sweep.shard_range(n_cand, world, rank)
    pq = np.repeat(np.arange(n_queries), hi - lo)
    gid = np.tile(np.arange(lo, hi), n_queries)
    keys = torch.from_numpy(sweep.local_best_keys(sums[:, lo:hi].reshape(-1), pq, gid, n_queries))
    sweep.allreduce_best(keys)
    s, g = sweep.unpack_keys(keys.numpy())
    assert np.array_equal(s, sums.max(axis=1))
    assert np.array_equal(g, sums.argmax(axis=1))           # argmax = first (lowest id) maximum
    # winners exchange: the owner's (response, mean, cov) row reaches every rank unchanged
    full = np.random.default_rng(seed + 1).normal(size=(n_queries, n_cand, 13))
    mine = full[:, lo:hi].reshape(-1, 13)
    tab = torch.from_numpy(sweep.winners_payload(keys.numpy(), lo, hi, pq, gid, mine[:, 0], mine[:, 1:4], mine[:, 4:]))
    assert int((tab.numpy()[:, 0] != 0).sum()) == int(((g >= lo) & (g < hi)).sum())
    sweep.allreduce_winners(tab)
    assert np.array_equal(tab.numpy(), full[np.arange(n_queries), g])
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_best_keys_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 16, 101, 3), nprocs=2, join=True)
