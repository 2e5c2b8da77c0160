"""Host-side logic of the multi-GPU loop-closure sweep: how candidate chains are
sharded over ranks and how the per-query best response is reduced.

Candidates shard naturally (every (query, chain) pair is independent), so the data path has no
collective; the only exchange is ONE all-reduce(MAX) of a packed 64-bit key per query:

    key = (best integer correlation sum << 32) | (0xFFFFFFFF - global candidate id)

MAX picks the highest sum and, between equal sums, the LOWEST candidate id -- deterministic for any
rank count.  The same packing is produced on the device by b200sm_batch_reduce_keys.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of candidate chains owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_keys(best_sum: np.ndarray, global_id: np.ndarray) -> np.ndarray:
    s = np.asarray(best_sum, dtype=np.int64)
    g = np.asarray(global_id, dtype=np.int64)
    return (s << 32) | (0xFFFFFFFF - (g & 0xFFFFFFFF))


def unpack_keys(keys: np.ndarray):
    k = np.asarray(keys, dtype=np.int64)
    return (k >> 32).astype(np.int64), (0xFFFFFFFF - (k & 0xFFFFFFFF)).astype(np.int64)


def local_best_keys(best_sum: np.ndarray, pair_query: np.ndarray, pair_global_chain: np.ndarray, n_queries: int) -> np.ndarray:
    """Per-query max key over this rank's pairs (host restatement of k_best_keys); 0 where a query has no pair."""
    keys = pack_keys(best_sum, pair_global_chain)
    out = np.zeros(n_queries, dtype=np.int64)
    np.maximum.at(out, np.asarray(pair_query, dtype=np.int64), keys)
    return out


def allreduce_best(keys_tensor):
    """In-place all-reduce(MAX) of a torch int64 tensor of keys over the default process group."""
    import torch.distributed as dist
    dist.all_reduce(keys_tensor, op=dist.ReduceOp.MAX)
    return keys_tensor
