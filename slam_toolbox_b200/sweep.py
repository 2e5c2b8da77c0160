"""Host-side logic of the multi-GPU loop-closure sweep: how candidate chains are
sharded over ranks and how the per-query best response is reduced.

Candidates shard naturally (every (query, chain) pair is independent), so the data path has no
collective.  The exchange of round 2 is `gather_winners` below: winner records built on the device
(b200sm_batch_winner_records), ONE all-gather, local selection (b200sm_batch_winners_select) -- no host round trip,
valid for penalised sweeps too.  The key-only exchange of round 1 is kept for callers that need just the winning id:
ONE all-reduce(MAX) of a packed 64-bit key per query:

    key = (best integer correlation sum << 32) | (0xFFFFFFFF - global candidate id)

MAX picks the highest sum and, between equal sums, the LOWEST candidate id -- deterministic for any
rank count.  The same packing is produced on the device by b200sm_batch_reduce_keys.

The winners' results follow in a second, equally small exchange: the rank that owns the winning candidate
of a query contributes its (response, mean[3], cov[9]) row, every other rank contributes zeros, and one
all-reduce(SUM) of the [Q, 13] table leaves the row on every rank -- x + 0 is exact, so the values are the
owner's bits.  Per-candidate results stay on their owners (TryCloseLoop consumes every passing chain,
Mapper.cpp:1500-1561) and come back through b200sm_batch_fetch.
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


def winners_payload(reduced_keys: np.ndarray, lo: int, hi: int, pair_query: np.ndarray, pair_global_chain: np.ndarray,
                    response: np.ndarray, mean: np.ndarray, cov: np.ndarray) -> np.ndarray:
    """[Q, 13] float64 table for the winners exchange: row q = (response, mean[3], cov[9]) of the pair (q, winning
    candidate) if this rank owns that candidate (lo <= id < hi), else zeros.  reduced_keys = keys after all-reduce."""
    _, gid = unpack_keys(reduced_keys)
    nq = len(gid)
    out = np.zeros((nq, 13), dtype=np.float64)
    pq = np.asarray(pair_query, dtype=np.int64)
    pc = np.asarray(pair_global_chain, dtype=np.int64)
    cov = np.asarray(cov, dtype=np.float64).reshape(len(pq), 9)
    own = (gid >= lo) & (gid < hi) & (np.asarray(reduced_keys) != 0)
    if own.any():
        hit = np.nonzero(own[pq] & (pc == gid[pq]))[0]     # the pairs that are some query's winner
        q = pq[hit]
        out[q, 0] = np.asarray(response, dtype=np.float64)[hit]
        out[q, 1:4] = np.asarray(mean, dtype=np.float64)[hit]
        out[q, 4:13] = cov[hit]
    return out


def allreduce_winners(table_tensor):
    """In-place all-reduce(SUM) of the [Q, 13] winners table (torch float64) over the default process group."""
    import torch.distributed as dist
    dist.all_reduce(table_tensor, op=dist.ReduceOp.SUM)
    return table_tensor


class WinnerExchange:
    """Device buffers + the three steps of the multi-GPU winner exchange for one ScanMatcher handle:
    records (device kernels) -> all_gather_into_tensor (NCCL over NVLink; a plain copy for world size 1) -> select."""

    def __init__(self, matcher, n_queries: int, world: int):
        import torch
        self.sm, self.nq, self.world = matcher, n_queries, world
        nbytes = n_queries * type(matcher).winner_record_bytes()
        self.send = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        self.recv = torch.zeros(world * nbytes, dtype=torch.uint8, device="cuda")

    def gather(self, id_offset: int):
        """After batch_run: this rank's records, then the collective (asynchronous on the current stream)."""
        self.sm.batch_winner_records(self.send.data_ptr(), id_offset)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.recv, self.send)
        else:
            self.recv.copy_(self.send)

    def select(self):
        """(global ids, response, mean[Q,3], cov[Q,3,3]) of every query's winner -- identical on every rank."""
        return self.sm.batch_winners_select(self.recv.data_ptr(), self.world, self.nq)
