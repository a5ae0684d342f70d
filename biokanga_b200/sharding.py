"""Read sharding across GPUs (one process per GPU under torch.distributed).

The align path shards by reads with no exchange step: every rank holds a full copy of
the index and aligns one contiguous range of reads; ranges start on even read numbers so PE1/PE2 of a pair
stay together (the reference relies on that adjacency, Aligner.cpp:9685).  The only collective is the `sum`
all-reduce of the counter vector (`bkx_align_stats` / `bkx_pe_stats`); records are concatenated in read order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


def shard_range(n_reads: int, rank: int, world: int):
    """[begin, end) of the reads rank `rank` aligns: contiguous, even-sized except possibly the last."""
    per = -(-n_reads // world)
    per += per & 1
    b = min(n_reads, per * rank)
    e = min(n_reads, per * (rank + 1))
    return b, e


def stats_to_array(stats) -> np.ndarray:
    return np.frombuffer(bytes(stats), dtype=np.uint64).astype(np.int64)


def array_to_stats(arr, cls=abi.AlignStats):
    raw = np.ascontiguousarray(arr, dtype=np.int64).astype(np.uint64).tobytes()
    return cls.from_buffer_copy(raw)


def all_reduce_stats(stats, dist, device=None):
    """Sum a bkx_align_stats / bkx_pe_stats over all ranks (NCCL on GPUs, gloo on CPU)."""
    import torch
    t = torch.from_numpy(stats_to_array(stats))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return array_to_stats(t.cpu().numpy(), type(stats))


def gather_results(local: np.ndarray, n_reads: int, rank: int, world: int, dist):
    """Concatenate per-rank record arrays in read order on every rank."""
    import torch
    out = np.zeros(n_reads, dtype=abi.RESULT_DTYPE)
    per = shard_range(n_reads, 0, world)[1]
    buf = np.zeros(per, dtype=abi.RESULT_DTYPE)
    buf[:len(local)] = local
    t = torch.from_numpy(buf.view(np.uint8).copy())
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    for r in range(world):
        b, e = shard_range(n_reads, r, world)
        out[b:e] = parts[r].numpy().view(abi.RESULT_DTYPE)[:e - b]
    return out
