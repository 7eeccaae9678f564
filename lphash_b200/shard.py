"""Sharding of the hot path over GPUs: the query and the build scan split by contig
range with no data-path exchange — every k-mer's code depends only on its own k bases and the
replicated, read-only index.

    plan = shard.plan(offsets, world_size)       contiguous contig ranges of near-equal base count
    b, o = shard.take(bases, offsets, plan[r])   rank r's batch (views, offsets rebased to 0)
    mm0  = shard.mm_count_starts(offsets, plan, m, mm_count)   starting m-mer ordinal per shard
                                                  (the `mm_count` in/out argument of
                                                  minimizer::from_string, /root/reference/include/
                                                  minimizer.hpp:14,59, carried across shards)
    shard.concat_codes(parts) / concat_records(parts)          results in shard (= input) order

Host logic only; the per-shard work goes through the C ABI (api.Mphf.query_batch,
api.scan_superkmers).  `gather_*` helpers move the per-rank results to rank 0 with
torch.distributed (any backend; tests use gloo with world_size 2).
"""
from __future__ import annotations

import numpy as np


def plan(offsets, world_size: int) -> list[tuple[int, int]]:
    """[(c_begin, c_end)) per rank: contiguous, covering all contigs, cut on contig boundaries at
    the points closest to equal base counts.  Ranks may receive empty ranges when there are fewer
    contigs than ranks."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if n <= 0:
        return [(0, 0)] * world_size
    first, total = int(offsets[0]), int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world_size):
        target = first + (total * r) // world_size
        c = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        # offsets[c] >= target; pick the nearer of the two boundaries around the target
        if c > 0 and target - int(offsets[c - 1]) < int(offsets[min(c, n)]) - target:
            c -= 1
        c = min(max(c, cuts[-1]), n)
        cuts.append(c)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def take(bases, offsets, rng: tuple[int, int]):
    """The batch of contigs [c0, c1): (bases view, offsets rebased to start at 0)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    c0, c1 = rng
    lo, hi = int(offsets[c0]), int(offsets[c1])
    return np.asarray(bases)[lo:hi], (offsets[c0:c1 + 1] - offsets[c0]).astype(np.uint64)


def mm_count_starts(offsets, shards: list[tuple[int, int]], m: int, mm_count: int = 0) -> list[int]:
    """Value of the running m-mer ordinal at the start of every shard: mm_count + number of m-mers
    (L-m+1 for L >= m) of all earlier contigs."""
    lens = np.diff(np.asarray(offsets, dtype=np.uint64)).astype(np.int64)
    mmers = np.maximum(lens - m + 1, 0)
    csum = np.concatenate([[0], np.cumsum(mmers)])
    return [int(mm_count + csum[c0]) for c0, _ in shards]


def concat_codes(parts):
    """parts = [(codes, code_offsets)] in shard order -> (codes, code_offsets) of the whole batch."""
    codes = np.concatenate([np.asarray(p[0], dtype=np.uint64) for p in parts]) if parts else np.empty(0, np.uint64)
    offs = [np.zeros(1, dtype=np.uint64)]
    base = 0
    for c, o in parts:
        o = np.asarray(o, dtype=np.uint64)
        offs.append(o[1:] + np.uint64(base))
        base += int(o[-1])
    return codes, np.concatenate(offs)


def concat_records(parts):
    return np.concatenate(parts) if parts else np.empty(0)


def gather_to_rank0(arr: np.ndarray, dist, rank: int, world: int):
    """Rank 0 receives every rank's 1-D array (in rank order); others return None."""
    import torch
    arr = np.ascontiguousarray(arr)
    raw = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy())
    n = torch.tensor([raw.numel()], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    if rank == 0:
        out = [arr]
        for r in range(1, world):
            buf = torch.empty(int(sizes[r]), dtype=torch.uint8)
            dist.recv(buf, src=r)
            out.append(buf.numpy().view(arr.dtype))
        return out
    dist.send(raw, dst=0)
    return None
