"""Sharding of a corpus across ranks (one process per GPU).

Sentences are independent work units (the reference resets its arena per
sentence: analyzer_impl.h:48-55), so the scored path needs no collective: each
rank scores a contiguous block of sentences, balanced by codepoints (the
sequential depth of a sentence), with the model replicated.  NCCL is used only
to (optionally) scatter packed input shards and to sum a handful of counters.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(codepoints_per_sentence: np.ndarray, world: int) -> np.ndarray:
    """[world+1] sentence indices: contiguous shards with (nearly) equal codepoint totals."""
    n = len(codepoints_per_sentence)
    cum = np.concatenate([[0], np.cumsum(codepoints_per_sentence, dtype=np.int64)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left"))
        i = min(max(i, bounds[-1]), n)
        bounds.append(i)
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def shard_of(batch, rank: int, world: int):
    """Sub-batch of `batch` (abi.StagedBatch) owned by `rank`."""
    cps = np.diff(batch.sent_codepoint_off.astype(np.int64))
    b = shard_bounds(cps, world)
    return batch.select(int(b[rank]), int(b[rank + 1])), (int(b[rank]), int(b[rank + 1]))


def reduce_counters(values: dict[str, float], dist=None, device=None) -> dict[str, float]:
    """Sum a small dict of counters over all ranks (all_reduce over NCCL / gloo)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(values)
    import torch
    keys = sorted(values)
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}
