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


def relative_paths(windows, results) -> tuple[np.ndarray, np.ndarray]:
    """1-best paths of a rank's windows as (path length per sentence, path nodes relative to the first
    node of their sentence): comparable across ranks, whatever the windowing."""
    lens = np.concatenate([r.best_path_len for r in results]).astype(np.int64)
    rel = []
    for w, r in zip(windows, results):
        sbo, bno = w.sent_boundary_off.astype(np.int64), w.boundary_node_off.astype(np.int64)
        first_node = bno[sbo[:-1]]
        for s in range(w.num_sentences):
            k = int(r.best_path_len[s])
            rel.append(r.best_path_nodes[int(sbo[s]):int(sbo[s]) + k].astype(np.int64) - first_node[s])
    nodes = np.concatenate(rel) if rel else np.zeros(0, np.int64)
    return lens, nodes


def gather_in_corpus_order(lens: np.ndarray, nodes: np.ndarray, lo: int, hi: int, total: int, dist, device) -> dict | None:
    """BASELINE configs[4]: the 1-best paths of every shard back on rank 0 in corpus order (one gather of
    the per-sentence path lengths and one of the path nodes: NCCL on the GPU box, gloo in the CPU test).
    Rank 0 checks that the pieces cover sentences [0, total) without gap or overlap and returns a digest
    of the stitched result; the other ranks return None."""
    import hashlib

    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = torch.tensor([len(lens), len(nodes), lo, hi], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().numpy() for m in metas]
    max_l, max_n = max(int(m[0]) for m in metas), max(int(m[1]) for m in metas)
    tl = torch.zeros(max(1, max_l), dtype=torch.int64, device=device)
    tl[:len(lens)] = torch.from_numpy(np.ascontiguousarray(lens)).to(device)
    tn = torch.zeros(max(1, max_n), dtype=torch.int64, device=device)
    tn[:len(nodes)] = torch.from_numpy(np.ascontiguousarray(nodes)).to(device)
    gl = [torch.zeros_like(tl) for _ in range(world)] if rank == 0 else None
    gn = [torch.zeros_like(tn) for _ in range(world)] if rank == 0 else None
    dist.gather(tl, gl, dst=0)
    dist.gather(tn, gn, dst=0)
    if rank != 0:
        return None
    cursor, ok = 0, True
    all_l, all_n = [], []
    for r in range(world):
        nl, nn, rlo, rhi = (int(x) for x in metas[r])
        ok = ok and rlo == cursor and rhi - rlo == nl
        cursor = rhi
        all_l.append(gl[r][:nl].cpu().numpy())
        all_n.append(gn[r][:nn].cpu().numpy())
    ok = ok and cursor == total
    all_l, all_n = np.concatenate(all_l), np.concatenate(all_n)
    ok = ok and int(all_l.sum()) == len(all_n)
    return {"in_corpus_order": bool(ok), "sentences": int(len(all_l)), "path_nodes": int(len(all_n)),
            "shard_sentences": [int(m[0]) for m in metas],
            "paths_sha256_16": hashlib.sha256(all_l.tobytes() + all_n.tobytes()).hexdigest()[:16],
            "lens": all_l, "nodes": all_n}
