"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch by
codepoints, score their shards independently (no data-path collective), and
all_reduce their counters; stitched shard results equal the whole-batch result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle_lib
    from conftest import load_golden
    from jumanpp_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, model, cfg, batch = load_golden("default")
    sub, (lo, hi) = shard.shard_of(batch, rank, world)
    res, st = oracle_lib.oracle_score(model, cfg, sub)      # stand-in scorer: no GPU on this box
    tot = shard.reduce_counters({"sentences": sub.num_sentences, "nodes": sub.num_nodes, "gathers": st.gathers}, dist)
    np.save(Path(out_dir) / f"beam_{rank}.npy", res.node_beam.view(np.uint8))
    # BASELINE configs[4]: the shards' 1-best paths back on rank 0 in corpus order (bench.py --shard-corpus)
    from jumanpp_b200 import abi
    paths = abi.best_paths(sub, res)
    res.best_path_len = np.array([len(p) for p in paths], np.uint32)
    res.best_path_nodes = np.full(sub.num_boundaries, 0xFFFFFFFF, np.uint32)
    for s_, p_ in enumerate(paths):
        b0 = int(sub.sent_boundary_off[s_])
        for k_, (bnd, pos) in enumerate(p_):
            res.best_path_nodes[b0 + k_] = int(sub.boundary_node_off[b0 + bnd]) + pos
    lens, nodes = shard.relative_paths([sub], [res])
    g = shard.gather_in_corpus_order(lens, nodes, lo, hi, batch.num_sentences, dist, "cpu")
    if rank == 0:
        assert g["in_corpus_order"] and g["sentences"] == batch.num_sentences
        np.save(Path(out_dir) / "gathered_lens.npy", g["lens"])
        np.save(Path(out_dir) / "gathered_nodes.npy", g["nodes"])
    else:
        assert g is None
    np.save(Path(out_dir) / f"meta_{rank}.npy", np.array([lo, hi, tot["sentences"], tot["nodes"], tot["gathers"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    from conftest import load_golden
    d, model, cfg, batch = load_golden("default")
    whole, st = oracle_lib.oracle_score(model, cfg, batch)
    m0, m1 = np.load(tmp_path / "meta_0.npy"), np.load(tmp_path / "meta_1.npy")
    assert m0[0] == 0 and m0[1] == m1[0] and m1[1] == batch.num_sentences and 0 < m0[1] < batch.num_sentences
    assert m0[2] == m1[2] == batch.num_sentences and m0[3] == m1[3] == batch.num_nodes
    assert m0[4] == m1[4] == st.gathers
    stitched = np.concatenate([np.load(tmp_path / "beam_0.npy"), np.load(tmp_path / "beam_1.npy")])
    assert stitched.tobytes() == whole.node_beam.view(np.uint8).tobytes()
    # the gathered paths are the whole batch's paths, in corpus order
    from jumanpp_b200 import abi
    paths = abi.best_paths(batch, whole)
    want_len = np.array([len(p) for p in paths], np.int64)
    want_nodes = np.array([int(batch.boundary_node_off[int(batch.sent_boundary_off[s]) + bnd]) + pos
                           - int(batch.boundary_node_off[int(batch.sent_boundary_off[s])])
                           for s, p in enumerate(paths) for bnd, pos in p], np.int64)
    assert np.array_equal(np.load(tmp_path / "gathered_lens.npy"), want_len)
    assert np.array_equal(np.load(tmp_path / "gathered_nodes.npy"), want_nodes)


def test_shard_bounds_balance_codepoints():
    from jumanpp_b200 import shard
    rng = np.random.RandomState(3)
    cps = rng.randint(8, 129, size=10000)
    for world in (1, 2, 4, 8):
        b = shard.shard_bounds(cps, world)
        assert b[0] == 0 and b[-1] == len(cps) and np.all(np.diff(b) >= 0)
        loads = [cps[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) - min(loads) <= 2 * 128
