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


def test_shard_bounds_balance_codepoints():
    from jumanpp_b200 import shard
    rng = np.random.RandomState(3)
    cps = rng.randint(8, 129, size=10000)
    for world in (1, 2, 4, 8):
        b = shard.shard_bounds(cps, world)
        assert b[0] == 0 and b[-1] == len(cps) and np.all(np.diff(b) >= 0)
        loads = [cps[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) - min(loads) <= 2 * 128
