"""The oracle against the compiled reference on lattices of the bench's shape (synthetic dictionary and
corpus, Jumandic beam 5 + RNN re-ranking with layer size 128) -- beyond the committed golden files, and
regenerated on every run.  Skipped where oracle/_ref is not built."""
import argparse
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
import parity

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.mark.parametrize("workload,seed,n", [("jumandic_beam5", 23, 300), ("jumandic_beam5", 977, 200),
                                             ("jumandic_beam20", 101, 120), ("jumandic_beam5_linear", 5, 300)])
def test_oracle_equals_reference_on_a_fresh_bench_like_sample(tmp_path, workload, seed, n):
    import bench
    from jumanpp_b200 import abi, jppb
    if not bench.HARNESS.exists():
        pytest.skip("oracle/_ref/jpp_ref_harness is not built")
    args = argparse.Namespace(workload=workload, cache=str(tmp_path), dict_rows=20000, sentences=n, min_len=4,
                              max_len=96, seed=seed, log2w=20, wseed=5, stage_chunk=2048, gpus=1, batch_form="inline", window=131072,
                              shard_corpus=False)
    info = bench.prepare_workload(args, 0)
    model, cfg, windows, _ = bench.load_staged(info, args, pin=lambda a: a)
    batch = windows[0]
    out = info["dir"] / "expected.jppb"
    cmd = [str(bench.HARNESS), "stage", f"dict={info['dict']}", f"corpus={info['corpus']}", f"out={out}", "expected=1",
           f"max_sent={n}", *bench.harness_model_args(args, info["wl"], info["rnn"])]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = jppb.load(out)
    for b in (batch, batch.compact_entries()):
        ora, _ = oracle_lib.oracle_score(model, cfg, b)
        # parity.py: patterns / links / linear scores bit-equal everywhere; totals and RNN scores bit-equal
        # on the elements of the final paths (the reference leaves the RNN slot of the others uninitialised)
        rnn = info["wl"]["rnn"]
        rep = parity.compare_with_expected(d, b, ora, rnn=rnn)
        if rnn:
            assert rep.pop("final_path_elements") > 500
        assert all(v == 0 for v in rep.values()), rep
