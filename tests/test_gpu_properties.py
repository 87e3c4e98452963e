"""GPU properties at a size the golden files do not reach: a few thousand synthetic sentences
(lattices built by the compiled reference's host pipeline, oracle/_ref), RNN re-ranking on,
layer size 128 as in the bench.  Checked: idempotence, independence of sentence order and of
how the batch is cut (windows, lanes, compact rows), structural invariants of the beams and
of the 1-best paths, the oracle on a bounded prefix and the reference's own post-state."""
import argparse
import dataclasses
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu

N_SENT = 6000


@pytest.fixture(scope="module")
def workload(tmp_path_factory):
    import bench
    from jumanpp_b200 import abi, jppb
    if not bench.HARNESS.exists():
        pytest.skip("oracle/_ref/jpp_ref_harness is not built (it ships with the repo snapshot to the GPU box)")
    cache = tmp_path_factory.mktemp("wl")
    args = argparse.Namespace(workload="jumandic_beam5", cache=str(cache), dict_rows=30000, sentences=N_SENT, min_len=4,
                              max_len=96, seed=11, log2w=22, wseed=3, stage_chunk=2048, gpus=1, batch_form="inline",
                              shard_corpus=False, window=1 << 30)
    info = bench.prepare_workload(args, 0)
    model, cfg, windows, _ = bench.load_staged(info, args, pin=lambda a: a)
    batch = windows[0]
    return dict(args=args, info=info, model=model, cfg=cfg, batch=batch, bench=bench, abi=abi, jppb=jppb)


@pytest.fixture(scope="module")
def scored(workload):
    from jumanpp_b200.scorer import LatticeScorer
    sc = LatticeScorer(workload["model"], workload["cfg"])
    res = sc.compute_scores(workload["batch"])
    again = sc.compute_scores(workload["batch"])
    sc.close()
    return res, again


def per_sentence_nodes(batch):
    sbo, bno = batch.sent_boundary_off, batch.boundary_node_off
    return [(int(bno[int(sbo[s])]), int(bno[int(sbo[s + 1])])) for s in range(batch.num_sentences)]


def test_idempotent(workload, scored):
    res, again = scored
    assert res.node_beam.tobytes() == again.node_beam.tobytes()
    assert res.node_beam_scores.tobytes() == again.node_beam_scores.tobytes()
    assert np.array_equal(res.best_path_nodes, again.best_path_nodes)


def test_independent_of_sentence_order_and_cut(workload, scored, monkeypatch):
    """Sentences are independent: reversed order, three separate sub-batches, lanes and compact rows all
    give the bytes of the whole-batch run, sentence by sentence."""
    from jumanpp_b200.scorer import LatticeScorer
    abi, batch, res = workload["abi"], workload["batch"], scored[0]
    ns = batch.num_sentences
    spans = per_sentence_nodes(batch)
    sc = LatticeScorer(workload["model"], workload["cfg"])
    try:
        rev = abi.StagedBatch.concat([batch.select(s, s + 1) for s in range(ns - 1, ns - 401, -1)])
        r = sc.compute_scores(rev)
        off = 0
        for s in range(ns - 1, ns - 401, -1):
            a, b = spans[s]
            assert r.node_beam[off:off + b - a].tobytes() == res.node_beam[a:b].tobytes(), s
            assert r.node_beam_scores[off:off + b - a].tobytes() == res.node_beam_scores[a:b].tobytes(), s
            off += b - a
        for lo, hi in ((0, ns // 3), (ns // 3, ns // 2), (ns // 2, ns)):
            r = sc.compute_scores(batch.select(lo, hi))
            a, b = spans[lo][0], spans[hi - 1][1]
            assert r.node_beam.tobytes() == res.node_beam[a:b].tobytes()
        monkeypatch.setenv("JPP_B200_LANE_MIN_SENTENCES", "500")
        for lanes in ("1", "2", "4"):
            monkeypatch.setenv("JPP_B200_LANES", lanes)
            for b in (batch, batch.compact_entries()):
                r = sc.compute_scores(b)
                assert r.node_beam.tobytes() == res.node_beam.tobytes(), lanes
                assert r.node_beam_scores.tobytes() == res.node_beam_scores.tobytes(), lanes
                assert np.array_equal(r.best_path_len, res.best_path_len)
                assert np.array_equal(r.best_path_nodes, res.best_path_nodes)
    finally:
        sc.close()


def test_beam_and_path_invariants(workload, scored):
    abi, batch, cfg, res = workload["abi"], workload["batch"], workload["cfg"], scored[0]
    beam = res.node_beam
    fake = beam["left"] == 0xFFFF
    # fake slots only pad the tail of a row, and carry the lowest float
    assert not (fake[:, :-1] & ~fake[:, 1:]).any()
    assert (beam["prev_beam"][fake] == 0xFFFF).all()
    # live slots: links stay inside the boundary's `ends` list and the previous node's beam
    sbo, bno, beo = batch.sent_boundary_off, batch.boundary_node_off, batch.boundary_end_off
    nodes_per_bnd = np.diff(bno.astype(np.int64))
    node_bnd = np.repeat(np.arange(len(nodes_per_bnd)), nodes_per_bnd)
    n_left = (beo[1:].astype(np.int64) - beo[:-1].astype(np.int64))[node_bnd]
    bnd_sent = np.repeat(np.arange(batch.num_sentences), np.diff(sbo.astype(np.int64)))
    local_bnd = (np.arange(len(bnd_sent)) - sbo[bnd_sent].astype(np.int64))[node_bnd]
    live = ~fake & (local_bnd >= 2)[:, None]          # the two BOS rows hold the seed element
    assert (beam["left"][live] < np.repeat(n_left[:, None], cfg.beam_size, 1)[live]).all()
    assert (beam["prev_beam"][live] < cfg.beam_size).all()
    # the EOS row is the re-ranked beam: totals (linear + weighted RNN) in descending order
    for s in range(0, batch.num_sentences, 37):
        eos = int(bno[int(sbo[s + 1]) - 1])
        row = beam[eos]
        t = row["total_score"][row["left"] != 0xFFFF]
        assert (np.diff(t) <= 0).all(), s
    # 1-best paths: device-resolved == host walk over the beams; consecutive nodes touch;
    # the path covers the sentence from codepoint 0 to its end
    walked = abi.best_paths(batch, res)
    got = res.paths(batch)
    ni = batch.node_info
    ncp = np.diff(batch.sent_codepoint_off.astype(np.int64))
    for s in range(batch.num_sentences):
        b0 = int(sbo[s])
        want = [int(bno[b0 + b]) + p for b, p in walked[s][:-1]]
        assert got[s] == want, s
        if not want:
            continue
        assert int(ni["start"][want[0]]) == 0
        for a, b in zip(want, want[1:]):
            assert int(ni["end"][a]) == int(ni["start"][b])
        assert int(ni["end"][want[-1]]) == int(ncp[s])


def test_oracle_on_a_prefix(workload, scored):
    batch, res = workload["batch"], scored[0]
    n = 300
    sub = batch.select(0, n)
    hi = per_sentence_nodes(batch)[n - 1][1]
    ora, _ = oracle_lib.oracle_score(workload["model"], workload["cfg"], sub)
    assert np.array_equal(res.node_patterns[:hi], ora.node_patterns)
    assert res.node_t0_score[:hi].tobytes() == ora.node_t0_score.tobytes()
    assert np.array_equal(res.node_beam["left"][:hi], ora.node_beam["left"])
    assert np.array_equal(res.node_beam["prev_beam"][:hi], ora.node_beam["prev_beam"])
    # linear connection scores bit-exact; totals of the EOS rows carry the RNN score (1e-4 relative)
    live = ora.node_beam["left"] != 0xFFFF
    assert res.node_beam_scores[:hi, :, 0][live].tobytes() == ora.node_beam_scores[:, :, 0][live].tobytes()
    a, b = res.node_beam["total_score"][:hi][live], ora.node_beam["total_score"][live]
    assert np.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_reference_post_state(workload, scored):
    """The compiled reference's own beams for the first sentences of the same corpus."""
    bench, abi, jppb = workload["bench"], workload["abi"], workload["jppb"]
    args, info, batch, res = workload["args"], workload["info"], workload["batch"], scored[0]
    n = 400
    out = info["dir"] / f"expected_{n}.jppb"
    cmd = [str(bench.HARNESS), "stage", f"dict={info['dict']}", f"corpus={info['corpus']}", f"out={out}", "expected=1",
           f"max_sent={n}", *bench.harness_model_args(args, info["wl"], info["rnn"])]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = jppb.load(out)
    hi = per_sentence_nodes(batch)[n - 1][1]
    assert np.array_equal(d["exp_patterns"], res.node_patterns[:hi])
    assert np.array_equal(d["exp_beam_lp"][:, :, 0], res.node_beam["left"][:hi])
    assert np.array_equal(d["exp_beam_lp"][:, :, 1], res.node_beam["prev_beam"][:hi])
    live = d["exp_beam_lp"][:, :, 0] != 0xFFFF
    assert d["exp_beam_scores"][:, :, 0][live].tobytes() == res.node_beam_scores[:hi, :, 0][live].tobytes()
    sub = batch.select(0, n)
    exp_beam = np.zeros((hi, res.node_beam.shape[1]), abi.BEAM_EL_DT)
    exp_beam["left"], exp_beam["prev_beam"] = d["exp_beam_lp"][:, :, 0], d["exp_beam_lp"][:, :, 1]
    exp_beam["total_score"] = d["exp_beam_total"]
    exp = abi.ScoreResult(exp_beam, d["exp_beam_scores"], d["exp_patterns"], res.node_t0_score[:hi])
    mine = dataclasses.replace(res, node_beam=res.node_beam[:hi])
    assert abi.best_paths(sub, mine) == abi.best_paths(sub, exp)
