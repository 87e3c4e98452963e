"""GPU parity: the CUDA path, called through the C ABI (libjpp_b200.so) with
host buffers, against (a) the reference's golden vectors and (b) the CPU oracle.

Bar: patterns, beam links, total scores and perceptron connection scores are
bit-exact.  RNN: every integer decision (tree, beams, EOS order) exact; RNN
float scores within 1e-4 relative of the oracle (CUDA expf vs libm expf), as
BASELINE.json's north star states."""
import numpy as np
import pytest

import oracle_lib
import parity
from conftest import FULL_BEAM, LINEAR_GBEAM, RNN, load_golden

pytestmark = pytest.mark.gpu


def gpu_score(model, cfg, batch):
    from jumanpp_b200.scorer import LatticeScorer
    sc = LatticeScorer(model, cfg)
    try:
        res = sc.compute_scores(batch)
        cnt = sc.counters()
    finally:
        sc.close()
    return res, cnt


@pytest.mark.parametrize("name", LINEAR_GBEAM + FULL_BEAM)
def test_linear_path_matches_reference_and_oracle(name):
    d, model, cfg, batch = load_golden(name)
    res, cnt = gpu_score(model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=False)
    assert rep == {"patterns": 0, "beam_links": 0, "beam_totals": 0, "conn_scores": 0}, rep
    ora, stats = oracle_lib.oracle_score(model, cfg, batch)
    assert res.node_beam.tobytes() == ora.node_beam.tobytes()
    assert np.array_equal(res.node_patterns, ora.node_patterns)
    assert res.node_t0_score.tobytes() == ora.node_t0_score.tobytes()
    assert cnt["gathers"] == stats.gathers
    assert cnt["kernel_ms"]["beam_walk"] > 0
    from jumanpp_b200 import abi
    bno, sbo = batch.boundary_node_off, batch.sent_boundary_off
    want = [[int(bno[int(sbo[s]) + b]) + p for b, p in path[:-1]] for s, path in enumerate(abi.best_paths(batch, res))]
    assert res.paths(batch) == want


@pytest.mark.parametrize("name", RNN)
def test_rnn_path(name, rnn_dir):
    d, model, cfg, batch = load_golden(name, rnn_dir)
    res, cnt = gpu_score(model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=True, rnn_rtol=1e-4)
    n = rep.pop("final_path_elements")
    assert n > 0 and all(v == 0 for v in rep.values()), rep
    # how many are even bit-exact (informational; libm vs CUDA expf)
    exact = parity.compare_with_expected(d, batch, res, rnn=True)
    print(name, "bit-exact mismatches:", exact)


def test_set_config_and_reuse():
    from jumanpp_b200 import abi
    from jumanpp_b200.scorer import LatticeScorer
    d, model, cfg, batch = load_golden("default")
    d2, _, cfg2, batch2 = load_golden("beam1")
    sc = LatticeScorer(model, cfg)
    try:
        r1 = sc.compute_scores(batch)
        r1b = sc.compute_scores(batch)
        assert r1.node_beam.tobytes() == r1b.node_beam.tobytes()
        sc.set_config(abi.ScoringConfig(1, 1, 1, 1))
        r2 = sc.compute_scores(batch)
        ora, _ = oracle_lib.oracle_score(model, abi.ScoringConfig(1, 1, 1, 1), batch)
        assert r2.node_beam.tobytes() == ora.node_beam.tobytes()
    finally:
        sc.close()


def test_batch_split_and_order_invariance():
    d, model, cfg, batch = load_golden("default")
    whole, _ = gpu_score(model, cfg, batch)
    ns = batch.num_sentences
    parts = [batch.select(0, ns // 2), batch.select(ns // 2, ns)]
    off = 0
    for p in parts:
        r, _ = gpu_score(model, cfg, p)
        n = p.num_nodes
        assert r.node_beam.tobytes() == whole.node_beam[off:off + n].tobytes()
        off += n


def test_compact_entry_rows():
    """Entry buffers sent as one row per distinct entry + a row index per node (ABI 2)."""
    import dataclasses
    from jumanpp_b200.scorer import LatticeScorer, ScorerError
    d, model, cfg, batch = load_golden("default")
    inline, _ = gpu_score(model, cfg, batch)
    compact = batch.compact_entries()
    got, _ = gpu_score(model, cfg, compact)
    assert got.node_beam.tobytes() == inline.node_beam.tobytes()
    assert np.array_equal(got.node_patterns, inline.node_patterns)
    bad = dataclasses.replace(compact, node_entry_row=compact.node_entry_row.copy())
    bad.node_entry_row[3] = len(compact.entry_rows) + 5
    sc = LatticeScorer(model, cfg)
    try:
        # a row index past the table fails ITS sentence (ABI 3: jpp_b200_result.sentence_status) ...
        res = sc.compute_scores(bad)
        assert res.sentence_status[0] & abi_mod().SENT_BAD_ENTRY and not res.sentence_status[1:].any()
        # ... and the call, when the caller did not ask for per-sentence status
        blind = sc.allocate_result(bad)
        blind.sentence_status = None
        with pytest.raises(ScorerError) as e:
            sc.compute_scores(bad, blind)
        assert e.value.code == abi_mod().INVALID_PARAMETER
        again = sc.compute_scores(compact)        # the handle stays usable
        assert again.node_beam.tobytes() == inline.node_beam.tobytes()
    finally:
        sc.close()


def test_compact_rows_with_rnn(rnn_dir):
    d, model, cfg, batch = load_golden("rnn_h128", rnn_dir)
    inline, _ = gpu_score(model, cfg, batch)
    compact = batch.compact_entries()
    assert compact.node_rnn_id is None and compact.row_rnn_id is not None
    got, _ = gpu_score(model, cfg, compact)
    assert got.node_beam.tobytes() == inline.node_beam.tobytes()
    assert got.node_beam_scores.tobytes() == inline.node_beam_scores.tobytes()


@pytest.mark.parametrize("name", ["default", "rnn_h128"])
def test_lanes_give_the_same_result(name, rnn_dir, monkeypatch):
    """jpp_b200_score_batch cuts large batches in two halves that run concurrently (two streams, two
    sets of buffers); forced here on a small batch, compact and inline entry forms."""
    d, model, cfg, batch = load_golden(name, rnn_dir if name.startswith("rnn") else None)
    monkeypatch.setenv("JPP_B200_LANES", "1")
    one, _ = gpu_score(model, cfg, batch)
    monkeypatch.setenv("JPP_B200_LANE_MIN_SENTENCES", "2")
    for lanes, b in (("2", batch), ("2", batch.compact_entries()), ("3", batch), ("4", batch.compact_entries()),
                     ("7", batch)):
        monkeypatch.setenv("JPP_B200_LANES", lanes)
        two, cnt = gpu_score(model, cfg, b)
        assert two.node_beam.tobytes() == one.node_beam.tobytes()
        assert two.node_beam_scores.tobytes() == one.node_beam_scores.tobytes()
        assert np.array_equal(two.node_patterns, one.node_patterns)
        assert two.node_t0_score.tobytes() == one.node_t0_score.tobytes()
        assert two.paths(batch) == one.paths(batch)
        assert cnt["sentences"] == batch.num_sentences and cnt["nodes"] == batch.num_nodes


def abi_mod():
    from jumanpp_b200 import abi
    return abi


def test_errors_are_reported_not_swallowed():
    from jumanpp_b200 import abi
    from jumanpp_b200.scorer import LatticeScorer, ScorerError
    d, model, cfg, batch = load_golden("default")
    with pytest.raises(ScorerError) as e:
        LatticeScorer(model, abi.ScoringConfig(0, 6, 1, 5))
    assert e.value.code == abi.INVALID_PARAMETER
    with pytest.raises(ScorerError) as e:
        LatticeScorer(model, abi.ScoringConfig(5, 6, 1, 0))
    assert e.value.code == abi.INVALID_PARAMETER
    with pytest.raises(ScorerError) as e:
        LatticeScorer(model, abi.ScoringConfig(40, 6, 1, 5))   # beams above 32 are refused, not truncated
    assert e.value.code == abi.NOT_IMPLEMENTED


# The default build picks the compiled-in Jumandic program and, for layer size 128, the tcgen05 level kernel.  The
# kernels every OTHER model gets (table-driven feature program, FFMA level kernel) are held to the same golden
# vectors: the switches are read when a handle is created.
@pytest.mark.parametrize("name", ["default", "beam20", "beam3_ties", "dyn_default", "ref_codegen_fixture"])
def test_table_driven_program_matches_reference_and_oracle(name, monkeypatch):
    monkeypatch.setenv("JPP_B200_NO_STATIC", "1")
    d, model, cfg, batch = load_golden(name)
    res, cnt = gpu_score(model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=False)
    assert rep == {"patterns": 0, "beam_links": 0, "beam_totals": 0, "conn_scores": 0}, rep
    ora, stats = oracle_lib.oracle_score(model, cfg, batch)
    assert res.node_beam.tobytes() == ora.node_beam.tobytes()
    assert np.array_equal(res.node_patterns, ora.node_patterns)
    assert res.node_t0_score.tobytes() == ora.node_t0_score.tobytes()
    assert cnt["gathers"] == stats.gathers


def test_ffma_level_kernel_for_layer_size_128(rnn_dir, monkeypatch):
    monkeypatch.setenv("JPP_B200_RNN_TC", "0")
    monkeypatch.setenv("JPP_B200_NO_STATIC", "1")
    d, model, cfg, batch = load_golden("rnn_h128", rnn_dir)
    res, _ = gpu_score(model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=True, rnn_rtol=1e-4)
    n = rep.pop("final_path_elements")
    assert n > 0 and all(v == 0 for v in rep.values()), rep
