"""Pins the CPU oracle (oracle/jpp_oracle.c) against golden vectors produced by
the unmodified reference (tests/golden/make_golden.py): pattern hashes, every
beam element and every connection score must be bit-identical, including the
order of exactly-equal scores and the RNN re-ranked EOS beam."""
import numpy as np
import pytest

import oracle_lib
import parity
from conftest import FULL_BEAM, LINEAR_GBEAM, RNN, load_golden


@pytest.mark.parametrize("name", LINEAR_GBEAM + FULL_BEAM)
def test_linear_path_bit_exact(name):
    d, model, cfg, batch = load_golden(name)
    res, stats = oracle_lib.oracle_score(model, cfg, batch, threads=2)
    rep = parity.compare_with_expected(d, batch, res, rnn=False)
    assert rep == {"patterns": 0, "beam_links": 0, "beam_totals": 0, "conn_scores": 0}, rep
    assert stats.gathers > 0


@pytest.mark.parametrize("name", RNN)
def test_rnn_path_bit_exact(name, rnn_dir):
    d, model, cfg, batch = load_golden(name, rnn_dir)
    res, stats = oracle_lib.oracle_score(model, cfg, batch, threads=2)
    rep = parity.compare_with_expected(d, batch, res, rnn=True)
    n = rep.pop("final_path_elements")
    assert n > 0
    assert rep == {"patterns": 0, "beam_links": 0, "conn_scores": 0, "final_path_totals": 0,
                   "final_path_rnn_scores": 0}, rep
    assert stats.rnn_nodes > 0


def test_threads_do_not_change_results():
    d, model, cfg, batch = load_golden("default")
    r1, _ = oracle_lib.oracle_score(model, cfg, batch, threads=1)
    r4, _ = oracle_lib.oracle_score(model, cfg, batch, threads=4)
    assert r1.node_beam.tobytes() == r4.node_beam.tobytes()  # fake slots hold NaN: compare bytes
    assert np.array_equal(r1.node_patterns, r4.node_patterns)


def test_reference_fixture_is_the_codegen_test_sentence():
    # src/jumandic/shared/jumandic_codegen_test.cc:40-127: codegen.mdic, 64 weights
    d, model, cfg, batch = load_golden("ref_codegen_fixture")
    assert model.log2_weights == 6 and model.weights is not None and model.weights.size == 64
    assert batch.num_sentences == 1
    assert cfg.global_beam_size == 0 and cfg.beam_size == 5
    text = "".join(chr(c) for c in batch.codepoints["codepoint"])
    assert text.startswith("５５１年も")
