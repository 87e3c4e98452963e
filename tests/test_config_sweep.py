"""Scoring configurations beyond the golden files: the device code (CPU build, tests/emul) against the
oracle on the golden lattices under other beam / global-beam / right-check settings.  The oracle is
pinned to the reference on the golden configurations; here it carries that pin to the code paths those
do not reach (right check > 1, global beam > 8 with small beams, partition + sort, full beam with other
beam sizes, RNN with other beams)."""
import dataclasses

import numpy as np
import pytest

import oracle_lib
from conftest import load_golden
from jumanpp_b200 import abi
from test_emulated_kernels import emul, run_emul  # noqa: F401  (fixture + helper)

LINEAR = [  # beam, global beam, right check, right beam
    (5, 6, 2, 4), (5, 6, 3, 5), (7, 9, 1, 5), (8, 8, 2, 3), (2, 3, 1, 2), (6, 5, 1, 6), (3, 12, 3, 2),
    (16, 24, 2, 6), (32, 32, 1, 8), (5, 6, 0, 0), (4, 10, 0, 0), (5, 0, 0, 0), (2, 0, 0, 0), (9, -1, 0, 0),
]


@pytest.mark.parametrize("cfgv", LINEAR)
def test_linear_configs(emul, cfgv):
    d, model, cfg, batch = load_golden("default")
    cfg = abi.ScoringConfig(*cfgv)
    ora, _ = oracle_lib.oracle_score(model, cfg, batch)
    got = run_emul(emul, model, cfg, batch)
    live = ora.node_beam["left"] != 0xFFFF
    assert np.array_equal(got.node_beam["left"], ora.node_beam["left"])
    assert np.array_equal(got.node_beam["prev_beam"], ora.node_beam["prev_beam"])
    assert got.node_beam["total_score"][live].tobytes() == ora.node_beam["total_score"][live].tobytes()
    assert got.node_beam_scores[:, :, 0][live].tobytes() == ora.node_beam_scores[:, :, 0][live].tobytes()
    assert np.array_equal(got.node_patterns, ora.node_patterns)


@pytest.mark.parametrize("cfgv", [(5, 6, 2, 3), (3, 4, 1, 3), (8, 8, 1, 5), (12, 16, 2, 4), (5, 6, 0, 0)])
def test_rnn_configs(emul, cfgv, rnn_dir):
    d, model, cfg, batch = load_golden("rnn_h128", rnn_dir)
    cfg = abi.ScoringConfig(*cfgv)
    ora, _ = oracle_lib.oracle_score(model, cfg, batch)
    got = run_emul(emul, model, cfg, batch)
    live = ora.node_beam["left"] != 0xFFFF
    assert np.array_equal(got.node_beam["left"], ora.node_beam["left"])
    assert np.array_equal(got.node_beam["prev_beam"], ora.node_beam["prev_beam"])
    # same float order on both sides (fmaf + libm expf in the CPU build): bit-equal, RNN included
    assert got.node_beam["total_score"][live].tobytes() == ora.node_beam["total_score"][live].tobytes()
    assert got.node_beam_scores[live].tobytes() == ora.node_beam_scores[live].tobytes()
