"""The CUDA path under scoring configurations the golden files do not cover (the group-cooperative
kernels with 32-lane groups, beams above 8, right check > 1, right check off, full beam with other
sizes, RNN with other beams): GPU == oracle, bit for bit on the linear path."""
import numpy as np
import pytest

import oracle_lib
from conftest import load_golden
from jumanpp_b200 import abi
from test_config_sweep import LINEAR

pytestmark = pytest.mark.gpu


def gpu_score(model, cfg, batch):
    from jumanpp_b200.scorer import LatticeScorer
    sc = LatticeScorer(model, cfg)
    try:
        return sc.compute_scores(batch)
    finally:
        sc.close()


@pytest.mark.parametrize("cfgv", LINEAR)
def test_linear_configs(cfgv):
    d, model, cfg, batch = load_golden("default")
    cfg = abi.ScoringConfig(*cfgv)
    ora, _ = oracle_lib.oracle_score(model, cfg, batch)
    got = gpu_score(model, cfg, batch)
    live = ora.node_beam["left"] != 0xFFFF
    assert np.array_equal(got.node_beam["left"], ora.node_beam["left"])
    assert np.array_equal(got.node_beam["prev_beam"], ora.node_beam["prev_beam"])
    assert got.node_beam["total_score"][live].tobytes() == ora.node_beam["total_score"][live].tobytes()
    assert got.node_beam_scores[:, :, 0][live].tobytes() == ora.node_beam_scores[:, :, 0][live].tobytes()
    assert np.array_equal(got.node_patterns, ora.node_patterns)


@pytest.mark.parametrize("cfgv", [(5, 6, 2, 3), (3, 4, 1, 3), (8, 8, 1, 5), (12, 16, 2, 4), (5, 6, 0, 0)])
def test_rnn_configs(cfgv, rnn_dir):
    d, model, cfg, batch = load_golden("rnn_h128", rnn_dir)
    cfg = abi.ScoringConfig(*cfgv)
    ora, _ = oracle_lib.oracle_score(model, cfg, batch)
    got = gpu_score(model, cfg, batch)
    live = ora.node_beam["left"] != 0xFFFF
    assert np.array_equal(got.node_beam["left"], ora.node_beam["left"])
    assert np.array_equal(got.node_beam["prev_beam"], ora.node_beam["prev_beam"])
    assert got.node_beam_scores[:, :, 0][live].tobytes() == ora.node_beam_scores[:, :, 0][live].tobytes()
    # totals carry RNN scores on the final paths: CUDA expf vs libm expf
    assert np.allclose(got.node_beam["total_score"][live], ora.node_beam["total_score"][live], rtol=1e-4, atol=1e-4)
