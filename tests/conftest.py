import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"

# parameters of the synthetic RNN models the rnn_* golden files were staged with
# (tests/golden/make_golden.py); the .nnet blocks depend only on these numbers
RNN_MODELS = {
    "rnn_default": dict(layer_size=32, maxent_order=3, maxent_size=1 << 16, nce_lnz=9.0, seed=2, vocab=5000),
    "rnn_beam20": dict(layer_size=64, maxent_order=4, maxent_size=1 << 15, nce_lnz=7.5, seed=5, vocab=3000),
    "rnn_h128": dict(layer_size=128, maxent_order=3, maxent_size=1 << 16, nce_lnz=9.0, seed=11, vocab=4000),
    "rnn_h48": dict(layer_size=48, maxent_order=2, maxent_size=1 << 14, nce_lnz=8.0, seed=13, vocab=2000),
    "rnn_h256": dict(layer_size=256, maxent_order=3, maxent_size=1 << 15, nce_lnz=9.0, seed=17, vocab=2000),
    "dyn_rnn": dict(layer_size=32, maxent_order=3, maxent_size=1 << 16, nce_lnz=9.0, seed=2, vocab=5000),
}

LINEAR_GBEAM = ["default", "beam20", "beam20_all", "gbeam20_partition", "beam3_ties", "beam1", "rcheck0",
                "ref_codegen_fixture_gbeam", "dyn_default", "dyn_beam3_ties"]
FULL_BEAM = ["ref_codegen_fixture", "dyn_full"]
RNN = ["rnn_default", "rnn_beam20", "rnn_h128", "rnn_h48", "rnn_h256", "dyn_rnn"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rnn_dir(tmp_path_factory):
    """Regenerate the synthetic .nnet files the rnn_* fixtures refer to."""
    from jumanpp_b200 import synth
    d = tmp_path_factory.mktemp("rnn")
    for name, p in RNN_MODELS.items():
        words = ["</s>", "<unk>"] + [f"w{i}" for i in range(p["vocab"] - 2)]
        synth.write_rnn_model(d / name, words, p["layer_size"], p["maxent_order"], p["maxent_size"], p["nce_lnz"],
                              p["seed"])
    return d


def load_golden(name, rnn_dir=None):
    from jumanpp_b200 import abi, jppb
    d = jppb.load(GOLDEN / f"{name}.jppb.gz")
    prefix = str(rnn_dir / name) if name in RNN_MODELS else None
    model = abi.ModelSpec.from_jppb(d, prefix)
    cfg = abi.ScoringConfig.from_jppb(d)
    batch = abi.StagedBatch.from_jppb(d)
    return d, model, cfg, batch
