"""Independent check of the RNN arithmetic (VERDICT r1 "missing" 6).

The RNN float parity of this repository is pinned to "reference sources + oracle/eigen_shim"
(Eigen is a configure-time download and absent).  The reference vendors a second, Eigen-free
implementation of the same model, src/rnn/legacy/rnnlmlib_static.cpp; oracle/rnn_legacy_check.cc
scores the same random word sequences through it and through the reference's MikolovRnn (+ shim),
with the semantics and the tolerance of the reference's own comparison
(mikolov_rnn_test.cc:240-277, 300-327: log10 scores, 1e-3 absolute).  The binary is built by
oracle/build_ref.py where /root/reference is mounted and travels in oracle/_ref/."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from jumanpp_b200 import synth  # noqa: E402

BIN = ROOT / "oracle" / "_ref" / "rnn_legacy_check"


@pytest.mark.parametrize("layer,order", [(16, 3), (48, 3), (128, 3), (256, 4), (64, 2)])
def test_mikolov_rnn_with_shim_equals_legacy(tmp_path, layer, order):
    if not BIN.exists():
        pytest.skip("oracle/_ref/rnn_legacy_check not built (needs /root/reference)")
    words = ["</s>", "<unk>"] + [f"w{i}" for i in range(2998)]
    synth.write_rnn_model(tmp_path / "lm", words, layer, order, 1 << 18, 9.0, 11 + layer)
    r = subprocess.run([str(BIN), str(tmp_path / "lm"), "150", "16", "5"], capture_output=True, text=True, check=True)
    o = json.loads(r.stdout.strip().splitlines()[-1])
    assert o["steps"] == 150 * 16 and o["layer_size"] == layer
    # the reference's own tolerance between its two implementations
    assert o["max_abs_log10_apply_vs_legacy"] < 1e-3
    assert o["max_abs_log10_parallel_vs_legacy"] < 1e-3
    # what it really is: float32 rounding of two summation orders (the legacy code sums in
    # double / 8-way unrolled float, the shim in 32 float partials)
    assert o["max_abs_log10_parallel_vs_legacy"] < 2e-5
    assert o["max_abs_context_vs_legacy"] < 2e-6
    # the single-step and the batched entry points of MikolovRnn agree exactly under the shim
    assert o["max_abs_log10_parallel_vs_apply"] == 0.0
