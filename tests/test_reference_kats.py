"""Known-answer tests the reference's own unit tests hold for pieces of this path,
restated against the oracle's building blocks."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib


@pytest.fixture(scope="module")
def lib():
    l = oracle_lib.lib()
    l.jpp_oracle_pack_candidate.restype = C.c_uint64
    l.jpp_oracle_pack_candidate.argtypes = [C.c_float, C.c_uint16, C.c_uint16]
    l.jpp_oracle_candidate_score.restype = C.c_float
    l.jpp_oracle_candidate_score.argtypes = [C.c_uint64]
    l.jpp_oracle_unrolled4.restype = C.c_float
    l.jpp_oracle_unrolled4.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    l.jpp_oracle_mix.restype = C.c_uint64
    l.jpp_oracle_mix.argtypes = [C.c_uint64, C.c_uint64]
    return l


def test_beam_candidate_roundtrip(lib):
    # reference src/core/analysis/score_processor_test.cc:48-61
    for score, left, beam in [(-1.0, 5, 15), (1.5, 55, 155), (0.0, 0, 0)]:
        p = lib.jpp_oracle_pack_candidate(score, left, beam)
        assert lib.jpp_oracle_candidate_score(p) == score
        assert (p >> 16) & 0xFFFF == left and p & 0xFFFF == beam


def test_beam_candidates_order_by_score(lib):
    # score_processor_test.cc:63-89
    pk = lib.jpp_oracle_pack_candidate
    bc0, bc, bc2, bc3, bc4 = pk(0, 0, 0), pk(-1.5, 5, 15), pk(-0.1, 5, 15), pk(1.5, 55, 155), pk(2.5, 556, 155)
    assert bc < bc2 < bc0 < bc3 < bc4


def test_beam_candidates_order_by_coordinates(lib):
    # score_processor_test.cc:91-113: equal score -> (left, beam) decides
    pk = lib.jpp_oracle_pack_candidate
    b = [pk(0.836, 0, 0), pk(0.836, 0, 2), pk(0.836, 0, 3), pk(0.836, 1, 0), pk(0.836, 1, 3)]
    assert b == sorted(b) and len(set(b)) == 5


def test_unrolled4_perceptron_sums(lib):
    # src/core/analysis/perceptron_test.cc:11-44 (masking is a no-op for in-range indices)
    def compute(w, idx):
        w = np.asarray(w, np.float32)
        i = np.asarray(idx, np.uint32)
        return lib.jpp_oracle_unrolled4(w.ctypes.data, i.ctypes.data, len(idx))
    w = [0.1] * 16
    for idx, exp in [([1], 0.1), ([1, 2], 0.2), ([1, 2, 3], 0.3), ([1, 2, 5, 9], 0.4), ([1, 7, 5, 9, 3], 0.5)]:
        assert compute(w, idx) == pytest.approx(exp)
    w = [10.0 ** k for k in range(16)]
    for idx, exp in [([1], 10.0), ([1, 2], 110.0), ([1, 2, 3], 1110.0), ([6, 7, 5, 9], 1011100000.0),
                     ([8, 7, 5, 9, 3], 1110101000.0)]:
        assert compute(w, idx) == pytest.approx(exp)


def test_fast_hash_rot_mix_definition(lib):
    # util/fast_hash_rot.h:38-43 with the constants of util/seahash.h:15,17
    seed, mult = 0x16f11fe89b0d677c, 0x6eed0e9da4d94a4f
    for data in (0, 1, 3, 0x7a11ed00000000, 0xFFFFFFFFFFFFFFFF, 0x51239ab41f1f):
        v = ((seed ^ data) * mult) & 0xFFFFFFFFFFFFFFFF
        v = ((v << 32) | (v >> 32)) & 0xFFFFFFFFFFFFFFFF
        assert lib.jpp_oracle_mix(seed, data) == v


def test_synthetic_weight_generators_agree():
    # python (synth.make_weights) == C oracle == the CUDA fill kernel (checked in the gpu tests)
    from jumanpp_b200 import synth
    w = synth.make_weights(12, 7, 0.5)
    out = np.zeros(1 << 12, np.float32)
    oracle_lib.lib().jpp_oracle_fill_weights(out.ctypes.data, 12, 7, 0.5)
    assert out.tobytes() == w.tobytes()
    assert abs(float(w.std()) - 0.5) < 0.02
