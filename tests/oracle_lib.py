"""TEST-ONLY loader for the CPU oracle (oracle/libjpp_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module; the product package (jumanpp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from jumanpp_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "libjpp_oracle.so"


class OracleStats(C.Structure):
    _fields_ = [("sentences", C.c_uint64), ("boundaries", C.c_uint64), ("nodes", C.c_uint64),
                ("gathers", C.c_uint64), ("rnn_nodes", C.c_uint64), ("rnn_scores", C.c_uint64)]


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "jpp_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-march=x86-64-v3", "-fPIC", "-shared",
                        f"-I{ROOT / 'include'}", f"-I{ORACLE_DIR}", str(src), "-o", str(LIB), "-lpthread", "-lm"],
                       check=True)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(str(LIB))
        _lib.jpp_oracle_score_batch.restype = C.c_int
        _lib.jpp_oracle_score_batch.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Config), C.POINTER(abi.Batch),
                                                C.POINTER(abi.Result), C.c_void_p, C.c_int, C.POINTER(OracleStats)]
        _lib.jpp_oracle_fill_weights.restype = None
        _lib.jpp_oracle_fill_weights.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_float]
    return _lib


_weight_cache: dict[tuple[int, int, float], np.ndarray] = {}


def weights_for(model: abi.ModelSpec) -> np.ndarray:
    if model.weights is not None:
        return model.weights
    key = (model.log2_weights, model.weight_seed, model.weight_scale)
    if key not in _weight_cache:
        from jumanpp_b200 import synth
        _weight_cache[key] = synth.make_weights(model.log2_weights, model.weight_seed, model.weight_scale)
    return _weight_cache[key]


def oracle_score(model: abi.ModelSpec, cfg: abi.ScoringConfig, batch: abi.StagedBatch, threads: int = 1):
    res = abi.ScoreResult.allocate(batch.num_nodes, cfg.beam_size, model.num_scorers, model.num_stored_patterns)
    w = weights_for(model)
    ms, cs, bs, rs = model.as_struct(), cfg.as_struct(), batch.as_struct(), res.as_struct()
    st = OracleStats()
    rc = lib().jpp_oracle_score_batch(C.byref(ms), C.byref(cs), C.byref(bs), C.byref(rs), w.ctypes.data, threads,
                                      C.byref(st))
    if rc != 0:
        raise RuntimeError(f"jpp_oracle_score_batch failed: {rc}")
    return res, st
