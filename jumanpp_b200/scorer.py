"""Python mirror of the reference's scoring interface over the C ABI.

``LatticeScorer`` plays the role ``core::analysis::ScoreProcessor`` +
``AnalyzerImpl::computeScores`` play in the reference (score_processor.h:119-193,
analyzer_impl.cc:299-309), for a *batch* of built lattices: construct it from a
model + scoring configuration, call :meth:`compute_scores` with a
``StagedBatch``, get the post-state (beams, connection scores, patterns).

The CUDA library is mandatory: importing this module works anywhere, but
creating a scorer raises ``ScorerError`` if ``libjpp_b200.so`` is missing or no
CUDA device is usable.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import abi

PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("JPP_B200_LIB", PKG / "libjpp_b200.so"))  # override: kernel tuning variants

_STATUS_NAMES = {abi.OK: "Ok", abi.INVALID_PARAMETER: "InvalidParameter", abi.INVALID_STATE: "InvalidState",
                 abi.NOT_IMPLEMENTED: "NotImplemented", abi.CUDA_ERROR: "InvalidState(CUDA)"}


class ScorerError(RuntimeError):
    """Non-OK jumanpp::Status equivalent (reference src/util/status.hpp:17-23)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ScorerError(abi.INVALID_STATE, f"{LIB_PATH} is missing: build it with `python -m jumanpp_b200.build` "
                                             "(the scorer has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    H = C.c_void_p
    lib.jpp_b200_create.restype = C.c_int
    lib.jpp_b200_create.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Config), C.c_int, C.POINTER(H)]
    lib.jpp_b200_set_config.restype = C.c_int
    lib.jpp_b200_set_config.argtypes = [H, C.POINTER(abi.Config)]
    lib.jpp_b200_score_batch.restype = C.c_int
    lib.jpp_b200_score_batch.argtypes = [H, C.POINTER(abi.Batch), C.POINTER(abi.Result)]
    lib.jpp_b200_upload_batch.restype = C.c_int
    lib.jpp_b200_upload_batch.argtypes = [H, C.POINTER(abi.Batch)]
    lib.jpp_b200_score_resident.restype = C.c_int
    lib.jpp_b200_score_resident.argtypes = [H]
    lib.jpp_b200_download_result.restype = C.c_int
    lib.jpp_b200_download_result.argtypes = [H, C.POINTER(abi.Result)]
    lib.jpp_b200_add_entry_rows.restype = C.c_int
    lib.jpp_b200_add_entry_rows.argtypes = [H, C.POINTER(abi.EntryTableStruct)]
    lib.jpp_b200_gather_probe.restype = C.c_int
    lib.jpp_b200_gather_probe.argtypes = [H, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    lib.jpp_b200_get_counters.restype = C.c_int
    lib.jpp_b200_get_counters.argtypes = [H, C.POINTER(abi.Counters)]
    lib.jpp_b200_last_error.restype = C.c_char_p
    lib.jpp_b200_last_error.argtypes = [H]
    lib.jpp_b200_destroy.restype = None
    lib.jpp_b200_destroy.argtypes = [H]
    lib.jpp_b200_abi_version.restype = C.c_uint32
    lib.jpp_b200_build_info.restype = C.c_char_p
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["jpp_b200_create", "jpp_b200_set_config", "jpp_b200_score_batch", "jpp_b200_upload_batch",
                    "jpp_b200_score_resident", "jpp_b200_download_result", "jpp_b200_get_counters", "jpp_b200_add_entry_rows",
                    "jpp_b200_alloc_pinned", "jpp_b200_free_pinned", "jpp_b200_reserve_entry_rows", "jpp_b200_gather_probe",
                    "jpp_b200_last_error", "jpp_b200_destroy", "jpp_b200_abi_version", "jpp_b200_build_info"]


class LatticeScorer:
    def __init__(self, model: abi.ModelSpec, config: abi.ScoringConfig | None = None, device: int = 0):
        self._lib = load_library()
        self.model = model
        self.config = config or abi.ScoringConfig()
        self._h = C.c_void_p()
        ms, cs = model.as_struct(), self.config.as_struct()
        rc = self._lib.jpp_b200_create(C.byref(ms), C.byref(cs), device, C.byref(self._h))
        if rc != abi.OK:
            msg = self._lib.jpp_b200_last_error(None)
            self._h = C.c_void_p()
            raise ScorerError(rc, (msg or b"").decode())
        self._batch_keep = None

    # -- reference: AnalyzerImpl::setGlobalBeam / JumanppEnv::setBeamSize
    def set_config(self, config: abi.ScoringConfig) -> None:
        cs = config.as_struct()
        self._check(self._lib.jpp_b200_set_config(self._h, C.byref(cs)))
        self.config = config

    def _check(self, rc: int) -> None:
        if rc != abi.OK:
            raise ScorerError(rc, (self._lib.jpp_b200_last_error(self._h) or b"").decode())

    def allocate_result(self, batch: abi.StagedBatch) -> abi.ScoreResult:
        return abi.ScoreResult.allocate(batch.num_nodes, self.config.beam_size, self.model.num_scorers,
                                        self.model.num_stored_patterns, batch.num_sentences, batch.num_boundaries)

    # -- reference: AnalyzerImpl::computeScores for every sentence of the batch
    def compute_scores(self, batch: abi.StagedBatch, result: abi.ScoreResult | None = None,
                       paths_only: bool = False) -> abi.ScoreResult:
        """jpp_b200_score_batch: host buffers in, host buffers out.  `paths_only` asks for the
        device-resolved 1-best paths alone (what the CLI's default output needs)."""
        res = result or self.allocate_result(batch)
        bs, rs = batch.as_struct(), res.as_struct()
        if paths_only:
            rs.node_beam = None
            rs.node_beam_scores = None
            rs.node_patterns = None
            rs.node_t0_score = None
        self._check(self._lib.jpp_b200_score_batch(self._h, C.byref(bs), C.byref(rs)))
        return res

    # -- pre-decoded dictionary entries resident on the device
    def add_entry_rows(self, table: abi.EntryTable) -> None:
        ts = table.as_struct()
        self._check(self._lib.jpp_b200_add_entry_rows(self._h, C.byref(ts)))

    def upload(self, batch: abi.StagedBatch) -> None:
        bs = batch.as_struct()
        self._batch_keep = batch
        self._check(self._lib.jpp_b200_upload_batch(self._h, C.byref(bs)))

    def score_resident(self) -> None:
        self._check(self._lib.jpp_b200_score_resident(self._h))

    def download(self, result: abi.ScoreResult, beams_only: bool = False, paths_only: bool = False) -> abi.ScoreResult:
        """Copy back the requested part of the post-state: everything, the beams (+ paths), or only
        the device-resolved 1-best paths (what the CLI's default output needs)."""
        rs = result.as_struct()
        if beams_only or paths_only:
            rs.node_beam_scores = None
            rs.node_patterns = None
            rs.node_t0_score = None
        if paths_only:
            rs.node_beam = None
        self._check(self._lib.jpp_b200_download_result(self._h, C.byref(rs)))
        return result

    def counters(self) -> dict:
        c = abi.Counters()
        self._check(self._lib.jpp_b200_get_counters(self._h, C.byref(c)))
        names = ["t0", "beam_walk", "rnn_tree", "rnn_contexts", "rnn_scores", "rerank", "upload", "score_total"]
        return {"sentences": c.sentences, "boundaries": c.boundaries, "nodes": c.nodes, "gathers": c.gathers,
                "algorithmic_bytes": c.algorithmic_bytes, "rnn_nodes": c.rnn_nodes,
                "walk_gathers": c.walk_gathers, "walk_bytes": c.walk_bytes, "launches": c.launches,
                "walk_gathers_issued": c.walk_gathers_issued, "rnn_scores": c.rnn_scores,
                "failed_sentences": c.failed_sentences,
                "kernel_ms": {n: float(c.kernel_ms[i]) for i, n in enumerate(names)}}

    def gather_probe(self, gathers: int = 1 << 31, ilp: int = 8, blocks_per_sm: int = 8) -> float:
        """Diagnostic: random 4-byte reads per second this device serves from the handle's perceptron table."""
        out = C.c_double(0.0)
        self._check(self._lib.jpp_b200_gather_probe(self._h, gathers, ilp, blocks_per_sm, C.byref(out)))
        return float(out.value)

    def close(self) -> None:
        if self._h:
            self._lib.jpp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_copy(a: np.ndarray) -> np.ndarray:
    """Copy into page-locked host memory (torch's pinned allocator) so the H2D
    copies of a batch run at full PCIe speed; falls back to a plain copy when
    torch/CUDA is unavailable."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
            out = t.numpy().view(a.dtype).reshape(a.shape)
            out[...] = a
            return out
    except Exception:
        pass
    return np.ascontiguousarray(a)
