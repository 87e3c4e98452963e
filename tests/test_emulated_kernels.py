"""The device code of jumanpp_b200/csrc (the source the GPU runs) compiled as
plain C++ with the lanes of a warp run one after another (tests/emul, test
only) must reproduce the reference's golden vectors bit for bit.  This checks
kernel LOGIC without a GPU; the CUDA build itself is checked by the gpu tests."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
import parity
from conftest import FULL_BEAM, LINEAR_GBEAM, RNN, load_golden
from jumanpp_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
EMUL = ROOT / "tests" / "emul"


@pytest.fixture(scope="module")
def emul():
    src = EMUL / "jpp_b200_emul.cpp"
    lib = EMUL / "libjpp_b200_emul.so"
    deps = [src] + list((ROOT / "jumanpp_b200" / "csrc").glob("*.h")) + [ROOT / "include" / "jpp_b200.h"]
    if not lib.exists() or lib.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-march=x86-64-v3", "-fPIC", "-shared", "-w",
                        f"-I{ROOT / 'include'}", f"-I{ROOT / 'jumanpp_b200' / 'csrc'}", str(src), "-o", str(lib)],
                       check=True)
    return C.CDLL(str(lib))


def run_emul(emul, model, cfg, batch):
    res = abi.ScoreResult.allocate(batch.num_nodes, cfg.beam_size, model.num_scorers, model.num_stored_patterns,
                                   batch.num_sentences, batch.num_boundaries)
    w = oracle_lib.weights_for(model)
    ms, cs, bs, rs = model.as_struct(), cfg.as_struct(), batch.as_struct(), res.as_struct()
    err = C.create_string_buffer(512)
    rc = emul.jpp_b200_emul_score_batch(C.byref(ms), C.byref(cs), C.byref(bs), C.byref(rs), C.c_void_p(w.ctypes.data),
                                        err, 512)
    assert rc == 0, err.value
    return res


@pytest.mark.parametrize("name", LINEAR_GBEAM + FULL_BEAM)
def test_linear_kernels(emul, name):
    d, model, cfg, batch = load_golden(name)
    res = run_emul(emul, model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=False)
    assert rep == {"patterns": 0, "beam_links": 0, "beam_totals": 0, "conn_scores": 0}, rep
    # device-resolved 1-best paths == the walk a host would do over the beams
    bno, sbo = batch.boundary_node_off, batch.sent_boundary_off
    want = [[int(bno[int(sbo[s]) + b]) + p for b, p in path[:-1]] for s, path in enumerate(abi.best_paths(batch, res))]
    assert res.paths(batch) == want


@pytest.mark.parametrize("name", RNN)
def test_rnn_kernels(emul, name, rnn_dir):
    d, model, cfg, batch = load_golden(name, rnn_dir)
    res = run_emul(emul, model, cfg, batch)
    rep = parity.compare_with_expected(d, batch, res, rnn=True)
    rep.pop("final_path_elements")
    assert all(v == 0 for v in rep.values()), rep


def test_compact_entry_rows_give_the_same_result(emul):
    """jpp_b200_batch.node_entry_row / entry_rows (one row per distinct entry) against the inline form;
    also through concat (row tables stacked, indices rebased) and select."""
    d, model, cfg, batch = load_golden("default")
    inline = run_emul(emul, model, cfg, batch)
    compact = batch.compact_entries()
    assert compact.node_entry is None and len(compact.entry_rows) < batch.num_nodes
    assert compact.h2d_bytes() < batch.h2d_bytes()
    got = run_emul(emul, model, cfg, compact)
    assert got.node_beam.tobytes() == inline.node_beam.tobytes()
    assert (got.node_patterns == inline.node_patterns).all()
    ns = batch.num_sentences
    parts = [batch.select(0, ns // 2).compact_entries(), batch.select(ns // 2, ns).compact_entries()]
    again = abi.StagedBatch.concat(parts)
    got2 = run_emul(emul, model, cfg, again)
    assert got2.node_beam.tobytes() == inline.node_beam.tobytes()
    dedup = again.compact_entries()
    assert len(dedup.entry_rows) == len(compact.entry_rows)
    got3 = run_emul(emul, model, cfg, dedup.select(1, ns - 1))
    ref3 = run_emul(emul, model, cfg, batch.select(1, ns - 1))
    assert got3.node_beam.tobytes() == ref3.node_beam.tobytes()
    ora, _ = oracle_lib.oracle_score(model, cfg, compact)
    assert ora.node_beam.tobytes() == inline.node_beam.tobytes()


def test_compact_rows_carry_placeholders_and_rnn_ids(emul, rnn_dir):
    d, model, cfg, batch = load_golden("rnn_default", rnn_dir)
    inline = run_emul(emul, model, cfg, batch)
    compact = batch.compact_entries()
    assert compact.node_placeholders is None and compact.node_rnn_id is None
    assert compact.row_rnn_id is not None and compact.row_placeholders is not None
    got = run_emul(emul, model, cfg, compact)
    assert got.node_beam.tobytes() == inline.node_beam.tobytes()
    assert got.node_beam_scores.tobytes() == inline.node_beam_scores.tobytes()
    kept = batch.compact_entries(fold=False)      # rows = entry buffers only, per-node arrays kept
    assert kept.node_rnn_id is not None and kept.row_rnn_id is None
    got2 = run_emul(emul, model, cfg, kept)
    assert got2.node_beam.tobytes() == inline.node_beam.tobytes()
    ora, _ = oracle_lib.oracle_score(model, cfg, compact)
    ora0, _ = oracle_lib.oracle_score(model, cfg, batch)
    assert ora.node_beam.tobytes() == ora0.node_beam.tobytes()


@pytest.mark.parametrize("name", LINEAR_GBEAM + FULL_BEAM)
def test_compiled_in_t0_program_equals_the_table_driven_one(emul, name, monkeypatch):
    """jpp_b200_static_jumandic_t0.h (what k_t0_static runs) against the program tables and against the
    reference's expected state: stored pattern rows, unigram scores and everything downstream, bit for bit."""
    d, model, cfg, batch = load_golden(name)
    tables = run_emul(emul, model, cfg, batch)
    monkeypatch.setenv("JPP_B200_EMUL_STATIC_T0", "1")
    static = run_emul(emul, model, cfg, batch)
    assert np.array_equal(tables.node_patterns, static.node_patterns)
    assert np.array_equal(tables.node_t0_score.view(np.uint32), static.node_t0_score.view(np.uint32))
    rep = parity.compare_with_expected(d, batch, static, rnn=False)
    assert rep == {"patterns": 0, "beam_links": 0, "beam_totals": 0, "conn_scores": 0}, rep


def test_a_different_feature_program_is_not_taken_for_the_compiled_in_one(emul, monkeypatch):
    """matches_static_t0 must refuse a program that differs from the generated header in one table entry
    (such a model runs the table-driven kernels)."""
    import dataclasses
    d, model, cfg, batch = load_golden("default")
    ngrams = model.ngrams.copy()
    uni = [i for i in range(len(ngrams)) if ngrams[i][0] == 1]
    ngrams[uni[3]][2], ngrams[uni[4]][2] = ngrams[uni[4]][2], ngrams[uni[3]][2]     # two unigram features swap their patterns
    other = dataclasses.replace(model, ngrams=ngrams)
    run_emul(emul, other, cfg, batch)                                                 # table-driven: fine
    monkeypatch.setenv("JPP_B200_EMUL_STATIC_T0", "1")
    res = abi.ScoreResult.allocate(batch.num_nodes, cfg.beam_size, other.num_scorers, other.num_stored_patterns,
                                   batch.num_sentences, batch.num_boundaries)
    w = oracle_lib.weights_for(other)
    ms, cs, bs, rs = other.as_struct(), cfg.as_struct(), batch.as_struct(), res.as_struct()
    err = C.create_string_buffer(512)
    rc = emul.jpp_b200_emul_score_batch(C.byref(ms), C.byref(cs), C.byref(bs), C.byref(rs), C.c_void_p(w.ctypes.data), err, 512)
    assert rc != 0 and b"not the compiled-in one" in err.value
