"""pre-decoded dictionary entries resident on the device.  A batch in the
resident form stages nothing per dictionary node beyond its NodeInfo; only the sentences' extra
(unknown-word) nodes bring rows.  Results must equal the reference's post-state exactly as for the
inline form."""
import os

import numpy as np
import pytest

import parity
from conftest import load_golden
from jumanpp_b200 import abi
from jumanpp_b200.scorer import LatticeScorer, ScorerError

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["default", "beam20", "rcheck0", "ref_codegen_fixture", "rnn_h128", "rnn_beam20", "dyn_rnn"])
def test_resident_form_equals_reference(name, rnn_dir):
    d, model, cfg, batch = load_golden(name, rnn_dir)
    rb, table = batch.resident_form()
    assert rb.h2d_bytes() < batch.h2d_bytes()
    sc = LatticeScorer(model, cfg, device=0)
    with pytest.raises(ScorerError) as ei:   # no table yet
        sc.compute_scores(rb)
    assert ei.value.code == abi.INVALID_STATE
    # the table arrives in two pieces (a stager adds the entries it meets for the first time)
    half = len(table.entry_ptr) // 2
    for sl in (slice(0, half), slice(half, None)):
        sc.add_entry_rows(abi.EntryTable(np.ascontiguousarray(table.entry_ptr[sl]), np.ascontiguousarray(table.rows[sl]),
                                         None if table.rnn_id is None else np.ascontiguousarray(table.rnn_id[sl])))
    res = sc.compute_scores(rb)
    assert not res.sentence_status.any()
    rnn = model.rnn is not None
    rep = parity.compare_with_expected(d, batch, res, rnn=rnn, rnn_rtol=1e-4 if rnn else 0.0)
    rep.pop("final_path_elements", None)
    assert all(v == 0 for v in rep.values()), rep
    inline = sc.compute_scores(batch)
    assert inline.node_beam.tobytes() == res.node_beam.tobytes()
    assert np.array_equal(inline.best_path_nodes, res.best_path_nodes)
    # several lanes: each window brings its own slice of the extra rows
    os.environ["JPP_B200_LANE_MIN_SENTENCES"] = "2"
    try:
        lanes = sc.compute_scores(rb)
    finally:
        del os.environ["JPP_B200_LANE_MIN_SENTENCES"]
    assert lanes.node_beam.tobytes() == res.node_beam.tobytes()
    assert lanes.node_beam_scores.tobytes() == res.node_beam_scores.tobytes()
    # a dictionary entry the table does not have fails its sentence, not the batch
    bad = abi.StagedBatch(**{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in rb.__dict__.items()})
    n = int(np.nonzero(bad.node_info["entry_ptr"] >= 0)[0][len(bad.node_info) // 2])
    bad.node_info["entry_ptr"][n] = int(table.entry_ptr.max()) + 2
    got = sc.compute_scores(bad)
    bnd = int(np.searchsorted(bad.boundary_node_off, n, side="right") - 1)
    s = int(np.searchsorted(bad.sent_boundary_off, bnd, side="right") - 1)
    assert got.sentence_status[s] & abi.SENT_BAD_ENTRY
    assert got.sentence_status.astype(bool).sum() == 1
    sc.close()
