"""A malformed sentence fails alone (the CLI prints `# ERROR` for a sentence and
goes on, jumanpp.cc:166-171): the upload's validation kernel marks it in
jpp_b200_result.sentence_status, every scoring kernel skips it, and the other sentences of the batch
come out exactly as in a clean run.  Without a status array the call reports INVALID_PARAMETER."""
import dataclasses

import numpy as np
import pytest

from conftest import load_golden
from jumanpp_b200 import abi
from jumanpp_b200.scorer import LatticeScorer, ScorerError

pytestmark = pytest.mark.gpu


def _corrupt(batch: abi.StagedBatch, model):
    """Returns (bad batch, {sentence: expected status bit})."""
    b = dataclasses.replace(batch, **{f.name: (None if getattr(batch, f.name) is None else getattr(batch, f.name).copy())
                                      for f in dataclasses.fields(batch)})
    sbo, bno, beo = b.sent_boundary_off, b.boundary_node_off, b.boundary_end_off
    want = {}

    def boundary(s, k):
        return int(sbo[s]) + k

    # sentence 1: an ends entry that names a node its boundary does not have
    e = int(beo[boundary(1, 3)])
    b.ends["position"][e] = 60000
    want[1] = abi.SENT_BAD_ENDS
    # sentence 3: an ends entry that points forward
    e = int(beo[boundary(3, 4)])
    b.ends["boundary"][e] = 50
    want[3] = abi.SENT_BAD_ENDS
    # sentence 4: node span outside the sentence
    n = int(bno[boundary(4, 2)])
    b.node_info["end"][n] = 9999
    want[4] = abi.SENT_BAD_SPAN
    # sentence 6: node offsets going backwards inside the sentence
    k = boundary(6, 4)
    bno[k], bno[k + 1] = bno[k + 1], bno[k]
    want[6] = abi.SENT_BAD_OFFSETS
    # sentence 8: entry row out of range (compact form)
    if b.node_entry_row is not None:
        b.node_entry_row[int(bno[boundary(8, 3)])] = 0x7FFFFFFF
        want[8] = abi.SENT_BAD_ENTRY
    return b, want


@pytest.mark.parametrize("name,compact", [("default", False), ("default", True), ("rnn_h128", True), ("beam20", False)])
def test_bad_sentences_fail_alone(name, compact, rnn_dir):
    d, model, cfg, batch = load_golden(name, rnn_dir)
    if compact:
        batch = batch.compact_entries()
    sc = LatticeScorer(model, cfg, device=0)
    clean = sc.compute_scores(batch)
    assert not clean.sentence_status.any()
    bad, want = _corrupt(batch, model)
    if model.rnn is not None:  # an RNN id past the vocabulary
        if bad.row_rnn_id is not None:
            n = int(bad.boundary_node_off[int(bad.sent_boundary_off[10]) + 2])
            bad.row_rnn_id = bad.row_rnn_id.copy()
            row = int(bad.node_entry_row[n])
            bad.row_rnn_id[row] = model.rnn.vocab_size + 5
            # the row may be shared: every sentence that uses it fails, and must say so
            users = np.unique(np.searchsorted(bad.boundary_node_off, np.nonzero(bad.node_entry_row == row)[0], side="right") - 1)
            for s in np.unique(np.searchsorted(bad.sent_boundary_off, users, side="right") - 1):
                want[int(s)] = want.get(int(s), 0) | abi.SENT_BAD_RNN_ID
    got = sc.compute_scores(bad)
    cnt = sc.counters()
    status = got.sentence_status
    for s in range(batch.num_sentences):
        if s in want:
            assert status[s] & want[s], (s, status[s], want[s])
            assert got.best_path_len[s] == 0
        else:
            assert status[s] == 0, (s, status[s])
    assert cnt["failed_sentences"] == len(want)
    # everything of the sentences that were scored equals the clean run, bit for bit
    sbo, bno = batch.sent_boundary_off, batch.boundary_node_off
    for s in range(batch.num_sentences):
        if s in want:
            continue
        n0, n1 = int(bno[sbo[s]]), int(bno[sbo[s + 1]])
        assert got.node_beam[n0:n1].tobytes() == clean.node_beam[n0:n1].tobytes(), s
        assert np.array_equal(got.node_patterns[n0:n1], clean.node_patterns[n0:n1])
        assert got.node_beam_scores[n0:n1].tobytes() == clean.node_beam_scores[n0:n1].tobytes()
        assert got.best_path_len[s] == clean.best_path_len[s]
        b0 = int(sbo[s])
        k = int(got.best_path_len[s])
        assert np.array_equal(got.best_path_nodes[b0:b0 + k], clean.best_path_nodes[b0:b0 + k])
    # a caller that did not ask for per-sentence status is told that the batch had failures
    res = sc.allocate_result(bad)
    res.sentence_status = None
    with pytest.raises(ScorerError) as ei:
        sc.compute_scores(bad, res)
    assert ei.value.code == abi.INVALID_PARAMETER and "failed validation" in ei.value.message
    # the handle is still usable
    again = sc.compute_scores(batch)
    assert again.node_beam.tobytes() == clean.node_beam.tobytes()
    sc.close()


def test_empty_batch_is_a_no_op(rnn_dir):
    d, model, cfg, batch = load_golden("default")
    sc = LatticeScorer(model, cfg, device=0)
    # an empty batch is a no-op
    empty = batch.select(0, 0)
    res = sc.compute_scores(empty)
    assert len(res.best_path_len) == 0
    sc.close()
