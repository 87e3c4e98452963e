"""Shared comparison helpers for the parity tests (oracle vs reference golden
vectors, CUDA vs oracle)."""
from __future__ import annotations

import numpy as np

from jumanpp_b200 import abi


def final_path_elements(batch: abi.StagedBatch, node_beam: np.ndarray) -> set[tuple[int, int]]:
    """(node, slot) of every beam element reachable from an EOS beam element."""
    out: set[tuple[int, int]] = set()
    sbo, bno, beo = batch.sent_boundary_off, batch.boundary_node_off, batch.boundary_end_off
    for s in range(batch.num_sentences):
        b0 = int(sbo[s]); nb = int(sbo[s + 1]) - b0
        eos = int(bno[b0 + nb - 1])
        for k in range(node_beam.shape[1]):
            if node_beam[eos, k]["left"] == 0xFFFF:
                break
            bnd, node, slot = nb - 1, eos, k
            while bnd >= 2 and (node, slot) not in out:
                out.add((node, slot))
                el = node_beam[node, slot]
                ptr = batch.ends[int(beo[b0 + bnd]) + int(el["left"])]
                bnd = int(ptr["boundary"]); node = int(bno[b0 + bnd]) + int(ptr["position"]); slot = int(el["prev_beam"])
    return out


def compare_with_expected(d: dict, batch: abi.StagedBatch, res: abi.ScoreResult, rnn: bool,
                          rnn_rtol: float = 0.0) -> dict[str, int]:
    """Compare a result with the reference dump `d` (exp_* arrays). Returns mismatch counts."""
    rep: dict[str, int] = {}
    # the reference's dynamic feature path keeps every pattern per node (uni-only ones included);
    # the first num_stored_patterns columns are the ones the n-gram features of t1 / t2 read
    pat_bad = d["exp_patterns"][:, :res.node_patterns.shape[1]] != res.node_patterns
    if len(d["cfg"]) > 8 and int(d["cfg"][8]):
        # ... and it fills them while the lattice is built, where nodes that no path reaches get zero or
        # stale rows (their beams are fake: nothing ever reads them); the static path hashes every node
        unreachable = d["exp_beam_lp"][:, 0, 0] == 0xFFFF
        pat_bad = pat_bad & ~unreachable[:, None]
    rep["patterns"] = int(pat_bad.sum())
    lp, tot, sc = d["exp_beam_lp"], d["exp_beam_total"], d["exp_beam_scores"]
    fake_e = lp[:, :, 0] == 0xFFFF
    got = res.node_beam
    eos_nodes = batch.boundary_node_off[batch.sent_boundary_off[1:] - 1]
    is_eos = np.zeros(batch.num_nodes, bool); is_eos[eos_nodes] = True
    lp_bad = (lp[:, :, 0] != got["left"]) | (lp[:, :, 1] != got["prev_beam"])
    tb, ob = tot.view(np.uint32), got["total_score"].view(np.uint32)
    s0_bad = (sc[:, :, 0].view(np.uint32) != res.node_beam_scores[:, :, 0].view(np.uint32)) & ~fake_e
    if not rnn:
        rep["beam_links"] = int(lp_bad.sum())
        rep["beam_totals"] = int(((tb != ob) & ~fake_e).sum())
        rep["conn_scores"] = int(s0_bad.sum())
        return rep
    # RNN: links of non-EOS nodes are untouched by the re-rank; EOS row is the final beam
    rep["beam_links"] = int(lp_bad.sum())
    rep["conn_scores"] = int(s0_bad[~is_eos].sum())
    elems = final_path_elements(batch, got)
    bad_t = bad_r = 0
    for node, slot in elems:
        e, g = tot[node, slot], got["total_score"][node, slot]
        r_e, r_g = sc[node, slot, 1], res.node_beam_scores[node, slot, 1]
        if rnn_rtol == 0.0:
            bad_t += int(np.float32(e).view(np.uint32) != np.float32(g).view(np.uint32))
            bad_r += int(np.float32(r_e).view(np.uint32) != np.float32(r_g).view(np.uint32))
        else:
            bad_t += int(abs(e - g) > rnn_rtol * max(1.0, abs(e)))
            bad_r += int(abs(r_e - r_g) > rnn_rtol * max(1.0, abs(r_e)))
    rep["final_path_totals"] = bad_t
    rep["final_path_rnn_scores"] = bad_r
    rep["final_path_elements"] = len(elems)
    return rep
