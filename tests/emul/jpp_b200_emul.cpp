// TEST ONLY.  Serial CPU build of the *device* code in
// jumanpp_b200/csrc/jpp_b200_kernels.h and jpp_b200_rnn.h: the same source the
// GPU runs, with every kernel "launch" executed as a plain loop over its index.  It exists so that kernel
// logic can be compared with the oracle on a machine without a GPU.  It is
// built into tests/emul/libjpp_b200_emul.so, never into libjpp_b200.so, and
// nothing in the product loads it.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jpp_b200.h"
#include "jpp_b200_host.h"
#include "jpp_b200_kernels.h"
#include "jpp_b200_rnn.h"
#include "jpp_b200_t0_static.h"

using namespace jppb200;

extern "C" int jpp_b200_emul_score_batch(const jpp_b200_model* model, const jpp_b200_config* config,
                                         const jpp_b200_batch* b, jpp_b200_result* r, const float* weights,
                                         char* errbuf, int errlen) {
  HostModel hm;
  std::string err;
  jpp_b200_status st = compile_host_model(model, &hm, &err);
  if (st != JPP_B200_OK) {
    if (errbuf) snprintf(errbuf, (size_t)errlen, "%s", err.c_str());
    return (int)st;
  }
  const FeatTables& ft = hm.ft;
  ScoreCfg cfg{config->beam_size, config->global_beam_size, config->right_gbeam_check, config->right_gbeam_size,
               model->num_scorers, ft.dynamicOrder};
  const u32 NS = b->num_sentences, NB = b->sent_boundary_off[NS], NN = b->boundary_node_off[NB],
            NE = b->boundary_end_off[NB];
  const u32 B = (u32)cfg.beam, G = (u32)std::max(1, cfg.gbeam), S = cfg.S, SP = ft.nStored;
  const u32 RC = cfg.rcheck > 0 ? (u32)cfg.rcheck : 1u;
  std::vector<u32> nodeBnd(NN), bndSent(NB), endNode(NE), cutIdx(NN), gbNode((size_t)NB * G);
  std::vector<u32> selPos(NN);
  std::vector<float> bi0(NN);
  std::vector<u16> gbSlot((size_t)NB * G), gbLeftBnd((size_t)NB * G);
  std::vector<u8> gbCount(NB);
  std::vector<float> pre((size_t)NN * RC), cut(NN), eosLocal((size_t)NS * G), eosRnn((size_t)NS * G);
  std::vector<u64> eosGbeam((size_t)NS * G);
  unsigned long long counters[8] = {0};
  for (u32 s = 0; s < NS; ++s)
    for (u32 bb = b->sent_boundary_off[s]; bb < b->sent_boundary_off[s + 1]; ++bb) bndSent[bb] = s;
  for (u32 bb = 0; bb < NB; ++bb) {
    for (u32 n = b->boundary_node_off[bb]; n < b->boundary_node_off[bb + 1]; ++n) nodeBnd[n] = bb;
    const u32 b0 = b->sent_boundary_off[bndSent[bb]];
    for (u32 e = b->boundary_end_off[bb]; e < b->boundary_end_off[bb + 1]; ++e)
      endNode[e] = b->boundary_node_off[b0 + b->ends[e].boundary] + b->ends[e].position;
  }
  std::vector<float> t0tmp(NN);
  BatchView bv{};
  bv.numSentences = NS; bv.totalBoundaries = NB; bv.totalNodes = NN; bv.totalEnds = NE;
  bv.sentBndOff = b->sent_boundary_off; bv.sentCpOff = b->sent_codepoint_off; bv.bndNodeOff = b->boundary_node_off;
  bv.bndEndOff = b->boundary_end_off; bv.nodeInfo = b->node_info; bv.nodeEntry = b->node_entry; bv.nodeEntryRow = b->node_entry_row; bv.entryRows = b->entry_rows; bv.numEntryRows = b->num_entry_rows;
  bv.nodePlaceholders = ft.nPlaceholders ? b->node_placeholders : nullptr; bv.nodeRnnId = b->node_rnn_id;
  bv.rowPlaceholders = ft.nPlaceholders ? b->row_placeholders : nullptr; bv.rowRnnId = b->row_rnn_id;
  bv.ends = b->ends; bv.cps = b->codepoints;
  bv.nodeBnd = nodeBnd.data(); bv.bndSent = bndSent.data(); bv.endNode = endNode.data(); bv.sentOrder = nullptr;
  bv.beam = r->node_beam; bv.beamScores = r->node_beam_scores; bv.patterns = r->node_patterns; bv.patNodeStride = SP; bv.patPStride = 1;
  bv.t0 = r->node_t0_score ? r->node_t0_score : t0tmp.data();
  bv.pre = pre.data(); bv.cut = cut.data(); bv.cutIdx = cutIdx.data(); bv.selPos = selPos.data(); bv.bi0 = bi0.data();
  bv.gbNode = gbNode.data(); bv.gbSlot = gbSlot.data(); bv.gbLeftBnd = gbLeftBnd.data(); bv.gbCount = gbCount.data();
  bv.eosGbeam = eosGbeam.data(); bv.eosLocal = eosLocal.data(); bv.eosRnn = eosRnn.data(); bv.counters = counters;
  const WLoadDirect wl{weights};
  // JPP_B200_EMUL_STATIC_T0=1: the compiled-in per-node program (what k_t0_static runs) instead of the tables
  const bool staticT0 = getenv("JPP_B200_EMUL_STATIC_T0") && atoi(getenv("JPP_B200_EMUL_STATIC_T0")) != 0;
  if (staticT0 && !matches_static_t0(ft)) {
    if (errbuf) snprintf(errbuf, (size_t)errlen, "the model's feature program is not the compiled-in one");
    return (int)JPP_B200_INVALID_STATE;
  }
  for (u32 n = 0; n < NN; ++n) {
    if (staticT0) t0_node_static(ft, bv, wl, n);
    else t0_node(ft, bv, wl, n);
  }
  // wavefront walk: the kernels of jpp_b200.cu, one "launch" = one loop
  u32 maxNb = 0;
  for (u32 s = 0; s < NS; ++s) maxNb = std::max(maxNb, b->sent_boundary_off[s + 1] - b->sent_boundary_off[s]);
  std::vector<u32> waveNodeOff(maxNb + 2, 0), waveNodes(NN);
  for (u32 n = 0; n < NN; ++n) waveNodeOff[nodeBnd[n] - b->sent_boundary_off[bndSent[nodeBnd[n]]] + 1]++;
  for (u32 i = 0; i <= maxNb; ++i) waveNodeOff[i + 1] += waveNodeOff[i];
  {
    std::vector<u32> cur(waveNodeOff.begin(), waveNodeOff.end() - 1);
    for (u32 n = 0; n < NN; ++n) waveNodes[cur[nodeBnd[n] - b->sent_boundary_off[bndSent[nodeBnd[n]]]]++] = n;
  }
  std::vector<u64> gbKey((size_t)NS * G);
  std::vector<u32> gbLnode((size_t)NS * G), gbT2node((size_t)NS * G), t1node((size_t)NS * G);
  std::vector<u8> gbT1idx((size_t)NS * G), gCount(NS), uCount(NS);
  u32 selCount[3] = {0, 0, 0};
  std::vector<WaveItem> biItems((size_t)(NN + 1) * G), triItems((size_t)(NN + 1) * G);
  std::vector<float> scb((size_t)(NN + 1) * G), tri((size_t)(NN + 1) * G * 2);
  WaveView wv{};
  wv.gbKey = gbKey.data(); wv.gbLnode = gbLnode.data(); wv.gbT2node = gbT2node.data(); wv.gbT1idx = gbT1idx.data();
  wv.t1node = t1node.data(); wv.gCount = gCount.data(); wv.uCount = uCount.data();
  wv.selCount = selCount; wv.biItems = biItems.data(); wv.triItems = triItems.data(); wv.scb = scb.data(); wv.tri = tri.data(); wv.waveNodes = waveNodes.data(); wv.waveNodeOff = waveNodeOff.data();
  for (u32 s = 0; s < NS; ++s) wave_init_sentence(bv, cfg, s);
  for (u32 bnd = 2; bnd < maxNb && cfg.gbeam <= 0; ++bnd)  // full-beam loop
    for (u32 i = waveNodeOff[bnd]; i < waveNodeOff[bnd + 1]; ++i) counters[0] += wave_full_node(ft, bv, cfg, wl, waveNodes[i]);
  for (u32 bnd = 2; bnd < maxNb && cfg.gbeam > 0; ++bnd) {
    selCount[0] = selCount[1] = selCount[2] = 0;
    for (u32 s = 0; s < NS; ++s) wave_gbeam(bv, cfg, wv, s, bnd);
    if (cfg.rcheck > 0)
      for (u32 i = waveNodeOff[bnd]; i < waveNodeOff[bnd + 1]; ++i) wave_prescore(ft, bv, cfg, wv, wl, waveNodes[i]);
    for (u32 s = 0; s < NS; ++s) wave_cutoff(ft, bv, cfg, wv, s, bnd);
    for (u32 i = 0; i < selCount[1]; ++i) wave_item_bigram(ft, bv, wv, wl, biItems[i]);
    for (u32 i = 0; i < selCount[2]; ++i) wave_item_trigram(ft, bv, wv, wl, triItems[i]);
    for (u32 i = waveNodeOff[bnd]; i < waveNodeOff[bnd + 1]; ++i) {
      if (small_gbeam_ok(cfg)) wave_final_node<true>(bv, cfg, wv, waveNodes[i]);
      else wave_final_node<false>(bv, cfg, wv, waveNodes[i]);
    }
  }
  if (model->num_scorers == 2) {
    const jpp_b200_rnn_model* rm = model->rnn;
    RnnTables rt{rm->layer_size, rm->maxent_order, rm->maxent_size, rm->vocab_size, rm->nce_constant, rm->unk_constant,
                 rm->unk_length_penalty, rm->unk_id, rm->matrix, rm->embeddings, rm->nce_embeddings, rm->maxent};
    std::vector<u32> rnOff(NS + 1);
    size_t tot = 0;
    for (u32 s = 0; s < NS; ++s) {
      rnOff[s] = (u32)tot;
      tot += (size_t)G * (b->sent_boundary_off[s + 1] - b->sent_boundary_off[s] + 1) + 4;
    }
    rnOff[NS] = (u32)tot;
    std::vector<i32> id(tot), prev(tot), next(tot), bnd(tot), length(tot), depth(tot), rsRnode(tot), conn((size_t)NN * B, -1),
        bndHead(NB, -1);
    std::vector<u64> hash(tot);
    std::vector<u32> rnCount(NS), rsNode(tot), rsCount(NS);
    std::vector<u16> rsSlot(tot), rsEosG(tot);
    RnnWork rw{};
    rw.rnOff = rnOff.data(); rw.id = id.data(); rw.prev = prev.data(); rw.next = next.data(); rw.bnd = bnd.data();
    rw.length = length.data(); rw.depth = depth.data(); rw.hash = hash.data(); rw.rnCount = rnCount.data();
    rw.rsNode = rsNode.data(); rw.rsSlot = rsSlot.data(); rw.rsEosG = rsEosG.data(); rw.rsRnode = rsRnode.data();
    rw.rsCount = rsCount.data(); rw.connNode = conn.data(); rw.bndHead = bndHead.data();
    std::vector<u32> chainNode(tot); std::vector<u16> chainSlot(tot), chainBnd(tot);
    std::vector<i32> chainId(tot), chainLen(tot);
    rw.chainNode = chainNode.data(); rw.chainSlot = chainSlot.data(); rw.chainBnd = chainBnd.data();
    rw.chainId = chainId.data(); rw.chainLen = chainLen.data();
    const u32 H = rt.H;
    for (u32 s = 0; s < NS; ++s) {
      rnn_build_tree(bv, cfg, rw, s);
      const u32 base0 = rnOff[s], n = rnCount[s];
      const u32 nb = b->sent_boundary_off[s + 1] - b->sent_boundary_off[s];
      if (n < 2) continue;
      std::vector<float> ctx((size_t)n * H);
      for (u32 i = 0; i < H; ++i) {
        float v = 0.0f + rt.emb[i];
        ctx[(size_t)1 * H + i] = 1.0f / (1.0f + expf(-v));
      }
      // nodes are created prev-before-child, so index order is a valid schedule
      for (u32 q = 2; q < n; ++q) {
        if ((u32)bnd[base0 + q] + 1 >= nb) continue;
        const u32 p = (u32)prev[base0 + q] - base0;
        i32 wid = id[base0 + q];
        if (wid == -1) wid = 0;
        for (u32 i = 0; i < H; ++i) {
          float v = serial_dot32(rt.matrix + (size_t)i * H, &ctx[(size_t)p * H], H);
          v = v + rt.emb[(size_t)wid * H + i];
          ctx[(size_t)q * H + i] = 1.0f / (1.0f + expf(-v));
        }
      }
      for (u32 i = 0; i < rsCount[s]; ++i) {
        const u32 si = base0 + i;
        const i32 q = rsRnode[si];
        const i32 wid = id[q];
        float sc;
        if (wid == rt.unkId) {
          sc = rt.unkConstant + rt.unkLengthPenalty * (float)length[q];
        } else {
          const i32 p = prev[q];
          const i32 eid = wid == -1 ? 0 : wid;
          const float d = serial_dot32(rt.nce + (size_t)eid * H, &ctx[(size_t)((u32)p - base0) * H], H);
          sc = rnn_finish_score(rt, d, id[p], wid);
        }
        if (rsEosG[si] != 0xFFFFu) eosRnn[(size_t)s * G + rsEosG[si]] = sc;
        else bv.beamScores[((size_t)rsNode[si] * B + rsSlot[si]) * S + 1] = sc;
      }
      rerank_sentence(bv, cfg, model->score_weights[0], model->score_weights[1], s);
    }
  }
  if (r->best_path_len && r->best_path_nodes)
    for (u32 s = 0; s < NS; ++s) best_path_sentence(bv, cfg, r->best_path_len, r->best_path_nodes, s);
  return 0;
}
