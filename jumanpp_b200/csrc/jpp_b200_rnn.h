// jpp_b200_rnn.h -- RNN language-model rescoring of the EOS global beam and the
// final re-rank (reference analyzer_impl.cc:286-294).
//
// Stages (one kernel each, see jpp_b200.cu):
//   R1 rnn_build_tree   one thread per sentence: RnnIdResolver::resolveIdsAtGbeam
//                       + RnnIdContainer prefix tree (rnn_id_resolver.cc:173-276),
//                       integer only, including its collapse quirk;
//   R2 rnn_contexts     level by level (depth from BOS): ctx' = sigmoid(W^T ctx + emb[id])
//                       (mikolov_rnn_impl.h:202-214), one warp per tree node;
//   R3 rnn_scores       one warp per connection: nce[id].ctx + maxent - nce constant
//                       (mikolov_rnn_impl.h:216-257, index hash :98-131);
//   R4 rerank_sentence  one thread per sentence: adjustBeamScores + remakeEosBeam
//                       (score_processor.cc:521-576).
//
// Dot products use the order of oracle/eigen_shim (which stands in for the
// un-vendored Eigen): partial l sums k = l, l+32, ... with one fused multiply-add
// per term, then the 16, 8, 4, 2, 1 fold of the 32 partials.
#pragma once
#include "jpp_b200_kernels.h"

#if defined(__CUDACC__)
#include <math.h>
#else
#include <cmath>
#endif

namespace jppb200 {

// Per-batch RNN workspace (device arrays; node slots are per sentence:
// sentence s owns [rnOff[s], rnOff[s+1]) with capacity G*(nb+1)+4).
struct RnnWork {
  const u32* rnOff;   // [sentences+1]
  i32* id;            // RnnNode::id
  i32* prev;          // absolute rnode index or -1
  i32* next;          // nextInBnd (absolute) or -1
  i32* bnd;           // sentence-local boundary
  i32* length;
  i32* depth;         // words from BOS (bos1 = 0)
  u64* hash;
  u32* rnCount;       // [sentences] nodes created
  // score list (same capacity / offsets)
  u32* rsNode;        // lattice node of the connection (or 0 for EOS)
  u16* rsSlot;
  u16* rsEosG;        // 0xFFFF, or the EOS global-beam index
  i32* rsRnode;
  u32* rsCount;       // [sentences]
  i32* connNode;      // [nodes][B]  ptrCache_: connection -> rnode, -1 = unseen
  i32* bndHead;       // [boundaries] head of the published list of the boundary
  u32* chainNode;     // [rnode capacity] scratch: connections of the path being added
  u16* chainSlot;
  u16* chainBnd;
  i32* chainId;   // RNN id and surface length of the chain's nodes, fetched while walking back
  i32* chainLen;
  u32* scoreList;     // [rnode capacity] flat list of score entries of the whole batch
  u32* scoreTotal;    // [1]
  float* ctx;         // [total rnodes][H]
  // level lists
  u32* levelCount;    // [maxDepth+2]
  u32* levelOff;      // [maxDepth+2]
  u32* levelList;     // [total rnodes]
  u32 maxDepth;
};

// FastHash1::mix (util/fast_hash.h:46-51)
JPP_HD u64 fast_hash1_mix(u64 state, u64 data) {
  u64 v = (state ^ data) * kMult;
  return v ^ (v >> 32);
}

// R1: literal restatement of addPath / addPrevChain, iterative instead of recursive.
JPP_HD void rnn_build_tree(const BatchView& bv, const ScoreCfg& cfg, const RnnWork& rw, u32 sent) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam;
  const u32 base0 = rw.rnOff[sent];
  rw.rnCount[sent] = 0;
  rw.rsCount[sent] = 0;
  if (nb <= 3) return;
  const u32 eosBnd = nb - 1;
  const u32 eosG = bv.gbCount[b0 + eosBnd];
  const u32 eosE0 = bv.bndEndOff[b0 + eosBnd];
  u32 nrn = 0, nrs = 0;
  // addBos (:311-329)
  rw.id[base0] = 0; rw.prev[base0] = -1; rw.next[base0] = -1; rw.bnd[base0] = 0; rw.length[base0] = 0;
  rw.depth[base0] = -1; rw.hash[base0] = 0;
  rw.id[base0 + 1] = 0; rw.prev[base0 + 1] = (i32)base0; rw.next[base0 + 1] = -1; rw.bnd[base0 + 1] = 1;
  rw.length[base0 + 1] = 0; rw.depth[base0 + 1] = 0; rw.hash[base0 + 1] = 0xdeadbeef0000ULL;
  nrn = 2;
  rw.bndHead[b0 + 1] = (i32)(base0 + 1);
  const u32 bos1 = bv.bndNodeOff[b0 + 1];
  rw.connNode[(size_t)bos1 * B] = (i32)(base0 + 1);
  for (u32 g = 0; g < eosG; ++g) {
    const u64 key = bv.eosGbeam[(size_t)sent * GMAX + g];
    const u32 left0 = (u32)(key >> 16) & 0xFFFFu;
    // addPrevChain: walk back to the first connection already in ptrCache_,
    // remembering the new connections (deepest last)
    u32 len = 0;
    i32 first, prev;
    {
      u32 curNode = bv.endNode[eosE0 + left0];
      u32 curBnd = bv.ends[eosE0 + left0].boundary;
      u32 curSlot = (u32)key & 0xFFFFu;
      // the connection's own record and its cache entry are independent loads: issue both, then
      // decide (one round trip per step of the walk back instead of two)
      i32 known;
      for (;;) {
        known = rw.connNode[(size_t)curNode * B + curSlot];
        const jpp_b200_beam_element el = bv.beam[(size_t)curNode * B + curSlot];
        if (known != -1) break;
        rw.chainNode[base0 + len] = curNode;
        rw.chainSlot[base0 + len] = (u16)curSlot;
        rw.chainBnd[base0 + len] = (u16)curBnd;
        {  // id / length of this node: off the critical path here, ready when the chain is unwound
          const jpp_b200_node_info ni = bv.nodeInfo[curNode];
          const i32 nid = node_rnn_id(bv, curNode, sent, ni.entry_ptr);
          rw.chainId[base0 + len] = nid;
          rw.chainLen[base0 + len] = (i32)ni.end - (i32)ni.start;
        }
        ++len;
        const u32 e = bv.bndEndOff[b0 + curBnd] + el.left;
        curSlot = el.prev_beam;
        curBnd = bv.ends[e].boundary;
        curNode = bv.endNode[e];
      }
      first = prev = known;
    }
    // unwind: deepest new connection first, the EOS fake connection (k == -1) last; hash and
    // depth of `prev` travel in registers (they were just computed unless the step collapsed)
    u64 prevHash = rw.hash[prev];
    i32 prevDepth = rw.depth[prev];
    // nodes created since the last collapse of this path: exactly the ones addPath publishes
    u32 sinceCollapse = 0;
    // record of step k, loaded one step ahead so that its latency hides behind the list scan
    u32 nNode = 0, nSlot = 0, nBnd = 0;
    i32 nId = 0, nLen = 0;
    if (len > 0) {
      const u32 j = base0 + len - 1;
      nNode = rw.chainNode[j]; nSlot = rw.chainSlot[j]; nBnd = rw.chainBnd[j]; nId = rw.chainId[j]; nLen = rw.chainLen[j];
    }
    for (i32 k = (i32)len - 1; k >= -1; --k) {
      u32 cNode = 0, cSlot = 0, cb;
      i32 id, length;
      if (k >= 0) {
        cNode = nNode; cSlot = nSlot; cb = nBnd; id = nId; length = nLen;
        if (k > 0) {
          const u32 j = base0 + (u32)k - 1;
          nNode = rw.chainNode[j]; nSlot = rw.chainSlot[j]; nBnd = rw.chainBnd[j]; nId = rw.chainId[j]; nLen = rw.chainLen[j];
        }
      } else {
        id = 0; length = 0; cb = eosBnd;  // nodeCache_[{eos,0}] = {eos,0,0} (:306-308)
      }
      const u64 data = (u64)(u32)id | ((u64)(u16)length << 32);
      const u64 hv = fast_hash1_mix(prevHash, data);
      // crdCache_.find(coord): newest published node with this (boundary, length, id)
      i32 it = -1;
      for (i32 q = rw.bndHead[b0 + cb]; q != -1; q = rw.next[q])
        if (rw.length[q] == length && rw.id[q] == id) { it = q; break; }
      bool collapsed = false;
      if (it != -1)
        for (i32 q = it; q != -1; q = rw.next[q])
          if (rw.hash[q] == hv) { collapsed = true; break; }
      i32 mine;
      if (collapsed) {
        mine = it;  // "nodePtr = it->second" (:224-231)
        const u32 si = base0 + nrs++;
        rw.rsNode[si] = cNode; rw.rsSlot[si] = (u16)cSlot; rw.rsEosG[si] = k < 0 ? (u16)g : (u16)0xFFFFu;
        rw.rsRnode[si] = mine;
        first = mine;
        prevHash = rw.hash[mine];
        prevDepth = rw.depth[mine];
        sinceCollapse = 0;
      } else {
        mine = (i32)(base0 + nrn++);
        rw.id[mine] = id; rw.prev[mine] = prev; rw.next[mine] = -1; rw.bnd[mine] = (i32)cb;
        rw.length[mine] = length; rw.hash[mine] = hv; rw.depth[mine] = prevDepth + 1;
        prevHash = hv;
        prevDepth = prevDepth + 1;
        ++sinceCollapse;
        if (cb + 1 < nb && rw.levelCount) {  // a context is needed for every node before EOS
#if defined(__CUDA_ARCH__)
          atomicAdd(rw.levelCount + prevDepth, 1u);
#else
          rw.levelCount[prevDepth] += 1;
#endif
        }
      }
      if (k >= 0) rw.connNode[(size_t)cNode * B + cSlot] = mine;
      prev = mine;
    }
    // publish (addPath :247-266): from the tip back to `first`.  New nodes of a path get
    // consecutive ids and link to each other, so the walk over rw.prev is a count-down, and the
    // boundary of the node made in unwind step k is chainBnd[k]: no dependent loads left here.
    {
      const i32 tip = (i32)(base0 + nrn) - 1;
      for (u32 j = 0; j < sinceCollapse; ++j) {
        const i32 lastN = tip - (i32)j;
        const i32 k = (i32)j - 1;
        const u32 cb = k < 0 ? eosBnd : (u32)rw.chainBnd[base0 + k];
        rw.next[lastN] = rw.bndHead[b0 + cb];
        rw.bndHead[b0 + cb] = lastN;
        const u32 si = base0 + nrs++;
        if (k < 0) {
          rw.rsNode[si] = 0; rw.rsSlot[si] = 0; rw.rsEosG[si] = (u16)g;
        } else {
          rw.rsNode[si] = rw.chainNode[base0 + k]; rw.rsSlot[si] = rw.chainSlot[base0 + k]; rw.rsEosG[si] = (u16)0xFFFFu;
        }
        rw.rsRnode[si] = lastN;
      }
      (void)first;
    }
  }
  rw.rnCount[sent] = nrn;
  rw.rsCount[sent] = nrs;
  if (rw.scoreList && nrs) {  // flat list of this batch's score entries
    u32 base;
#if defined(__CUDA_ARCH__)
    base = atomicAdd(rw.scoreTotal, nrs);
#else
    base = *rw.scoreTotal;
    *rw.scoreTotal += nrs;
#endif
    for (u32 i = 0; i < nrs; ++i) rw.scoreList[base + i] = base0 + i;
  }
}

// dot product, oracle/eigen_shim order (serial form; the kernels keep the 32 partials in registers)
JPP_HD float serial_dot32(const float* a, const float* b, u32 n) {
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (u32 k = 0; k < n; ++k) {
#if defined(__CUDA_ARCH__)
    part[k & 31] = __fmaf_rn(a[k], b[k], part[k & 31]);
#else
    part[k & 31] = fmaf(a[k], b[k], part[k & 31]);
#endif
  }
  for (int stride = 16; stride >= 1; stride >>= 1)
    for (int l = 0; l < stride; ++l) part[l] = part[l] + part[l + stride];
  return part[0];
}

// maxent part of one score (mikolov_rnn_impl.h:98-131,225-237); every context
// slot holds the previous word id (rnn_scorer_gbeam.cc:170-186)
JPP_HD float rnn_maxent(const RnnTables& rt, i32 prevId, i32 id) {
  const u64 PR[36] = {108641969, 116049371, 125925907, 133333309, 145678979, 175308587, 197530793, 234567803, 251851741,
                      264197411, 330864029, 399999781, 407407183, 459258997, 479012069, 545678687, 560493491, 607407037,
                      629629243, 656789717, 716048933, 718518067, 725925469, 733332871, 753085943, 755555077, 782715551,
                      790122953, 812345159, 814814293, 893826581, 923456189, 940740127, 953085797, 985184539, 990122807};
  const u64 hashMax = rt.maxentSize - rt.vocab;
  float me = 0.0f;
  for (u32 i = 0; i < rt.order; ++i) {
    u64 x = PR[0] * PR[1];
    for (u32 j = 1; j <= i; ++j) {
      const u64 primeIdx = ((u64)i * PR[j] + j) % 36;
      x += PR[primeIdx] * ((u64)(int64_t)prevId + 1);
    }
    const u64 idx = ((x % hashMax) + (u64)(int64_t)id) % hashMax;
    const float wv = rt.maxent[idx];
    me = (i == 0) ? wv : me + wv;
  }
  return me;
}

JPP_HD float rnn_finish_score(const RnnTables& rt, float dot, i32 prevId, i32 id) {
  float s = dot;
  if (rt.order > 0) s = s + rnn_maxent(rt, prevId, id); else s = 0.0f;
  return s - rt.nceConstant;
}

// remakeEosBeam (analyzer_impl.cc:286-294): the EOS row from the re-scored global beam
JPP_HD void rerank_eos(const BatchView& bv, const ScoreCfg& cfg, float w0, float w1, u32 sent) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam, S = cfg.S;
  const u32 eosB = b0 + nb - 1;
  const u32 eosNode = bv.bndNodeOff[eosB];
  const u32 eosE0 = bv.bndEndOff[eosB];
  const u32 G = bv.gbCount[eosB];
  float full[MAX_BEAM], loc[MAX_BEAM], rnn[MAX_BEAM];
  u64 gb2[MAX_BEAM];
  for (u32 g = 0; g < G; ++g) {
    const u64 key = bv.eosGbeam[(size_t)sent * GMAX + g];
    const u32 left = (u32)(key >> 16) & 0xFFFFu, bi = (u32)key & 0xFFFFu;
    const u32 lnode = bv.endNode[eosE0 + left];
    const float beamScore = bv.beam[(size_t)lnode * B + bi].total_score;
    gb2[g] = pack_cand(beamScore, left, bi);
    const float s0 = bv.eosLocal[(size_t)sent * GMAX + g];
    const float s1 = bv.eosRnn[(size_t)sent * GMAX + g];
    float localScore = 0.0f;
    localScore += s0 * w0;
    localScore += s1 * w1;
    loc[g] = s0;
    rnn[g] = s1;
    full[g] = localScore + beamScore;
  }
  u32 idx[MAX_BEAM];
  const int len = beam_order(idx, (int)G, full, (int)B);
  jpp_b200_beam_element* row = bv.beam + (size_t)eosNode * B;
  float* srow = bv.beamScores + (size_t)eosNode * B * S;
  int k = 0;
  for (; k < len && k < (int)B; ++k) {
    const u32 g = idx[k];
    jpp_b200_beam_element e;
    e.left = (u16)(gb2[g] >> 16);
    e.prev_beam = (u16)gb2[g];
    e.total_score = full[g];
    row[k] = e;
    srow[(size_t)k * S] = loc[g];
    srow[(size_t)k * S + 1] = rnn[g];
  }
  for (; k < (int)B; ++k) {
    row[k] = fake_elem();
    srow[(size_t)k * S] = 0.0f;
    srow[(size_t)k * S + 1] = 0.0f;
  }
}

// R4: adjustBeamScores + remakeEosBeam, one thread per sentence
JPP_HD void rerank_sentence(const BatchView& bv, const ScoreCfg& cfg, float w0, float w1, u32 sent) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  if (nb <= 3) return;
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam, S = cfg.S;
  for (u32 bnd = 3; bnd < nb; ++bnd) {
    const u32 gB = b0 + bnd;
    if (bv.bndNodeOff[gB + 1] == bv.bndNodeOff[gB]) continue;
    const u32 G = bv.gbCount[gB];
    for (u32 i = 0; i < G; ++i) {
      const u32 node = bv.gbNode[(size_t)gB * GMAX + i];
      const u32 slot = bv.gbSlot[(size_t)gB * GMAX + i];
      const u32 nbnd = bv.gbLeftBnd[(size_t)gB * GMAX + i];
      jpp_b200_beam_element* el = bv.beam + (size_t)node * B + slot;
      const float* sc = bv.beamScores + ((size_t)node * B + slot) * S;
      float localScore = 0.0f;
      localScore += sc[0] * w0;
      localScore += sc[1] * w1;
      const u32 pnode = bv.endNode[bv.bndEndOff[b0 + nbnd] + el->left];
      localScore += bv.beam[(size_t)pnode * B + el->prev_beam].total_score;
      el->total_score = localScore;
    }
  }
  rerank_eos(bv, cfg, w0, w1, sent);
}

}  // namespace jppb200
