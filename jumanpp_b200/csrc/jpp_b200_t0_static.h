// jpp_b200_t0_static.h -- T0 with the compiled-in Jumandic program (jpp_b200_static_jumandic_t0.h).
//
// Same results as t0_node (jpp_b200_kernels.h), bit for bit: the same primitive values, the same hash chains,
// the same table reads summed in the same order -- but the program is code, not tables.  The table-driven
// t0_node executes ~35 000 instructions per node (every pattern reference is a walk through five small
// tables and a read of the local primitive array; ncu r2e: 51 % of k_t0's stall samples sit on that read and
// the multiply that waits for it), and a 4 MB perceptron table that lives in L2 makes it only 0.9 ms faster
// (profiles/README.md r2i): it is bound by its own instruction stream, not by the gathers.  Here primitives and
// patterns are registers, dead as soon as the row slot / unigram index they feed is out, and all 32 unigram
// reads of a node are in flight together.
#pragma once
#include "jpp_b200_kernels.h"
#include "jpp_b200_static_jumandic_t0.h"

namespace jppb200 {

template <typename WL>
JPP_HD void t0_node_static(const FeatTables& ft, const BatchView& bv, const WL& wl, u32 node) {
  typedef JumandicT0Program P;
  static_assert(P::kUni >= 8, "uniVars = 4 below (pattern_feature_codegen.cc:202-208)");
  const u32 gb = bv.nodeBnd[node];
  if (gb == 0xFFFFFFFFu) return;  // node of no valid boundary (failed sentence)
  const u32 sent = bv.bndSent[gb];
  const u32 localBnd = gb - bv.sentBndOff[sent];
  u64* pout = bv.patterns + (size_t)node * bv.patNodeStride;
  const size_t pstr = bv.patPStride;
  const jpp_b200_node_info ni = bv.nodeInfo[node];
  if (localBnd < 2) {  // BOS rows, lattice_builder.cc:171-176
    for (u32 p = 0; p < (u32)P::kStored; ++p) pout[p * pstr] = (u64)(u32)ni.entry_ptr;
    bv.t0[node] = 0.0f;
    return;
  }
  const u32 cp0 = bv.sentCpOff[sent];
  const u32 ncp = bv.sentCpOff[sent + 1] - cp0;
  const jpp_b200_codepoint* cps = bv.cps + cp0;
  const NodeEntry ne = node_entry(ft, bv, node, sent, ni.entry_ptr);
  u32 id[P::kUni];
  P::run(ne.entry, ne.ph, ni, cps, ncp, ft.mask, pout, pstr, id);
  // unigram score: the order of t0_node (rows 0..R-2: "=" for the first four terms then "+=", summed p0 + p1 + ...;
  // the last row / the dynamic order: computeUnrolled4RawPerceptron, four accumulators started at 0)
  const bool lastRow = ft.dynamicOrder || (node + 1 == bv.bndNodeOff[gb + 1]);
  float wv[P::kUni];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int u = 0; u < P::kUni; ++u) wv[u] = wl.one(id[u]);
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int u = 0; u < P::kUni; ++u) {
    if (!lastRow && u < 4) acc[u % 4] = wv[u];
    else acc[u % 4] += wv[u];
  }
  float s = acc[0];
  for (int v = 1; v < 4; ++v) s = s + acc[v];
  bv.t0[node] = s;
}

}  // namespace jppb200
