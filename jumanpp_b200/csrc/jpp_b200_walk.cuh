// jpp_b200_walk.cuh -- the fused beam walk (one kernel for all boundaries of all sentences) and
// the compiled-in n-gram program.  Included by jpp_b200.cu after the group-cooperative device
// functions (gbeam_group, cutoff_group_block, block_max_u32).
//
// STATIC = the n-gram program is the compiled-in one (jpp_b200_static_jumandic.h, the analogue
// of the reference's generated PartialNgramFeatureApply): the 112-byte pattern rows of the right
// node / t1 / t2 are loaded ONCE per item into registers with 16-byte loads and indexed by
// literals, where the table-driven form loads two or three 8-byte slots per feature through L1
// (ncu r2: 8.96 G sector requests for 4.19 G weight gathers -- more than half of the L1 traffic
// of the walk was pattern slots).  Sums replay exactly the accumulator orders of gsum_triple /
// gsum_dual (same additions per accumulator, same final association).
#pragma once

#include "jpp_b200_static_jumandic.h"

namespace {

using namespace jppb200;

#ifndef JPP_FUSED_MINBLOCKS_STATIC
#define JPP_FUSED_MINBLOCKS_STATIC 4   // blocks of 256 threads per SM (measured, walk of the bench batch: 2 -> 42.5 ms, 3 -> 39.2 ms, 4 -> 37.5 ms)
#endif

// ------------------------------------------------------------ rows in registers ----
template <int N>
__device__ __forceinline__ void load_row(const u64* __restrict__ base, u64 (&r)[N]) {
  static_assert(N % 2 == 0, "rows are read as 16-byte pairs");
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(base);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const ulonglong2 v = __ldg(p + i);
    r[2 * i] = v.x;
    r[2 * i + 1] = v.y;
  }
}

typedef JumandicNgramProgram SP;

#ifndef JPP_STATIC_BATCH
#define JPP_STATIC_BATCH 8
#endif
constexpr int kGatherBatch = JPP_STATIC_BATCH;  // table reads a thread has in flight in the bigram sums (a multiple of 8)
static_assert(kGatherBatch % 8 == 0, "accumulator j % NV / j % 2 / j % 4 of feature f = batch * k + j needs 8 | batch");

struct LeftRow {
  u64 r[SP::kStored];
  __device__ __forceinline__ void load(const u64* __restrict__ base) { load_row(base, r); }
  __device__ __forceinline__ u64 operator[](int s) const { return r[s]; }
};

// One pass over the bigram features of (right node row P, left node row T1): the NV interleaved
// accumulators of the generated applyBiStep2 rows, the r1/r2 pair of applyBiTriFullKernel and the
// four of computeUnrolled4RawPerceptron, fed from the same gathered values (gsum_triple's orders).
__device__ __forceinline__ void static_bigram_sums(const float* __restrict__ w, u32 mask, const u64 (&P)[SP::kStored],
                                                   const LeftRow& T1, float& sumNV, float& sumPair, float& sum4) {
  constexpr int NV = SP::kBiVars;
  float a[NV], c[2] = {0.0f, 0.0f}, b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = 0.0f;
#pragma unroll
  for (int B = 0; B < (SP::kBi + kGatherBatch - 1) / kGatherBatch; ++B) {
    u32 id[kGatherBatch];
    float x[kGatherBatch];
#define X(f, st, s0, s1) \
  if ((f) / kGatherBatch == B) id[(f) % kGatherBatch] = (u32)mix(mix(st, P[s0]), T1[s1]) & mask;
    JPP_STATIC_BIGRAMS(X)
#undef X
#pragma unroll
    for (int j = 0; j < kGatherBatch; ++j)
      if (B * kGatherBatch + j < SP::kBi) x[j] = gather_weight(w, id[j]);
#pragma unroll
    for (int j = 0; j < kGatherBatch; ++j)
      if (B * kGatherBatch + j < SP::kBi) {  // feature B*batch+j: accumulator j % NV / j % 2 / j % 4 (the batch is a multiple of each)
        a[j % NV] += x[j];
        c[j % 2] += x[j];
        b[j % 4] += x[j];
      }
  }
  sumNV = a[0];
#pragma unroll
  for (int j = 1; j < NV; ++j) sumNV += a[j];
  sumPair = c[0] + c[1];
  sum4 = ((b[0] + b[1]) + b[2]) + b[3];
}

// The same for the trigram features of (P, T1, T2); c0 / c1 stay apart for applyBiTri (scbuf + r1 + r2).
__device__ __forceinline__ void static_trigram_sums(const float* __restrict__ w, u32 mask, const u64 (&P)[SP::kTriSlots],
                                                    const u64 (&T1)[SP::kTriSlots], const u64 (&T2)[SP::kTriSlots], float& sumNV,
                                                    float& c0, float& c1, float& sum4) {
  constexpr int NV = SP::kTriVars;
  float a[NV], c[2] = {0.0f, 0.0f}, b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = 0.0f;
#pragma unroll
  for (int B = 0; B < (SP::kTri + 7) / 8; ++B) {
    u32 id[8];
    float x[8];
#define X(f, st, s0, s1, s2) \
  if ((f) / 8 == B) id[(f) % 8] = (u32)mix(mix(mix(st, P[s0]), T1[s1]), T2[s2]) & mask;
    JPP_STATIC_TRIGRAMS(X)
#undef X
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (B * 8 + j < SP::kTri) x[j] = gather_weight(w, id[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (B * 8 + j < SP::kTri) {
        a[j % NV] += x[j];
        c[j % 2] += x[j];
        b[j % 4] += x[j];
      }
  }
  sumNV = a[0];
#pragma unroll
  for (int j = 1; j < NV; ++j) sumNV += a[j];
  c0 = c[0];
  c1 = c[1];
  sum4 = ((b[0] + b[1]) + b[2]) + b[3];
}

// W2 with the rows in registers: the same results as wave_prescore
__device__ __forceinline__ void wave_prescore_static(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg,
                                                     const WaveView& wv, const float* __restrict__ w, u32 node) {
  const u32 gB = bv.nodeBnd[node];
  const u32 sent = bv.bndSent[gB];
  const u32 G = wv.gCount[sent], U = wv.uCount[sent];
  const u32 theMax = (u32)cfg.rcheck < G ? (u32)cfg.rcheck : G;
  if (theMax == 0) return;
  const u32 GMAX = (u32)cfg.gbeam, mask = ft.mask;
  const bool lastRow = ft.dynamicOrder || (node + 1 == bv.bndNodeOff[gB + 1]);
  u64 P[SP::kStored];
  load_row(bv.patterns + (size_t)node * SP::kStored, P);
  u64 P3[SP::kTriSlots];
#pragma unroll
  for (int i = 0; i < SP::kTriSlots; ++i) P3[i] = P[i];
  const float t0 = bv.t0[node];
  for (u32 i = 0; i < theMax; ++i) {
    LeftRow T1;
    u64 T1t[SP::kTriSlots], T2[SP::kTriSlots];
    T1.load(bv.patterns + (size_t)wv.gbLnode[(size_t)sent * GMAX + i] * SP::kStored);
    load_row(bv.patterns + (size_t)wv.gbT2node[(size_t)sent * GMAX + i] * SP::kStored, T2);
#pragma unroll
    for (int q = 0; q < SP::kTriSlots; ++q) T1t[q] = T1[q];
    float sc = t0;
    float sNV, sPair, s4;
    static_bigram_sums(w, mask, P, T1, sNV, sPair, s4);
    sc += lastRow ? s4 : sNV;
    if (i == 0) bv.bi0[node] = (U == 1 || ft.dynamicOrder) ? s4 : sPair;  // applyBiTriFullKernel row of unique t1 0
    float tNV, c0, c1, t4;
    static_trigram_sums(w, mask, P3, T1t, T2, tNV, c0, c1, t4);
    sc += lastRow ? t4 : tNV;
    bv.pre[(size_t)i * bv.totalNodes + node] = sc;
  }
}

// W4 items with the rows in registers: the same results as wave_item_bigram / wave_item_trigram
__device__ __forceinline__ void wave_item_bigram_static(const FeatTables& ft, const BatchView& bv, const WaveView& wv,
                                                        const float* __restrict__ w, const WaveItem& it) {
  u64 P[SP::kStored];
  LeftRow T1;
  load_row(bv.patterns + (size_t)it.a * SP::kStored, P);
  T1.load(bv.patterns + (size_t)it.b * SP::kStored);
  float sNV, sPair, s4;
  static_bigram_sums(w, ft.mask, P, T1, sNV, sPair, s4);
  wv.scb[it.slot & ~ITEM_LAST] = ((it.slot & ITEM_LAST) || ft.dynamicOrder) ? s4 : sPair;
}

__device__ __forceinline__ void wave_item_trigram_static(const FeatTables& ft, const BatchView& bv, const WaveView& wv,
                                                         const float* __restrict__ w, const WaveItem& it) {
  u64 P[SP::kTriSlots], T1[SP::kTriSlots], T2[SP::kTriSlots];
  load_row(bv.patterns + (size_t)it.a * SP::kStored, P);
  load_row(bv.patterns + (size_t)it.b * SP::kStored, T1);
  load_row(bv.patterns + (size_t)it.c * SP::kStored, T2);
  float tNV, c0, c1, t4;
  static_trigram_sums(w, ft.mask, P, T1, T2, tNV, c0, c1, t4);
  float* out = wv.tri + 2 * (size_t)(it.slot & ~ITEM_LAST);
  if ((it.slot & ITEM_LAST) || ft.dynamicOrder) { out[0] = t4; out[1] = 0.0f; }
  else { out[0] = c0; out[1] = c1; }
}

// ------------------------------------------------------------------- the kernel ----
// Fused walk: a block of 256 threads owns 32 sentences (neighbours in the longest-first order, so
// of similar length) and walks them through ALL their boundaries: the device functions of the
// wavefront kernels, __syncthreads() where those have a kernel boundary, the lists of selected
// nodes / items private to the block (counters in shared memory, no global atomics).  Blocks drift
// apart in phase, so while one block chases the dependent loads of its global beams another one
// keeps the gather pipe busy -- across kernels that overlap does not exist.  All warps of a block
// are in the same phase at the same time, which also keeps the instruction working set of an SM to
// a few phases.  A variant in which every WARP walked its own four sentences (`__syncwarp` only, no
// block barriers at all; git history: "Warp-scope beam walk") was measured at 128 / 85 / 64 registers
// per thread and lost every time -- 69.6 / 60.4 / 57.6 ms against 37.4 ms (profiles/README.md r2b,
// r2g): without other warps of the same block to run, every lane-0 serial section (the exact
// nth_element replay, the beam sorts) idles 31 lanes for its whole length.
//
// GW = lanes per sentence in W1 (8 when beam and global beam are <= 8: the block's 32 sentences in
// one pass; 32 for wider beams: four passes of eight sentences), KB = beam slots a lane holds in
// W1, SMALL = register beam build of W5 (global beam <= 8), STATIC = compiled-in n-gram program.
#ifndef JPP_FUSED_SENT_DEFAULT
#define JPP_FUSED_SENT_DEFAULT 32
#endif
constexpr u32 kFusedSentDefault = JPP_FUSED_SENT_DEFAULT;  // sentences per block (8 threads each): 32 or 16
#ifndef JPP_FUSED_MINBLOCKS
#define JPP_FUSED_MINBLOCKS 4      // table-driven program: the same
#endif

template <int GW, int KB, bool SMALL, bool STATIC, int SENT>
__global__ void __launch_bounds__(SENT * 8, (STATIC ? JPP_FUSED_MINBLOCKS_STATIC : JPP_FUSED_MINBLOCKS) * (32 / SENT))
    k_walk_fused(const __grid_constant__ FeatTables ft, BatchView bv, ScoreCfg cfg, WaveView wv, const float* __restrict__ w,
                 u32 itemCap) {
  constexpr u32 kFusedSent = SENT, kFusedThreads = SENT * 8;
  __shared__ u32 sCount[4];                 // selected nodes, bigram items, trigram items of the current wave
  __shared__ u32 sNodeOff[kFusedSent + 1];  // right nodes of the block's sentences at the current boundary (prefix)
  __shared__ u32 sNode0[kFusedSent];
  __shared__ u32 sSent[kFusedSent], sB0[kFusedSent], sNb[kFusedSent];
  const u32 NS = bv.numSentences;
  const u32 firstSent = blockIdx.x * kFusedSent;
  if (firstSent >= NS) return;
  const u32 nHere = NS - firstSent < kFusedSent ? NS - firstSent : kFusedSent;
  // the block's private lists
  WaveView lw = wv;
  lw.selCount = sCount;
  lw.biItems = wv.biItems + (size_t)blockIdx.x * itemCap;
  lw.triItems = wv.triItems + (size_t)blockIdx.x * itemCap;
  lw.scb = wv.scb + (size_t)blockIdx.x * itemCap;
  lw.tri = wv.tri + 2 * (size_t)blockIdx.x * itemCap;
  const WLoadDirect wl{w};
  u32 myNb = 0;
  if (threadIdx.x < kFusedSent) {
    u32 s = 0xFFFFFFFFu, b0 = 0;
    if (threadIdx.x < nHere) {
      s = bv.sentOrder[firstSent + threadIdx.x];
      myNb = sent_span(bv, s, b0);
    }
    sSent[threadIdx.x] = s;
    sB0[threadIdx.x] = b0;
    sNb[threadIdx.x] = myNb;
  }
  // longest sentence of the block (the first one in the longest-first order, unless that one failed
  // validation on the device after the order was made)
  const u32 maxNb = block_max_u32(myNb, sCount);
  constexpr u32 kGroups = kFusedThreads / GW;  // sentences W1 takes per pass
  const u32 wgrp = threadIdx.x / GW, wsub = threadIdx.x % GW;
  for (u32 bnd = 2; bnd < maxNb; ++bnd) {
    // W1 + the block's node list of this wave
    if (threadIdx.x < 3) sCount[threadIdx.x] = 0;
#pragma unroll 1
    for (u32 q = wgrp; q < kFusedSent; q += kGroups) {
      const u32 sent = sSent[q], nb = sNb[q];
      u32 R = 0, n0 = 0;
      if (sent != 0xFFFFFFFFu && bnd < nb && nb > 3) {
        n0 = bv.bndNodeOff[sB0[q] + bnd];
        R = bv.bndNodeOff[sB0[q] + bnd + 1] - n0;
        gbeam_group<KB, GW>(bv, cfg, lw, sent, bnd);
      }
      if (wsub == 0) { sNodeOff[q + 1] = R; sNode0[q] = n0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      u32 v = threadIdx.x < kFusedSent ? sNodeOff[threadIdx.x + 1] : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if ((int)threadIdx.x >= d) v += t;
      }
      if (threadIdx.x < kFusedSent) sNodeOff[threadIdx.x + 1] = v;
      if (threadIdx.x == 0) sNodeOff[0] = 0;
    }
    __syncthreads();
    const u32 T = sNodeOff[kFusedSent];
    // W2
    for (u32 i = threadIdx.x; i < T; i += kFusedThreads) {
      u32 lo = 0, hi = kFusedSent;  // sentence j with sNodeOff[j] <= i < sNodeOff[j + 1]
      while (hi - lo > 1) {
        const u32 m = (lo + hi) >> 1;
        if (sNodeOff[m] <= i) lo = m; else hi = m;
      }
      const u32 node = sNode0[lo] + (i - sNodeOff[lo]);
      if (STATIC) wave_prescore_static(ft, bv, cfg, lw, w, node);
      else wave_prescore(ft, bv, cfg, lw, wl, node);
    }
    __syncthreads();
    // W3 (positions and items are private to the block: counters in shared memory)
    cutoff_group_block<kFusedSent>(ft, bv, cfg, lw, firstSent + nHere, bnd, firstSent);
    __syncthreads();
    // W4
    {
      const u32 n1 = sCount[1], n2 = sCount[2];
      for (u32 i = threadIdx.x; i < n1 + n2; i += kFusedThreads) {
        if (i < n1) {
          if (STATIC) wave_item_bigram_static(ft, bv, lw, w, lw.biItems[i]);
          else wave_item_bigram(ft, bv, lw, wl, lw.biItems[i]);
        } else {
          if (STATIC) wave_item_trigram_static(ft, bv, lw, w, lw.triItems[i - n1]);
          else wave_item_trigram(ft, bv, lw, wl, lw.triItems[i - n1]);
        }
      }
    }
    __syncthreads();
    // W5
    for (u32 i = threadIdx.x; i < T; i += kFusedThreads) {
      u32 lo = 0, hi = kFusedSent;
      while (hi - lo > 1) {
        const u32 m = (lo + hi) >> 1;
        if (sNodeOff[m] <= i) lo = m; else hi = m;
      }
      wave_final_node<SMALL>(bv, cfg, lw, sNode0[lo] + (i - sNodeOff[lo]));
    }
    __syncthreads();
  }
}

}  // namespace
