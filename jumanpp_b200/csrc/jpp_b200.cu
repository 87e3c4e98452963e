// jpp_b200.cu -- C ABI (include/jpp_b200.h) + CUDA kernels of the B200 lattice
// scorer.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false
// (see jumanpp_b200/build.py).  There is no CPU path in this library: every
// entry point that scores needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "jpp_b200.h"
#include "jpp_b200_host.h"
#include "jpp_b200_kernels.h"
#include "jpp_b200_rnn.h"
#include "jpp_b200_t0_static.h"

using namespace jppb200;

namespace {

thread_local std::string g_create_error;

// ------------------------------------------------------------- kernels ------
__global__ void k_fill_weights(float* w, u64 n, u64 seed, float scale) {
  const float f = (float)((double)scale * 1.7320508075688772);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 x = i + (seed << 40);
    x += 0x9E3779B97F4A7C15ULL;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    const u32 s = (u32)(z & 0xFFFF) + (u32)((z >> 16) & 0xFFFF) + (u32)((z >> 32) & 0xFFFF) + (u32)((z >> 48) & 0xFFFF);
    const float v = __fsub_rn(__fmul_rn((float)s, 1.0f / 65536.0f), 2.0f);
    w[i] = __fmul_rn(v, f);
  }
}

// derived index arrays of a batch
__global__ void k_boundary_index(BatchView bv) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= bv.numSentences) return;
  for (u32 b = bv.sentBndOff[s]; b < bv.sentBndOff[s + 1]; ++b) bv.bndSent[b] = s;
}

// Validation of everything a caller supplies that the scoring kernels index with (one thread per
// boundary; reads are guarded, the only writes are status bits).  A sentence that fails is skipped
// by every later kernel -- sent_span() gives it zero boundaries -- so a malformed lattice costs its
// own sentence, not the batch and not the process (reference: one `# ERROR` and on to the next
// sentence, jumanpp.cc:166-171).
struct ValidateArgs {
  u32 totalCodepoints, vocab, hasRnn;
  u8* status;  // [sentences], padded to a multiple of 4 bytes
};

__device__ __forceinline__ void mark_failed(u8* status, u32 sent, u32 bits) {
  atomicOr(reinterpret_cast<unsigned*>(status + (sent & ~3u)), bits << (8u * (sent & 3u)));
}

__global__ void k_validate(BatchView bv, ValidateArgs va) {
  const u32 gB = blockIdx.x * blockDim.x + threadIdx.x;
  if (gB >= bv.totalBoundaries) return;
  const u32 s = bv.bndSent[gB];
  const u32 b0 = bv.sentBndOff[s];
  const u32 nb = bv.sentBndOff[s + 1] - b0;
  const u32 lb = gB - b0;
  u32 bad = 0;
  const u32 n0 = bv.bndNodeOff[gB], n1 = bv.bndNodeOff[gB + 1];
  const u32 e0 = bv.bndEndOff[gB], e1 = bv.bndEndOff[gB + 1];
  const u32 c0 = bv.sentCpOff[s], c1 = bv.sentCpOff[s + 1];
  if (n0 > n1 || n1 > bv.totalNodes || e0 > e1 || e1 > bv.totalEnds || c0 > c1 || c1 > va.totalCodepoints) bad |= JPP_B200_SENT_BAD_OFFSETS;
  if (!bad) {
    const u32 R = n1 - n0, ncp = c1 - c0;
    if (((lb < 2 || lb + 1 == nb) && R != 1) || R > 65535u) bad |= JPP_B200_SENT_BAD_SHAPE;
    for (u32 n = n0; n < n1; ++n) {
      const jpp_b200_node_info ni = bv.nodeInfo[n];
      if (lb >= 2 && lb + 1 != nb && (ni.start > ni.end || ni.end > ncp)) bad |= JPP_B200_SENT_BAD_SPAN;
      i32 id = 0;
      if (bv.nodeEntry) {
      } else if (bv.nodeEntryRow) {
        const u32 row = bv.nodeEntryRow[n];
        if (row >= bv.numEntryRows) bad |= JPP_B200_SENT_BAD_ENTRY;
        else if (bv.rowRnnId) id = bv.rowRnnId[row];
      } else if (lb >= 2) {  // resident form
        const i32 ep = ni.entry_ptr;
        if (ep >= 0) {
          const i32 row = (u32)ep < bv.dicPtrLimit ? bv.dicRowOfPtr[ep] : -1;
          if (row < 0) bad |= JPP_B200_SENT_BAD_ENTRY;
          else if (bv.dicRnnId) id = bv.dicRnnId[row];
        } else if (ep != kEntryBos && ep != kEntryEos) {
          const u32 x0 = bv.sentExtraOff[s], x1 = bv.sentExtraOff[s + 1];
          if (x1 < x0 || x1 > bv.numEntryRows || (u32)~ep >= x1 - x0) bad |= JPP_B200_SENT_BAD_ENTRY;
          else if (bv.rowRnnId) id = bv.rowRnnId[x0 + (u32)~ep];
        }
      }
      if (bv.nodeRnnId) id = bv.nodeRnnId[n];
      if (va.hasRnn && (id < -1 || (id >= 0 && (u32)id >= va.vocab))) bad |= JPP_B200_SENT_BAD_RNN_ID;
    }
    for (u32 e = e0; e < e1; ++e) {
      const jpp_b200_node_ptr p = bv.ends[e];
      if (p.boundary >= lb) { bad |= JPP_B200_SENT_BAD_ENDS; continue; }
      // the left boundary's own offsets are checked by its thread; only compare here
      const u32 l0 = bv.bndNodeOff[b0 + p.boundary], l1 = bv.bndNodeOff[b0 + p.boundary + 1];
      if (l1 < l0 || p.position >= l1 - l0) bad |= JPP_B200_SENT_BAD_ENDS;
    }
  }
  if (bad) mark_failed(va.status, s, bad);
}

__global__ void k_node_index(BatchView bv) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bv.totalBoundaries) return;
  const u32 s = bv.bndSent[b];
  if (bv.sentStatus[s]) return;  // its nodes keep nodeBnd = 0xFFFFFFFF: no kernel touches them
  for (u32 n = bv.bndNodeOff[b]; n < bv.bndNodeOff[b + 1]; ++n) bv.nodeBnd[n] = b;
  const u32 b0 = bv.sentBndOff[s];
  for (u32 e = bv.bndEndOff[b]; e < bv.bndEndOff[b + 1]; ++e) {
    const jpp_b200_node_ptr p = bv.ends[e];
    bv.endNode[e] = bv.bndNodeOff[b0 + p.boundary] + p.position;
  }
}

// Shared-memory hot-weight cache of a CTA (WLoadCached): kCacheSlots 8-byte slots.
constexpr u32 kCacheSlots = 16384;  // 128 KB: the hottest 16k indices are ~3/4 of all gathers
constexpr u32 kCacheThreads = 1024; // one CTA per SM shares one cache

__device__ __forceinline__ WLoadCached make_cache(const float* w, unsigned long long* slots) {
  for (u32 i = threadIdx.x; i < kCacheSlots; i += blockDim.x) slots[i] = 0xFFFFFFFFull;  // empty tag
  __syncthreads();
  return WLoadCached{w, slots, kCacheSlots - 1};
}

// T0: one thread per node.  No hot-weight cache here: measured 13.4 ms without vs 19.9 ms with
// (profiles/README.md r1h) -- most unigram features are high-cardinality (surface, base form).
template <bool LEAN>
__global__ void __launch_bounds__(128) k_t0(const FeatTables* __restrict__ ftg, BatchView bv, const float* __restrict__ w) {
  // per-thread (divergent) table indices here: shared memory, not the constant bank
  __shared__ FeatTables ft;
  extern __shared__ __align__(16) unsigned long long sPrim[];  // [nPrim][128]: the threads' primitive values
  {
    const u32* src = reinterpret_cast<const u32*>(ftg);
    u32* dst = reinterpret_cast<u32*>(&ft);
    for (u32 i = threadIdx.x; i < sizeof(FeatTables) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const WLoadDirect wl{w};
  const u32 node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node < bv.totalNodes) t0_node<WLoadDirect, LEAN>(ft, bv, wl, node, reinterpret_cast<u64*>(sPrim) + threadIdx.x, blockDim.x);
}

// T0 with the compiled-in program (jpp_b200_t0_static.h): no program tables, so none to copy into shared memory;
// the handful of scalars it still reads from FeatTables are uniform (constant bank)
#ifndef JPP_T0_STATIC_MINBLOCKS
#define JPP_T0_STATIC_MINBLOCKS 6   // 80 registers: primitives, the patterns in flight and 32 table reads
#endif
__global__ void __launch_bounds__(128, JPP_T0_STATIC_MINBLOCKS) k_t0_static(const __grid_constant__ FeatTables ft, BatchView bv, const float* __restrict__ w) {
  const WLoadDirect wl{w};
  const u32 node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node < bv.totalNodes) t0_node_static(ft, bv, wl, node);
}

// ---- wavefront walk kernels (see jpp_b200_kernels.h "wavefront walk") ----
// wave index: nodes grouped by their local boundary index
__global__ void k_wave_count(BatchView bv, u32* cnt) {
  const u32 gB = blockIdx.x * blockDim.x + threadIdx.x;
  if (gB >= bv.totalBoundaries) return;
  if (bv.sentStatus[bv.bndSent[gB]]) return;
  const u32 R = bv.bndNodeOff[gB + 1] - bv.bndNodeOff[gB];
  if (R) atomicAdd(cnt + (gB - bv.sentBndOff[bv.bndSent[gB]]), R);
}

__global__ void k_wave_fill(BatchView bv, const u32* off, u32* cursor, u32* waveNodes) {
  const u32 gB = blockIdx.x * blockDim.x + threadIdx.x;
  if (gB >= bv.totalBoundaries) return;
  if (bv.sentStatus[bv.bndSent[gB]]) return;
  const u32 n0 = bv.bndNodeOff[gB];
  const u32 R = bv.bndNodeOff[gB + 1] - n0;
  if (!R) return;
  const u32 lb = gB - bv.sentBndOff[bv.bndSent[gB]];
  const u32 base = off[lb] + atomicAdd(cursor + lb, R);
  for (u32 r = 0; r < R; ++r) waveNodes[base + r] = n0 + r;
}

__global__ void __launch_bounds__(128) k_w_init(BatchView bv, ScoreCfg cfg) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < bv.numSentences) wave_init_sentence(bv, cfg, s);
}

__global__ void __launch_bounds__(128) k_w_gbeam(BatchView bv, ScoreCfg cfg, WaveView wv, u32 nActive, u32 bnd) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) wv.selCount[0] = wv.selCount[1] = wv.selCount[2] = 0;  // consumed by k_w_cutoff_group, a later launch on the same stream
  if (i < nActive) wave_gbeam(bv, cfg, wv, bv.sentOrder[i], bnd);
}

// W1, a group of GW lanes per sentence (the same function as wave_gbeam, which stays the serial
// statement of it for tests/emul): lane l of the group owns left node l (+GW, +2GW, ...), loads
// its B beam elements at once, and G rounds of a group-wide maximum over the packed keys pick
// the global beam -- keys are distinct (score_processor.h:81-115), so the selection has one
// answer whatever the order.  Lane g then resolves element g (left node, its boundary, the node
// before it) and the first-seen numbering of the unique left nodes (dedupT1,
// score_processor.cc:363-385).  GW = 8 (four sentences per warp) when the global beam has at
// most 8 elements -- a boundary has ~7 left nodes, so a full warp per sentence idles -- else 32.
// the group's work for sentence `sent` at local boundary `bnd`; every lane of the group calls it
template <int KB, int GW>
__device__ __forceinline__ void gbeam_group(const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, u32 sent, u32 bnd) {
  const u32 lane = threadIdx.x & 31u, sub = lane % GW;
  const unsigned gshift = lane - sub;
  const unsigned gmask = (GW == 32 ? 0xffffffffu : ((1u << GW) - 1u)) << gshift;
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  if (bnd >= nb || nb <= 3) return;
  const u32 gB = b0 + bnd;
  const u32 R = bv.bndNodeOff[gB + 1] - bv.bndNodeOff[gB];
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam;
  if (R == 0) {  // analyzer_impl.cc:263-266
    if (sub == 0) { wv.gCount[sent] = 0; wv.uCount[sent] = 0; bv.gbCount[gB] = 0; }
    return;
  }
  const u32 e0 = bv.bndEndOff[gB];
  const u32 L = bv.bndEndOff[gB + 1] - e0;
  u64 carried = 0;  // lane g of the group: g-th largest key so far (0 = none; no real key is 0)
  for (u32 c0 = 0; c0 < L; c0 += GW) {
    const u32 left = c0 + sub;
    u64 k[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) k[q] = 0;
    if (left < L) {
      const jpp_b200_beam_element* row = bv.beam + (size_t)bv.endNode[e0 + left] * B;
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        if ((u32)q < B) {
          const jpp_b200_beam_element el = row[q];
          k[q] = is_fake(el) ? 0ull : pack_cand(el.total_score, left, (u32)q);
        }
      }
    }
    u64 next = 0;
    for (u32 g = 0; g < GMAX; ++g) {
      u64 m = carried;
      int mi = -1;
#pragma unroll
      for (int q = 0; q < KB; ++q)
        if (k[q] > m) { m = k[q]; mi = q; }
      const u32 hi = (u32)(m >> 32);
      const u32 whi = __reduce_max_sync(gmask, hi);
      const u32 wlo = __reduce_max_sync(gmask, hi == whi ? (u32)m : 0u);
      const u64 wm = ((u64)whi << 32) | wlo;
      if (wm == 0) break;
      if (m == wm) {
        if (mi < 0) carried = 0;
#pragma unroll
        for (int q = 0; q < KB; ++q)
          if (q == mi) k[q] = 0;
      }
      if (sub == g) next = wm;
    }
    carried = next;
  }
  const u32 G = __popc(__ballot_sync(gmask, carried != 0));
  const bool have = sub < G;
  u32 lnode = 0xFFFFFFFFu, slot = 0, lbnd = 0, t2 = 0;
  if (have) {
    const u32 left = (u32)(carried >> 16) & 0xFFFFu;
    slot = (u32)carried & 0xFFFFu;
    lnode = bv.endNode[e0 + left];
    lbnd = bv.ends[e0 + left].boundary;
    const jpp_b200_beam_element el = bv.beam[(size_t)lnode * B + slot];
    t2 = bv.endNode[bv.bndEndOff[b0 + lbnd] + el.left];
  }
  u32 firstIdx = sub;
  for (u32 j = 0; j < G; ++j) {
    const u32 lj = __shfl_sync(gmask, lnode, j, GW);
    if (have && lj == lnode && j < firstIdx) firstIdx = j;
  }
  const unsigned firstMask = (__ballot_sync(gmask, have && firstIdx == sub) & gmask) >> gshift;
  const u32 u = __popc(firstMask & ((1u << firstIdx) - 1u));
  if (have) {
    const size_t o = (size_t)sent * GMAX;
    wv.gbKey[o + sub] = carried;
    wv.gbLnode[o + sub] = lnode;
    wv.gbT2node[o + sub] = t2;
    wv.gbT1idx[o + sub] = (u8)u;
    if (firstIdx == sub) wv.t1node[o + u] = lnode;
    bv.gbNode[(size_t)gB * GMAX + sub] = lnode;
    bv.gbSlot[(size_t)gB * GMAX + sub] = (u16)slot;
    bv.gbLeftBnd[(size_t)gB * GMAX + sub] = (u16)lbnd;
  }
  if (sub == 0) {
    bv.gbCount[gB] = (u8)G;
    wv.gCount[sent] = (u8)G;
    wv.uCount[sent] = (u8)__popc(firstMask);
  }
}

template <int KB, int GW>
__global__ void __launch_bounds__(128) k_w_gbeam_warp(BatchView bv, ScoreCfg cfg, WaveView wv, u32 nActive, u32 bnd) {
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) wv.selCount[0] = wv.selCount[1] = wv.selCount[2] = 0;
  const u32 w = tid / GW;
  if (w >= nActive) return;
  gbeam_group<KB, GW>(bv, cfg, wv, bv.sentOrder[w], bnd);
}

__global__ void __launch_bounds__(kCacheThreads) k_w_prescore(const __grid_constant__ FeatTables ft, BatchView bv, ScoreCfg cfg,
                                                           WaveView wv, const float* __restrict__ w, u32 first, u32 count) {
  extern __shared__ __align__(16) unsigned long long cacheSlots[];
  const WLoadCached wl = make_cache(w, cacheSlots);
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    wave_prescore(ft, bv, cfg, wv, wl, wv.waveNodes[first + i]);
}

// W3, 8 lanes per sentence, 32 sentences per block (serial statement: wave_cutoff).  A boundary
// has ~7 right nodes: lane r holds the cut score of node r and counts how many scores are larger
// / equal by shuffles; when no run of equal scores straddles the cut that count IS the
// nth_element selection.  When it does (or with more than 8 right nodes) lane 0 replays
// std::nth_element exactly as the serial code.  List positions come from a block-wide prefix
// over the 32 sentences and one atomic per block and list.
// body for a block of 256 threads = 32 sentences: `firstSent` is the block's first index into the
// longest-first sentence order; the three list counters wv.selCount may live in global memory
// (wavefront kernels: one atomic per block and list) or in the block's shared memory (fused walk)
template <u32 SPB>  // sentences per block = blockDim.x / 8
__device__ __forceinline__ void cutoff_group_block(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg,
                                                   const WaveView& wv, u32 nActive, u32 bnd, u32 firstSent) {
  constexpr u32 GW = 8;
  static_assert(SPB <= 32, "the block prefix below is one warp wide");
  __shared__ u32 sSel[SPB], sBi[SPB], sTri[SPB], sBase[3];
  __shared__ unsigned long long sCnt[3];
  const u32 lane = threadIdx.x & 31u, sub = lane % GW, grp = threadIdx.x / GW;
  const unsigned gshift = lane - sub;
  const unsigned gmask = ((1u << GW) - 1u) << gshift;
  const u32 sentIdx = firstSent + grp;
  if (threadIdx.x < 3) sCnt[threadIdx.x] = 0;
  __syncthreads();
  bool active = sentIdx < nActive;
  u32 sent = 0, gB = 0, n0 = 0, R = 0, G = 0, U = 0, L = 0;
  if (active) {
    sent = bv.sentOrder[sentIdx];
    u32 b0;
    const u32 nb = sent_span(bv, sent, b0);
    if (bnd >= nb || nb <= 3) {
      active = false;
    } else {
      gB = b0 + bnd;
      n0 = bv.bndNodeOff[gB];
      R = bv.bndNodeOff[gB + 1] - n0;
      if (R == 0) active = false;
    }
  }
  StepSizes z = step_sizes(cfg, R, 0);
  if (active) {
    G = wv.gCount[sent];
    U = wv.uCount[sent];
    L = bv.bndEndOff[gB + 1] - bv.bndEndOff[gB];
    z = step_sizes(cfg, R, G);
  }
  // Selection as a bit per right node (boundaries of up to 64 nodes; beyond that lane 0 marks
  // selPos in global memory as the serial code does).
#ifndef JPP_CUTOFF_SM_R
#define JPP_CUTOFF_SM_R 16
#endif
  // boundaries of up to SM_R right nodes select in shared memory (a boundary has ~7).  16, not 64: 5 KB instead of
  // 17 KB per block leaves room on an SM for the tensor-core level kernel of ANOTHER window next to four walk
  // blocks (end-to-end 72.6 -> 69.3 ms per 100 000 sentences with six windows side by side, r2l)
  constexpr u32 SM_R = JPP_CUTOFF_SM_R;
  static_assert(SM_R >= 8 && SM_R <= 64, "the selection is a 64-bit mask; rank counting uses 8 slots");
  __shared__ float sCut[SPB][SM_R];
  __shared__ u32 sIdx[SPB][SM_R];
  const bool bits = active && R <= SM_R;
  unsigned long long selBits = 0;
  u32 nsel = 0;
  if (active) {
    if (!(cfg.rcheck > 0 && R > z.toKeep)) {
      if (bits) selBits = ~0ull;
      else for (u32 r = sub; r < R; r += GW) bv.selPos[n0 + r] = 0;
      nsel = R;
    } else {
      nsel = z.toKeep;
      bool exact = true;  // std::nth_element has to be replayed
      if (R <= GW) {
        // rank count: when no run of equal scores straddles the cut it IS the selection
        const bool valid = sub < R;
        float v = 0;
        if (valid) for (u32 j = 0; j < z.full; ++j) v += bv.pre[(size_t)j * bv.totalNodes + n0 + sub];
        u32 greater = 0, eq = 0;
        for (u32 j = 0; j < R; ++j) {
          const float vj = __shfl_sync(gmask, v, j, GW);
          if (vj > v) ++greater;
          else if (!(v > vj)) ++eq;  // counts the lane itself
        }
        const bool sure = valid && greater + eq <= z.toKeep;
        const bool out = valid && greater >= z.toKeep;
        const unsigned amb = __ballot_sync(gmask, valid && !sure && !out) & gmask;
        if (amb == 0) {
          exact = false;
          selBits = (unsigned long long)((__ballot_sync(gmask, sure) & gmask) >> gshift);
        } else {
          if (valid) { sCut[grp][sub] = v; sIdx[grp][sub] = sub; }
        }
      } else if (bits) {
        for (u32 r = sub; r < R; r += GW) {
          float c = 0;
          for (u32 j = 0; j < z.full; ++j) c += bv.pre[(size_t)j * bv.totalNodes + n0 + r];
          sCut[grp][r] = c;
          sIdx[grp][r] = r;
        }
      }
      if (exact && bits) {
        __syncwarp(gmask);
        u32 lo = 0, hi = 0;
        if (sub == 0) {
          em_std_nth_element(sIdx[grp], (long)z.toKeep, (long)R, ScoreCmp{sCut[grp]});
          for (u32 k = 0; k < z.toKeep; ++k) {
            const u32 r = sIdx[grp][k];
            if (r < 32) lo |= 1u << r; else hi |= 1u << (r - 32);
          }
        }
        lo = __shfl_sync(gmask, lo, 0, GW);
        hi = __shfl_sync(gmask, hi, 0, GW);
        selBits = ((unsigned long long)hi << 32) | lo;
      } else if (exact) {
        if (sub == 0) {
          u32* idx = bv.cutIdx + n0;
          float* cut = bv.cut + n0;
          for (u32 r = 0; r < R; ++r) {
            float c = 0;
            for (u32 j = 0; j < z.full; ++j) c += bv.pre[(size_t)j * bv.totalNodes + n0 + r];
            cut[r] = c;
            idx[r] = r;
            bv.selPos[n0 + r] = NOT_SELECTED;
          }
          em_std_nth_element(idx, (long)z.toKeep, (long)R, ScoreCmp{cut});
          for (u32 k = 0; k < z.toKeep; ++k) bv.selPos[n0 + idx[k]] = 0;
        }
        __syncwarp(gmask);
      }
    }
  }
  const bool scoreTail = active && z.remaining > 0 && G > 0;
  const u32 u0 = first_bigram_item(cfg);
  const u32 biPer = scoreTail && U > u0 ? U - u0 : 0u;
  const u32 triPer = scoreTail ? z.remaining : 0u;
  if (sub == 0) {
    sSel[grp] = nsel;
    sBi[grp] = nsel * biPer;
    sTri[grp] = nsel * triPer;
    if (active && bv.counters) {
      unsigned long long g = 0, sb = 0, gi = 0;
      step_counters(ft, cfg, z, R, L, G, U, nsel, g, sb, gi);
      atomicAdd(&sCnt[0], g);
      atomicAdd(&sCnt[1], sb);
      atomicAdd(&sCnt[2], gi);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    u32 v[3] = {0, 0, 0};
    if (lane < SPB) { v[0] = sSel[lane]; v[1] = sBi[lane]; v[2] = sTri[lane]; }
    u32 inc[3] = {v[0], v[1], v[2]};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc[q], d);
        if ((int)lane >= d) inc[q] += t;
      }
    }
    if (lane == 31) {
      sBase[0] = inc[0] ? atomicAdd(wv.selCount + 0, inc[0]) : 0u;
      sBase[1] = inc[1] ? atomicAdd(wv.selCount + 1, inc[1]) : 0u;
      sBase[2] = inc[2] ? atomicAdd(wv.selCount + 2, inc[2]) : 0u;
    }
    if (lane < SPB) {
      sSel[lane] = inc[0] - v[0];
      sBi[lane] = inc[1] - v[1];
      sTri[lane] = inc[2] - v[2];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && bv.counters && (sCnt[0] | sCnt[1])) {
    atomicAdd(bv.counters + 0, sCnt[0]);
    atomicAdd(bv.counters + 1, sCnt[1]);
    atomicAdd(bv.counters + 2, sCnt[2]);
  }
  if (!active) return;
  const u32 base = sBase[0] + sSel[grp], biBase = sBase[1] + sBi[grp], triBase = sBase[2] + sTri[grp];
  u32 run = 0;
  for (u32 r0 = 0; r0 < R; r0 += GW) {
    const u32 r = r0 + sub;
    const bool valid = r < R;
    const bool sel = valid && (bits ? ((selBits >> r) & 1ull) != 0 : bv.selPos[n0 + r] != NOT_SELECTED);
    const unsigned m = (__ballot_sync(gmask, sel) & gmask) >> gshift;
    if (sel) {
      const u32 k = run + __popc(m & ((1u << sub) - 1u));
      const u32 pos = base + k;
      bv.selPos[n0 + r] = pos;
      if (scoreTail)
        emit_items(bv, cfg, wv, sent, z, G, U, n0 + r, pos, wv.biItems + biBase + (size_t)k * biPer,
                   wv.triItems + triBase + (size_t)k * triPer);
    } else if (valid) {
      bv.selPos[n0 + r] = NOT_SELECTED;
    }
    run += __popc(m);
  }
}

__global__ void __launch_bounds__(256) k_w_cutoff_group(const __grid_constant__ FeatTables ft, BatchView bv, ScoreCfg cfg,
                                                       WaveView wv, u32 nActive, u32 bnd) {
  cutoff_group_block<32>(ft, bv, cfg, wv, nActive, bnd, blockIdx.x * 32u);
}

// W4: bigram items first, trigram items after them; one 16-byte item per thread
__global__ void __launch_bounds__(kCacheThreads) k_w_items(const __grid_constant__ FeatTables ft, BatchView bv, WaveView wv,
                                                        const float* __restrict__ w) {
  extern __shared__ __align__(16) unsigned long long cacheSlots[];
  const WLoadCached wl = make_cache(w, cacheSlots);
  const u32 n1 = wv.selCount[1], n2 = wv.selCount[2];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) wave_item_bigram(ft, bv, wv, wl, wv.biItems[i]);
    else wave_item_trigram(ft, bv, wv, wl, wv.triItems[i - n1]);
  }
}

// W5: thread per right node of the wave
template <bool SMALL>
__global__ void __launch_bounds__(128) k_w_final(BatchView bv, ScoreCfg cfg, WaveView wv, u32 first, u32 count) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) wave_final_node<SMALL>(bv, cfg, wv, wv.waveNodes[first + i]);
}

// ------------------------------------------------------------ fused walk -----
// (the kernel itself: jpp_b200_walk.cuh)

// block-wide maximum through a 4-word shared scratch (all threads call it)
__device__ __forceinline__ u32 block_max_u32(u32 v, u32* scratch) {
  if (threadIdx.x == 0) scratch[3] = 0;
  __syncthreads();
  v = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31u) == 0) atomicMax(scratch + 3, v);
  __syncthreads();
  const u32 r = scratch[3];
  __syncthreads();
  return r;
}

inline bool fused_walk_ok(const ScoreCfg& cfg) {
  static const bool off = getenv("JPP_B200_WAVEFRONT") != nullptr;  // A/B and per-kernel profiling
  return !off && cfg.rcheck > 0 && cfg.gbeam > 0;  // any beam / global beam the library accepts (<= 32)
}

}  // namespace
#include "jpp_b200_walk.cuh"
namespace {

// full-beam mode: the whole boundary step of a right node in one thread
__global__ void __launch_bounds__(128) k_w_full(const __grid_constant__ FeatTables ft, BatchView bv, ScoreCfg cfg, WaveView wv,
                                                const float* __restrict__ w, u32 first, u32 count) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const WLoadDirect wl{w};
  u32 g = 0;
  if (i < count) g = wave_full_node(ft, bv, cfg, wl, wv.waveNodes[first + i]);
  if (bv.counters) {
    g = __reduce_add_sync(0xffffffffu, g);
    if ((threadIdx.x & 31) == 0 && g) atomicAdd(bv.counters + 0, (unsigned long long)g);
  }
}

__global__ void k_rebase(u32* p, u32 n, u32 base) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] -= base;
}

__global__ void k_fill_i32(i32* p, size_t n, i32 v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Thread = sentence, in input order.  Taking the sentences in the length-sorted order instead (a warp's 32 sentences
// then loop about equally often) was measured and is slower, 7.3 -> 7.8 ms (r2n): neighbours in input order are
// neighbours in memory, and this kernel is all dependent loads.
__global__ void k_rnn_tree(BatchView bv, ScoreCfg cfg, RnnWork rw) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < bv.numSentences) rnn_build_tree(bv, cfg, rw, s);
}

// level lists (the histogram is accumulated by k_rnn_tree); 8 lanes share a sentence's nodes
struct __align__(16) LevelRec {  // = TcRec of jpp_b200_rnn_tc.cuh
  u32 inRow, outRow, embRow, rnode;
};
__global__ void k_rnn_level_fill(BatchView bv, RnnWork rw, u32* cursor, const u32* __restrict__ ctxOff, LevelRec* levelRec) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 s = t >> 3, sub = t & 7u;
  if (s >= bv.numSentences) return;
  const u32 base0 = rw.rnOff[s];
  const u32 n = rw.rnCount[s];
  u32 b0unused;
  const u32 nb = sent_span(bv, s, b0unused);
  for (u32 i = 2 + sub; i < n; i += 8) {
    const u32 q = base0 + i;
    if ((u32)rw.bnd[q] + 1 < nb) {
      const u32 d = (u32)rw.depth[q];
      const u32 pos = atomicAdd(cursor + d, 1u);
      rw.levelList[rw.levelOff[d] + pos] = q;
      if (levelRec) {  // the tensor-core level kernel reads its rows from here (prev is a node of the same sentence)
        const u32 c0 = ctxOff[s];
        i32 id = rw.id[q];
        if (id == -1) id = 0;  // rnn_scorer_gbeam.cc:126-129
        LevelRec r;
        r.inRow = c0 + ((u32)rw.prev[q] - base0);
        r.outRow = c0 + i;
        r.embRow = (u32)id;
        r.rnode = q;
        levelRec[rw.levelOff[d] + pos] = r;
      }
    }
  }
}

// context row of an rnode: sentence-compact
__device__ __forceinline__ size_t ctx_row(const RnnWork& rw, const u32* ctxOff, const u32* rnSent, u32 q) {
  const u32 s = rnSent[q];
  return (size_t)ctxOff[s] + (q - rw.rnOff[s]);
}

__global__ void k_rnn_mark_sentence(BatchView bv, RnnWork rw, u32* rnSent) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= bv.numSentences) return;
  const u32 base0 = rw.rnOff[s];
  const u32 n = rw.rnCount[s];
  const u32 m = rw.rsCount[s] > n ? rw.rsCount[s] : n;
  for (u32 i = 0; i < m; ++i) rnSent[base0 + i] = s;
}

// BOS context of every sentence: sigmoid(W^T 0 + emb[0]) (rnn_scorer_gbeam.cc:37-46);
// dot32 of a zero vector is +0, so the state is sigmoid(0 + emb[0][i])
__global__ void k_rnn_bos(BatchView bv, RnnWork rw, RnnTables rt, const u32* ctxOff) {
  const u32 s = blockIdx.x;
  if (rw.rnCount[s] < 2) return;
  float* out = rw.ctx + ((size_t)ctxOff[s] + 1) * rt.H;
  for (u32 i = threadIdx.x; i < rt.H; i += blockDim.x) {
    float v = __fadd_rn(0.0f, rt.emb[i]);
    out[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
  }
  (void)bv;
}

// Level kernel: ctx' = sigmoid(W^T ctx + emb[id]) (mikolov_rnn_impl.h:202-214) for
// the tree nodes of one depth.  A warp takes a tile of NT nodes; lane l owns
// output i = i0 + l for all NT nodes, so every load of the (transposed)
// recurrent matrix is reused NT times from a register.
// Dot-product order = oracle/eigen_shim: 32 residue partials (k = r, r+32, ...)
// folded by the 16/8/4/2/1 tree.  The residues are visited in bit-reversed
// order, which is the depth-first order of that tree, so a partial is folded
// into its parent as soon as its sibling exists and only <= 6 values per
// (output, node) are live instead of 32 -- same additions, same operands, same
// order as the serial definition.
template <int NT>
__global__ void __launch_bounds__(128) k_rnn_level(RnnWork rw, RnnTables rt, const float* __restrict__ Wt,
                                                 const u32* ctxOff, const u32* rnSent, u32 level) {
  extern __shared__ __align__(16) float sctx[];
  const u32 cnt = rw.levelCount[level];
  const u32 off = rw.levelOff[level];
  const int lane = threadIdx.x & 31;
  const u32 warp = threadIdx.x >> 5;
  const u32 warpsPerBlock = blockDim.x >> 5;
  const u32 H = rt.H;
  float* cS = sctx + (size_t)warp * H * NT;  // [k][NT]
  const u32 tiles = (cnt + NT - 1) / NT;
  for (u32 tile = blockIdx.x * warpsPerBlock + warp; tile < tiles; tile += gridDim.x * warpsPerBlock) {
    const u32 first = tile * NT;
    const u32 nvalid = (cnt - first) < (u32)NT ? (cnt - first) : (u32)NT;
    u32 qn[NT];
    const float* emb[NT];
    float* cout[NT];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const u32 q = rw.levelList[off + first + ((u32)j < nvalid ? (u32)j : 0u)];
      qn[j] = q;
      const u32 p = (u32)rw.prev[q];
      const float* cin = rw.ctx + ctx_row(rw, ctxOff, rnSent, p) * H;
      cout[j] = rw.ctx + ctx_row(rw, ctxOff, rnSent, q) * H;
      i32 id = rw.id[q];
      if (id == -1) id = 0;  // rnn_scorer_gbeam.cc:126-129
      emb[j] = rt.emb + (size_t)id * H;
      for (u32 k = (u32)lane; k < H; k += 32) cS[(size_t)k * NT + j] = cin[k];
    }
    __syncwarp();
    for (u32 i0 = 0; i0 < H; i0 += 32) {
      const u32 i = i0 + (u32)lane;
      const u32 ic = i < H ? i : H - 1;
      float st[6][NT];
      int sp = 0;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        // residue of the r-th leaf in depth-first order = 5-bit reversal of r
        const int l = ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
        float part[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) part[j] = 0.0f;
        for (u32 k = (u32)l; k < H; k += 32) {
          const float wv = Wt[(size_t)k * H + ic];
          const float4* c4 = reinterpret_cast<const float4*>(cS + (size_t)k * NT);
#pragma unroll
          for (int j4 = 0; j4 < NT / 4; ++j4) {
            const float4 c = c4[j4];
            part[j4 * 4 + 0] = __fmaf_rn(wv, c.x, part[j4 * 4 + 0]);
            part[j4 * 4 + 1] = __fmaf_rn(wv, c.y, part[j4 * 4 + 1]);
            part[j4 * 4 + 2] = __fmaf_rn(wv, c.z, part[j4 * 4 + 2]);
            part[j4 * 4 + 3] = __fmaf_rn(wv, c.w, part[j4 * 4 + 3]);
          }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) st[sp][j] = part[j];
        ++sp;
        // binary-counter folding: after leaf r, fold once per trailing one bit of r
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          if (((r >> b) & 1) && ((r & ((1 << b) - 1)) == ((1 << b) - 1))) {
#pragma unroll
            for (int j = 0; j < NT; ++j) st[sp - 2][j] = __fadd_rn(st[sp - 2][j], st[sp - 1][j]);
            --sp;
          }
        }
      }
      if (i < H) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if ((u32)j < nvalid) {
            const float v = __fadd_rn(st[0][j], emb[j][i]);
            cout[j][i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
          }
        }
      }
    }
    (void)qn;
  }
}

// how many blocks share the outputs of one tile of tree nodes: the block's rows of the recurrent
// matrix (H / split rows of H floats) must fit shared memory; H / split stays a multiple of 4
__host__ __device__ inline u32 rnn_level_split(u32 H) {
  u32 split = 1;
  while ((size_t)H * H * 4 / split > 160 * 1024 && (H / (split * 2)) % 4 == 0) split *= 2;
  return split;
}

// Level kernel, layer sizes that are a multiple of 32 (KI = H / 32 <= 8): the same sums as
// k_rnn_level, arranged so that the additions of the 16/8/4/2/1 tree happen ACROSS lanes.
// Lane l owns residue l of every dot product: p_l = fma chain over k = l, l+32, ... (KI terms).
// A warp takes 8 tree nodes (their previous contexts sit in registers, lane l holding elements
// l + 32 i) and walks the outputs four at a time: 32 (node, output) pairs, 32 partials per lane,
// 4*KI*8 FMAs per lane against 4*KI shared-memory reads of the recurrent matrix.  The 32 x 32
// partials are then folded by a transpose-reduce: at lane distance 16 every lane keeps half of
// its pairs and receives the matching partials of its partner (p_l + p_{l+16}), then distance 8,
// 4, 2, 1 -- 31 shuffles for 32 complete dot products, and lane L ends up with pair L.  Operands
// and association are exactly those of oracle/eigen_shim dot32 (float + is commutative).
template <int KI>
__global__ void __launch_bounds__(256) k_rnn_level_x(RnnWork rw, RnnTables rt, const u32* ctxOff, const u32* rnSent, u32 level) {
  constexpr int TN = 8, TO = 4;
  extern __shared__ __align__(16) float Ws[];  // this block's rows of the recurrent matrix as stored: Ws[o * H + k]
  const u32 cnt = rw.levelCount[level];
  const u32 tiles = (cnt + TN - 1) / TN;
  const u32 warpsPerBlock = blockDim.x >> 5;
  if (blockIdx.x * warpsPerBlock >= tiles) return;
  constexpr u32 H = 32u * KI;
  // large layers: the outputs are split over blockIdx.y so that a block's rows fit shared memory
  const u32 oCount = H / gridDim.y, oBase = blockIdx.y * oCount;
  {
    const float4* src = reinterpret_cast<const float4*>(rt.matrix + (size_t)oBase * H);
    float4* dst = reinterpret_cast<float4*>(Ws);
    for (u32 i = threadIdx.x; i < oCount * H / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const u32 off = rw.levelOff[level];
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  for (u32 tile = blockIdx.x * warpsPerBlock + warp; tile < tiles; tile += gridDim.x * warpsPerBlock) {
    const u32 first = tile * TN;
    const u32 nvalid = (cnt - first) < (u32)TN ? (cnt - first) : (u32)TN;
    // lane j < 8 resolves node j; everyone then reads the 8 previous contexts
    u32 inRow = 0, outRow = 0, embRow = 0;
    {
      const u32 j = lane & 7u;
      const u32 q = rw.levelList[off + first + (j < nvalid ? j : 0u)];
      inRow = (u32)ctx_row(rw, ctxOff, rnSent, (u32)rw.prev[q]);
      outRow = (u32)ctx_row(rw, ctxOff, rnSent, q);
      i32 id = rw.id[q];
      if (id == -1) id = 0;  // rnn_scorer_gbeam.cc:126-129
      embRow = (u32)id;
    }
    // contexts of node pairs side by side: one packed FMA (fma.rn.f32x2, two IEEE single-precision
    // FMAs in one issue slot) advances the same residue of two nodes
    float2 c2[TN / 2][KI];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const u32 r = __shfl_sync(full, inRow, j);
      const float* cin = rw.ctx + (size_t)r * H;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        const float v = cin[lane + 32u * i];
        if (j & 1) c2[j / 2][i].y = v; else c2[j / 2][i].x = v;
      }
    }
    // lane L finishes pair L = j * TO + t: node j = L / 4, output o0 + (L % 4)
    const u32 myJ = lane >> 2, myT = lane & 3u;
    const float* myEmb = rt.emb + (size_t)__shfl_sync(full, embRow, myJ) * H;
    float* myOut = rw.ctx + (size_t)__shfl_sync(full, outRow, myJ) * H;
    for (u32 o0 = 0; o0 < oCount; o0 += TO) {
      float2 P2[TN / 2][TO];
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int t = 0; t < TO; ++t) P2[jj][t] = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int t = 0; t < TO; ++t) {
#pragma unroll
        for (int i = 0; i < KI; ++i) {
          const float w = Ws[(o0 + t) * H + 32u * i + lane];
          const float2 ww = make_float2(w, w);
#pragma unroll
          for (int jj = 0; jj < TN / 2; ++jj) P2[jj][t] = __ffma2_rn(ww, c2[jj][i], P2[jj][t]);
        }
      }
      float P[TN * TO];
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int t = 0; t < TO; ++t) {
          P[(2 * jj) * TO + t] = P2[jj][t].x;
          P[(2 * jj + 1) * TO + t] = P2[jj][t].y;
        }
      float v16[16], v8[8], v4[4], v2[2], v1;
      {
        const bool up = (lane & 16u) != 0;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const float a = P[x], b = P[x + 16];
          v16[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 16));
        }
      }
      {
        const bool up = (lane & 8u) != 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const float a = v16[x], b = v16[x + 8];
          v8[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 8));
        }
      }
      {
        const bool up = (lane & 4u) != 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float a = v8[x], b = v8[x + 4];
          v4[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 4));
        }
      }
      {
        const bool up = (lane & 2u) != 0;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const float a = v4[x], b = v4[x + 2];
          v2[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 2));
        }
      }
      {
        const bool up = (lane & 1u) != 0;
        const float a = v2[0], b = v2[1];
        v1 = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 1));
      }
      const u32 o = oBase + o0 + myT;
      const float v = __fadd_rn(v1, myEmb[o]);
      const float r = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
      if (myJ < nvalid) myOut[o] = r;
    }
  }
}

// One thread per score entry of the batch-wide flat list: nce[id] . ctx(prev) in the
// eigen_shim order (32 residue partials in registers, then the 16/8/4/2/1 fold),
// maxent hash, nce constant (mikolov_rnn_impl.h:216-257).
__global__ void __launch_bounds__(128) k_rnn_scores(BatchView bv, ScoreCfg cfg, RnnWork rw, RnnTables rt, const u32* ctxOff,
                                                  const u32* rnSent) {
  const u32 H = rt.H;
  const u32 total = *rw.scoreTotal;
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const u32 si = rw.scoreList[e];
    const u32 s = rnSent[si];
    const i32 q = rw.rsRnode[si];
    const i32 id = rw.id[q];
    float sc;
    if (id == rt.unkId) {
      sc = __fadd_rn(rt.unkConstant, __fmul_rn(rt.unkLengthPenalty, (float)rw.length[q]));  // rnn_scorer_gbeam.cc:236-240
    } else {
      const i32 p = rw.prev[q];
      const i32 eid = id == -1 ? 0 : id;
      const float* a = rt.nce + (size_t)eid * H;
      const float* b = rw.ctx + ctx_row(rw, ctxOff, rnSent, (u32)p) * H;
      float part[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j] = 0.0f;
      u32 k0 = 0;
      if ((H & 3u) == 0) {
        for (; k0 + 32 <= H; k0 += 32) {
          const float4* a4 = reinterpret_cast<const float4*>(a + k0);
          const float4* b4 = reinterpret_cast<const float4*>(b + k0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 x = a4[j4], y = b4[j4];
            part[j4 * 4 + 0] = __fmaf_rn(x.x, y.x, part[j4 * 4 + 0]);
            part[j4 * 4 + 1] = __fmaf_rn(x.y, y.y, part[j4 * 4 + 1]);
            part[j4 * 4 + 2] = __fmaf_rn(x.z, y.z, part[j4 * 4 + 2]);
            part[j4 * 4 + 3] = __fmaf_rn(x.w, y.w, part[j4 * 4 + 3]);
          }
        }
      }
      for (; k0 < H; k0 += 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + j < H) part[j] = __fmaf_rn(a[k0 + j], b[k0 + j], part[j]);
      }
#pragma unroll
      for (int stride = 16; stride >= 1; stride >>= 1) {
#pragma unroll
        for (int l = 0; l < stride; ++l) part[l] = __fadd_rn(part[l], part[l + stride]);
      }
      sc = rnn_finish_score(rt, part[0], rw.id[p], id);
    }
    const u16 g = rw.rsEosG[si];
    if (g != 0xFFFFu) bv.eosRnn[(size_t)s * cfg.gbeam + g] = sc;
    else bv.beamScores[((size_t)rw.rsNode[si] * cfg.beam + rw.rsSlot[si]) * cfg.S + 1] = sc;
  }
}

// The same scores, a warp per 32 entries, for layer sizes that are a multiple of 32 (KI = H / 32):
// lane l reads elements l + 32 i of every entry's two rows -- coalesced, one line per load
// instead of 32 -- and keeps residue l of the 32 dot products; the transpose-reduce of
// k_rnn_level_x then hands entry L's complete sum to lane L, which finishes the score.
template <int KI>
__global__ void __launch_bounds__(128) k_rnn_scores_x(BatchView bv, ScoreCfg cfg, RnnWork rw, RnnTables rt, const u32* ctxOff,
                                                    const u32* rnSent) {
  constexpr u32 H = 32u * KI;
  const u32 total = *rw.scoreTotal;
  const u32 lane = threadIdx.x & 31u;
  const u32 warpsPerBlock = blockDim.x >> 5;
  const unsigned full = 0xffffffffu;
  const u32 tiles = (total + 31) / 32;
  for (u32 tile = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); tile < tiles; tile += gridDim.x * warpsPerBlock) {
    const u32 e = tile * 32 + lane;
    const bool valid = e < total;
    u32 si = 0, s = 0, aRow = 0, bRow = 0;
    i32 q = 0, id = 0, p = 0;
    bool dot = false;
    if (valid) {
      si = rw.scoreList[e];
      s = rnSent[si];
      q = rw.rsRnode[si];
      id = rw.id[q];
      if (id != rt.unkId) {
        p = rw.prev[q];
        aRow = (u32)(id == -1 ? 0 : id);
        bRow = (u32)ctx_row(rw, ctxOff, rnSent, (u32)p);
        dot = true;
      }
    }
    float P[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) {
      P[x] = 0.0f;
      const bool need = __shfl_sync(full, dot ? 1 : 0, x) != 0;
      const u32 ar = __shfl_sync(full, aRow, x), br = __shfl_sync(full, bRow, x);
      if (need) {
        const float* a = rt.nce + (size_t)ar * H;
        const float* b = rw.ctx + (size_t)br * H;
#pragma unroll
        for (int i = 0; i < KI; ++i) P[x] = __fmaf_rn(a[lane + 32u * i], b[lane + 32u * i], P[x]);
      }
    }
    float v16[16], v8[8], v4[4], v2[2], v1;
    {
      const bool up = (lane & 16u) != 0;
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float a = P[x], b = P[x + 16];
        v16[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 16));
      }
    }
    {
      const bool up = (lane & 8u) != 0;
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const float a = v16[x], b = v16[x + 8];
        v8[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 8));
      }
    }
    {
      const bool up = (lane & 4u) != 0;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float a = v8[x], b = v8[x + 4];
        v4[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 4));
      }
    }
    {
      const bool up = (lane & 2u) != 0;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const float a = v4[x], b = v4[x + 2];
        v2[x] = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 2));
      }
    }
    {
      const bool up = (lane & 1u) != 0;
      const float a = v2[0], b = v2[1];
      v1 = __fadd_rn(up ? b : a, __shfl_xor_sync(full, up ? a : b, 1));
    }
    if (!valid) continue;
    float sc;
    if (!dot) sc = __fadd_rn(rt.unkConstant, __fmul_rn(rt.unkLengthPenalty, (float)rw.length[q]));  // rnn_scorer_gbeam.cc:236-240
    else sc = rnn_finish_score(rt, v1, rw.id[p], id);
    const u16 g = rw.rsEosG[si];
    if (g != 0xFFFFu) bv.eosRnn[(size_t)s * cfg.gbeam + g] = sc;
    else bv.beamScores[((size_t)rw.rsNode[si] * cfg.beam + rw.rsSlot[si]) * cfg.S + 1] = sc;
  }
}

__global__ void k_best_path(BatchView bv, ScoreCfg cfg, u32* pathLen, u32* pathNodes, u32 nodeBase) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < bv.numSentences) best_path_sentence(bv, cfg, pathLen, pathNodes, s, nodeBase);
}

// adjustBeamScores + remakeEosBeam (analyzer_impl.cc:286-294; serial statement: rerank_sentence).
// Totals of a boundary's global-beam elements depend on totals of EARLIER boundaries only, so
// the 8 lanes of a group take the elements of one boundary side by side and the group walks
// the boundaries in order; lane 0 then rebuilds the EOS beam.
__global__ void __launch_bounds__(256) k_rerank(BatchView bv, ScoreCfg cfg, float w0, float w1) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 sent = t >> 3, sub = threadIdx.x & 7u;
  if (sent >= bv.numSentences) return;
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31u) - sub);
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  if (nb <= 3) return;
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam, S = cfg.S;
  for (u32 bnd = 3; bnd < nb; ++bnd) {
    const u32 gB = b0 + bnd;
    const u32 G = (bv.bndNodeOff[gB + 1] == bv.bndNodeOff[gB]) ? 0u : bv.gbCount[gB];
    for (u32 i = sub; i < G; i += 8) {
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
    __syncwarp(gmask);  // the next boundary reads these totals
  }
  if (sub == 0) rerank_eos(bv, cfg, w0, w1, sent);
}

}  // namespace
#include "jpp_b200_rnn_tc.cuh"
namespace {

// ------------------------------------------------------------- helpers ------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

}  // namespace

namespace { struct WaveProfile; }

// ---------------------------------------------------------------- gather probe (diagnostic) ----
// What this GPU can do for the access pattern of the scored path and nothing else: every thread issues
// `perThread` independent 4-byte reads at hash-random indices of the handle's own perceptron table, ILP at a
// time, the values summed.  No lattice, no ordering, no barriers: the rate it reaches is the ceiling any
// kernel that gathers from this table can have ("HBM-gather roofline"), bench.py quotes the walk against it.
template <int ILP>
__global__ void __launch_bounds__(256) k_gather_probe(const float* __restrict__ w, u32 mask, u64 seed, u32 perThread, float* sink) {
  u64 x = (seed ^ ((u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull)) | 1ull;
  float acc = 0.f;
  for (u32 i = 0; i < perThread; i += ILP) {
    float v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      x ^= x >> 12; x ^= x << 25; x ^= x >> 27;   // xorshift64*
      const u32 idx = (u32)((x * 0x2545F4914F6CDD1Dull) >> 32) & mask;
      v[j] = __ldg(w + idx);
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += v[j];
  }
  if (acc == 1234567.0f) *sink = acc;  // keeps the loads alive
}

struct jpp_b200_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t upStream = nullptr;  // uploads + index kernels: highest priority (see upload_window)
  std::string error;
  HostModel hm;
  ScoreCfg cfg{};
  jpp_b200_config pubCfg{};
  FeatTables* dFt = nullptr;
  float* dWeights = nullptr;
  RnnTables rt{};
  bool hasRnn = false;
  // resident entry table (jpp_b200_add_entry_rows); lanes share the owner's device arrays
  struct EntryTable {
    i32* dRowOfPtr = nullptr; u32 ptrLimit = 0, ptrCap = 0;
    i32* dRows = nullptr; i32* dRnnId = nullptr; u32 numRows = 0, rowCap = 0 /* i32 elements */, idCap = 0;
  } dic;
  std::vector<i32> hostRowOfPtr;  // owner only: host copy of the direct map
  cudaStream_t tableStream = nullptr;  // owner only: uploads of table rows (never waits for a batch in flight)
  bool staticProgram = false;  // the uploaded n-gram program is the compiled-in one (jpp_b200_static_jumandic.h)
  bool staticT0 = false;       // ... and so is its per-node part (jpp_b200_static_jumandic_t0.h)
  bool rnnTc = true;           // layer size 128: tcgen05 level kernel (JPP_B200_RNN_TC=0 at create: the FFMA kernel)
  float* dRnnBlocks[4] = {nullptr, nullptr, nullptr, nullptr};
  float* dRnnWt = nullptr;  // transposed recurrent matrix
  float* dRnnWImage = nullptr;  // layer size 128: the two TF32 planes of W in the tensor-core operand layout (jpp_b200_rnn_tc.cuh)
  // window of the caller's batch that is resident (upload_window): offsets of its first sentence /
  // boundary / node; results go to the matching slices of the caller's arrays
  u32 windowSentBase = 0, windowBndBase = 0, windowNodeBase = 0;
  // second lane of jpp_b200_score_batch: own stream and buffers, the model memory of this handle
  jpp_b200_handle* lane = nullptr;   // next lane (a chain: at most kMaxLanes - 1 of them)
  std::mutex* computeToken = nullptr;  // lanes of one handle take turns on the GPU's compute (see jpp_b200_score_batch)
  bool ownsModel = true;
  WaveProfile* prof = nullptr;
  float scoreW[2] = {1.0f, 0.0f};
  int smCount = 148;
  // batch buffers
  enum { B_SBO, B_SCO, B_BNO, B_BEO, B_INFO, B_ENTRY, B_ENTRYROW, B_PH, B_RNNID, B_ENDS, B_CPS, B_NODEBND, B_BNDSENT, B_ENDNODE,
         B_ORDER, B_BEAM, B_BSCORES, B_PATS, B_T0, B_PRE, B_CUT, B_CUTIDX, B_SELPOS, B_BI0, B_GBNODE, B_GBSLOT, B_GBLBND, B_GBCNT,
         B_EOSG, B_EOSLOC, B_EOSRNN, B_COUNTERS, B_WORK,
         B_PATHLEN, B_PATHNODES, B_STATUS, B_EXTRAOFF,
         W_GBKEY, W_GBLNODE, W_GBT2, W_GBT1IDX, W_T1NODE, W_GCNT, W_UCNT, W_SELCNT, W_BIITEMS, W_TRIITEMS, W_SCB, W_TRI, W_NODES, W_OFF, W_CURSOR,
         R_OFF, R_ID, R_PREV, R_NEXT, R_BND, R_LEN, R_DEPTH, R_HASH, R_RNCNT, R_RSNODE, R_RSSLOT, R_RSEOSG, R_RSRNODE,
         R_RSCNT, R_CONN, R_BNDHEAD, R_CTX, R_CHNODE, R_CHSLOT, R_CHBND, R_CHID, R_CHLEN, R_SCLIST, R_SCTOTAL, R_LVLCNT, R_LVLOFF, R_LVLLIST, R_LVLREC, R_CTXOFF, R_RNSENT, R_CURSOR, R_CUBTMP,
         NUM_BUFS };
  DevBuf buf[NUM_BUFS];
  BatchView bv{};
  WaveView wv{};
  std::vector<u32> waveActive;   // [maxNb] sentences that have local boundary b
  std::vector<u32> waveNodeOff;  // [maxNb+1] host copy
  RnnWork rw{};
  bool resident = false;
  u32 walkLaunches = 0;
  u32 maxNb = 0;
  size_t rnTotal = 0;
  jpp_b200_counters counters{};
  std::vector<u8> hostStatus;    // [sentences of the resident window] jpp_b200_sentence_status bits
  u32 failedSentences = 0;
  cudaEvent_t ev[10] = {};
};

namespace {

jpp_b200_status fail(jpp_b200_handle* h, jpp_b200_status st, const std::string& msg) {
  if (h) h->error = msg; else g_create_error = msg;
  return st;
}

// no C++ exception crosses the C ABI
template <typename F>
jpp_b200_status guarded(jpp_b200_handle* h, F&& f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    try { return fail(h, JPP_B200_INVALID_STATE, "out of host memory"); } catch (...) { return JPP_B200_INVALID_STATE; }
  } catch (const std::exception& e) {
    try { return fail(h, JPP_B200_INVALID_STATE, std::string("unexpected exception: ") + e.what()); } catch (...) { return JPP_B200_INVALID_STATE; }
  } catch (...) {
    return JPP_B200_INVALID_STATE;
  }
}

// grows a device array, keeping its first `keep` elements
template <typename T>
cudaError_t grow_device(T** p, u32* cap, u32 want, u32 keep, cudaStream_t st) {
  if (want <= *cap) return cudaSuccess;
  u32 ncap = std::max<u32>(want, *cap + *cap / 2 + 1024);
  T* np = nullptr;
  cudaError_t e = cudaMalloc(&np, (size_t)ncap * sizeof(T));
  if (e != cudaSuccess) return e;
  if (*p && keep) e = cudaMemcpyAsync(np, *p, (size_t)keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (*p) cudaFree(*p);
  *p = np;
  *cap = ncap;
  return e;
}

#define CU_CHECK(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(h, JPP_B200_CUDA_ERROR, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
  } while (0)

// JPP_B200_PROFILE=1: CUDA events around every launch of the wave loop, summed per kernel and
// printed to stderr (diagnostic only: the extra events perturb the step a little)
struct WaveProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  size_t used = 0;
  void begin(cudaStream_t st, int k) {
    if (!on) return;
    if (used + 2 > ev.size()) {
      ev.resize(used + 2);
      cudaEventCreate(&ev[used]);
      cudaEventCreate(&ev[used + 1]);
    }
    cudaEventRecord(ev[used], st);
    kind.push_back(k);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(ev[used + 1], st);
    used += 2;
  }
  void report(const char* const* names, int nk) {
    if (!on) return;
    std::vector<double> ms(nk, 0.0);
    std::vector<int> cnt(nk, 0);
    for (size_t i = 0; i < kind.size(); ++i) {
      float t = 0;
      cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
      ms[kind[i]] += t;
      cnt[kind[i]] += 1;
    }
    for (int k = 0; k < nk; ++k)
      if (cnt[k]) fprintf(stderr, "[jpp_b200 profile] %-12s %5d launches %9.3f ms\n", names[k], cnt[k], ms[k]);
    used = 0;
    kind.clear();
  }
};

// The perceptron table (256 MiB on the bench model) does not fit the 126 MB L2 and competes there
// with gigabytes of streamed lattice data.  JPP_B200_L2_PERSIST=<MiB> sets aside that much L2 for
// persisting lines and marks a matching fraction of the table's lines persisting on the scoring
// stream (cudaAccessPolicyWindow); everything else on the stream stays normal.  Off by default.
void apply_l2_policy(jpp_b200_handle* h, cudaStream_t st, size_t tableBytes, const cudaDeviceProp* prop) {
  const char* e = getenv("JPP_B200_L2_PERSIST");
  if (!e || atoi(e) <= 0) return;
  size_t want = (size_t)atoi(e) << 20;
  want = std::min(want, (size_t)prop->persistingL2CacheMaxSize);
  if (want == 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return; }
  cudaStreamAttrValue attr{};
  const size_t window = std::min(tableBytes, (size_t)prop->accessPolicyMaxWindowSize);
  attr.accessPolicyWindow.base_ptr = h->dWeights;
  attr.accessPolicyWindow.num_bytes = window;
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)window);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
  fprintf(stderr, "[jpp_b200] L2 persisting: %zu MiB set aside (max %d MiB), window %zu MiB (max %d MiB), hit ratio %.2f\n",
          want >> 20, prop->persistingL2CacheMaxSize >> 20, window >> 20, prop->accessPolicyMaxWindowSize >> 20,
          attr.accessPolicyWindow.hitRatio);
}

jpp_b200_status check_config(jpp_b200_handle* h, const jpp_b200_config* c, u32 numScorers) {
  if (c->beam_size <= 0) return fail(h, JPP_B200_INVALID_PARAMETER, "beam size can not be zero for scoring");  // analyzer_impl.cc:53-56
  if (c->right_gbeam_check > 0 && c->right_gbeam_size <= 0)
    return fail(h, JPP_B200_INVALID_PARAMETER, "right global beam size should not be zero if you enable it");  // :58-61
  if (numScorers > 1 && c->global_beam_size <= 0)
    return fail(h, JPP_B200_INVALID_STATE, "additional scorers are supported only with global beam enabled");  // :83-88
  if (c->beam_size > MAX_BEAM || c->global_beam_size > MAX_BEAM || c->right_gbeam_size > MAX_BEAM || c->right_gbeam_check > MAX_BEAM)
    return fail(h, JPP_B200_NOT_IMPLEMENTED, "beam sizes above 32 are not supported by the GPU path");
  return JPP_B200_OK;
}

}  // namespace

extern "C" {

uint32_t jpp_b200_abi_version(void) { return JPP_B200_ABI_VERSION; }

const char* jpp_b200_build_info(void) {
  return "jpp_b200 sm_100a cuda-runtime " JPP_STR(CUDART_VERSION) " fmad=off";
}

const char* jpp_b200_last_error(const jpp_b200_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

static jpp_b200_status create_impl(const jpp_b200_model* model, const jpp_b200_config* config, int device,
                                jpp_b200_handle** out) {
  if (!model || !config || !out) return fail(nullptr, JPP_B200_INVALID_PARAMETER, "null argument");
  *out = nullptr;
  HostModel hm;
  std::string err;
  jpp_b200_status st = compile_host_model(model, &hm, &err);
  if (st != JPP_B200_OK) return fail(nullptr, st, err);
  st = check_config(nullptr, config, model->num_scorers);
  if (st != JPP_B200_OK) return st;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, JPP_B200_CUDA_ERROR,
                std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(nullptr, JPP_B200_INVALID_PARAMETER, "device index out of range");
  jpp_b200_handle* h = new (std::nothrow) jpp_b200_handle();
  if (!h) return fail(nullptr, JPP_B200_INVALID_STATE, "out of host memory");
  h->device = device;
  h->hm = hm;
  h->pubCfg = *config;
  h->cfg = ScoreCfg{config->beam_size, config->global_beam_size, config->right_gbeam_check, config->right_gbeam_size,
                    model->num_scorers, hm.ft.dynamicOrder};
  h->scoreW[0] = model->score_weights[0];
  h->scoreW[1] = model->score_weights[1];
  // the compiled-in program is used when the uploaded one equals it; the switches are read when the handle is made
  // (A/B measurements, and the tests that hold the table-driven / FFMA kernels to the same golden vectors)
  const bool noStatic = getenv("JPP_B200_NO_STATIC") != nullptr;
  h->staticProgram = matches_static_program(h->hm.ft) && !noStatic;
  h->staticT0 = matches_static_t0(h->hm.ft) && !noStatic && getenv("JPP_B200_NO_STATIC_T0") == nullptr;
  h->rnnTc = !(getenv("JPP_B200_RNN_TC") && atoi(getenv("JPP_B200_RNN_TC")) == 0);
#define CREATE_CHECK(expr)                                                                      \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      g_create_error = std::string(#expr) + ": " + cudaGetErrorString(e_);                     \
      jpp_b200_destroy(h);                                                                      \
      return JPP_B200_CUDA_ERROR;                                                               \
    }                                                                                           \
  } while (0)
  CREATE_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_CHECK(cudaGetDeviceProperties(&prop, device));
  h->smCount = prop.multiProcessorCount;
  // L2 <- DRAM fetch granularity.  A miss of a 4-byte table read moves far more than its 32-byte sector: the
  // gather probe reads 86 B from DRAM per gather at an 11 % L2 hit rate (ncu, profiles/r2i_probe.md) and is at
  // 97 % of the HBM rate at 74 G gathers/s.  JPP_B200_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity
  // (a hint, device-wide for this context) before anything is allocated.
  if (const char* e = getenv("JPP_B200_L2_FETCH")) {
    const int g = atoi(e);
    if (g == 32 || g == 64 || g == 128) {
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g) != cudaSuccess) cudaGetLastError();
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      fprintf(stderr, "[jpp_b200] L2 fetch granularity: asked %d, device reports %zu\n", g, got);
    }
  }
  CREATE_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_CHECK(cudaFuncSetAttribute(k_w_prescore, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kCacheSlots * 8)));
  CREATE_CHECK(cudaFuncSetAttribute(k_t0<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_PRIM * 128 * 8)));
  CREATE_CHECK(cudaFuncSetAttribute(k_t0<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_PRIM * 128 * 8)));
  CREATE_CHECK(cudaFuncSetAttribute(k_w_items, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kCacheSlots * 8)));
  for (auto& e : h->ev) CREATE_CHECK(cudaEventCreate(&e));
  CREATE_CHECK(cudaMalloc(&h->dFt, sizeof(FeatTables)));
  CREATE_CHECK(cudaMemcpyAsync(h->dFt, &h->hm.ft, sizeof(FeatTables), cudaMemcpyHostToDevice, h->stream));
  const size_t nw = size_t{1} << model->log2_weights;
  CREATE_CHECK(cudaMalloc(&h->dWeights, nw * sizeof(float)));
  if (model->weights) {
    CREATE_CHECK(cudaMemcpyAsync(h->dWeights, model->weights, nw * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  } else {
    k_fill_weights<<<h->smCount * 8, 256, 0, h->stream>>>(h->dWeights, nw, model->weight_seed, model->weight_scale);
    CREATE_CHECK(cudaGetLastError());
  }
  if (model->num_scorers == 2) {
    const jpp_b200_rnn_model* r = model->rnn;
    h->hasRnn = true;
    const size_t sizes[4] = {(size_t)r->layer_size * r->layer_size, (size_t)r->vocab_size * r->layer_size,
                             (size_t)r->vocab_size * r->layer_size, (size_t)r->maxent_size};
    const float* src[4] = {r->matrix, r->embeddings, r->nce_embeddings, r->maxent};
    for (int i = 0; i < 4; ++i) {
      CREATE_CHECK(cudaMalloc(&h->dRnnBlocks[i], sizes[i] * sizeof(float)));
      CREATE_CHECK(cudaMemcpyAsync(h->dRnnBlocks[i], src[i], sizes[i] * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    {
      const u32 H = r->layer_size;
      std::vector<float> wt((size_t)H * H);
      for (u32 i = 0; i < H; ++i)
        for (u32 k = 0; k < H; ++k) wt[(size_t)k * H + i] = r->matrix[(size_t)i * H + k];
      CREATE_CHECK(cudaMalloc(&h->dRnnWt, wt.size() * sizeof(float)));
      CREATE_CHECK(cudaMemcpy(h->dRnnWt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
      if (H == kTcH) {
        // W = Wh + Wl (top 19 bits + exact remainder), each plane in the shared-memory image of the UMMA A operand
        std::vector<float> img(2 * (size_t)H * H, 0.0f);
        for (u32 i = 0; i < H; ++i)
          for (u32 k = 0; k < H; ++k) {
            const float v = r->matrix[(size_t)i * H + k];
            u32 bits;
            memcpy(&bits, &v, 4);
            bits &= 0xFFFFE000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const float lo = v - hi;
            const u32 o = tc_operand_offset(H, i, k) / 4;
            img[o] = hi;
            img[(size_t)H * H + o] = lo;
          }
        CREATE_CHECK(cudaMalloc(&h->dRnnWImage, img.size() * sizeof(float)));
        CREATE_CHECK(cudaMemcpy(h->dRnnWImage, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
      }
      if (32 * H * sizeof(float) > 200 * 1024) {
        g_create_error = "rnn layer too large for the level kernel's shared-memory tile";
        jpp_b200_destroy(h);
        return JPP_B200_NOT_IMPLEMENTED;
      }
      CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(32 * H * sizeof(float))));
      if (H % 32 == 0 && H / 32 <= 8) {
        const int wsBytes = (int)(H * H * sizeof(float) / rnn_level_split(H));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
        CREATE_CHECK(cudaFuncSetAttribute(k_rnn_level_x<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsBytes));
      }
    }
    h->rt = RnnTables{r->layer_size, r->maxent_order, r->maxent_size, r->vocab_size, r->nce_constant, r->unk_constant,
                      r->unk_length_penalty, r->unk_id, h->dRnnBlocks[0], h->dRnnBlocks[1], h->dRnnBlocks[2],
                      h->dRnnBlocks[3]};
  }
  CREATE_CHECK(cudaStreamSynchronize(h->stream));
#undef CREATE_CHECK
  apply_l2_policy(h, h->stream, nw * sizeof(float), &prop);
  *out = h;
  return JPP_B200_OK;
}

jpp_b200_status jpp_b200_set_config(jpp_b200_handle* h, const jpp_b200_config* config) {
  if (!h) return JPP_B200_INVALID_PARAMETER;
  if (!config) return fail(h, JPP_B200_INVALID_PARAMETER, "null config");
  jpp_b200_status st = check_config(h, config, h->cfg.S);
  if (st != JPP_B200_OK) return st;
  h->pubCfg = *config;
  h->cfg.beam = config->beam_size;
  h->cfg.gbeam = config->global_beam_size;
  h->cfg.rcheck = config->right_gbeam_check;
  h->cfg.rbeam = config->right_gbeam_size;
  h->resident = false;  // result / scratch shapes depend on the configuration
  for (jpp_b200_handle* l = h->lane; l; l = l->lane) { l->pubCfg = h->pubCfg; l->cfg = h->cfg; l->resident = false; }
  return JPP_B200_OK;
}

void jpp_b200_destroy(jpp_b200_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->lane) jpp_b200_destroy(h->lane);
  delete h->prof;
  delete h->computeToken;
  for (auto& b : h->buf) b.release();
  if (h->ownsModel) {
    if (h->dFt) cudaFree(h->dFt);
    if (h->dWeights) cudaFree(h->dWeights);
    for (auto& p : h->dRnnBlocks) if (p) cudaFree(p);
    if (h->dRnnWt) cudaFree(h->dRnnWt);
    if (h->dRnnWImage) cudaFree(h->dRnnWImage);
    if (h->dic.dRowOfPtr) cudaFree(h->dic.dRowOfPtr);
    if (h->dic.dRows) cudaFree(h->dic.dRows);
    if (h->dic.dRnnId) cudaFree(h->dic.dRnnId);
    if (h->tableStream) cudaStreamDestroy(h->tableStream);
  }
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->upStream) cudaStreamDestroy(h->upStream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// Uploads sentences [lo, hi) of `b`.  The CSR offset arrays are copied as they are and rebased on
// the device, so a window of a larger batch costs the host nothing.
// `copiesQueued` (optional) is called once every host-to-device copy of the window has been handed to the
// copy engine: a caller that runs several windows lets the next one queue its copies then, not when this
// window's index kernels have finished.
static jpp_b200_status upload_window(jpp_b200_handle* h, const jpp_b200_batch* b, u32 lo, u32 hi,
                                     const std::function<void()>* copiesQueued = nullptr) {
  if (!h || !b) return JPP_B200_INVALID_PARAMETER;
  h->resident = false;
  CU_CHECK(h, cudaSetDevice(h->device));
  if (b->num_sentences == 0 || lo >= hi || hi > b->num_sentences) return fail(h, JPP_B200_INVALID_PARAMETER, "empty batch");
  if (!b->sent_boundary_off || !b->sent_codepoint_off || !b->boundary_node_off || !b->boundary_end_off || !b->node_info ||
      !b->ends || (!b->codepoints && b->sent_codepoint_off[b->num_sentences] > 0))
    return fail(h, JPP_B200_INVALID_PARAMETER, "batch has null arrays");
  const bool residentForm = !b->node_entry && !b->node_entry_row;
  if (residentForm) {
    if (!b->sent_extra_row_off || (b->num_entry_rows > 0 && !b->entry_rows))
      return fail(h, JPP_B200_INVALID_PARAMETER, "resident form needs sent_extra_row_off (and entry_rows for the extra nodes)");
    if (h->dic.numRows == 0) return fail(h, JPP_B200_INVALID_STATE, "resident form without a resident entry table: call jpp_b200_add_entry_rows first");
    if (b->sent_extra_row_off[hi] < b->sent_extra_row_off[lo] || b->sent_extra_row_off[hi] > b->num_entry_rows)
      return fail(h, JPP_B200_INVALID_PARAMETER, "sent_extra_row_off runs past entry_rows");
  } else if (!b->node_entry && (!b->entry_rows || b->num_entry_rows == 0)) {
    return fail(h, JPP_B200_INVALID_PARAMETER, "batch has null arrays");
  }
  const bool rowPh = !b->node_placeholders && !b->node_entry && b->row_placeholders;
  const bool rowRnn = !b->node_rnn_id && !b->node_entry && b->row_rnn_id;
  if (h->hm.ft.nPlaceholders > 0 && !b->node_placeholders && !rowPh && !(residentForm && b->num_entry_rows == 0))
    return fail(h, JPP_B200_INVALID_PARAMETER, "model has placeholders but batch.node_placeholders (or row_placeholders) is null");
  if (h->hasRnn && !b->node_rnn_id && !rowRnn && !(residentForm && b->num_entry_rows == 0))
    return fail(h, JPP_B200_INVALID_PARAMETER, "model has an RNN but batch.node_rnn_id (or row_rnn_id) is null");
  const u32 NS = hi - lo;
  const u32* sbo = b->sent_boundary_off + lo;
  const u32* sco = b->sent_codepoint_off + lo;
  const u32 bBase = sbo[0], cBase = sco[0];
  // the sentence -> boundary offsets size everything else: they must be sound as a whole.  What
  // hangs off them (node / end / codepoint offsets, indices) is checked per sentence on the device.
  for (u32 s = 0; s < NS; ++s)
    if (sbo[s + 1] < sbo[s]) return fail(h, JPP_B200_INVALID_PARAMETER, "sent_boundary_off is not monotonic");
  const u32 NB = sbo[NS] - bBase;
  const u32* bno = b->boundary_node_off + bBase;
  const u32* beo = b->boundary_end_off + bBase;
  const u32 nBase = bno[0], eBase = beo[0];
  if (bno[NB] < nBase || beo[NB] < eBase || sco[NS] < cBase)
    return fail(h, JPP_B200_INVALID_PARAMETER, "boundary_node_off / boundary_end_off / sent_codepoint_off end before they start");
  const u32 NN = bno[NB] - nBase;
  const u32 NE = beo[NB] - eBase;
  const u32 NC = sco[NS] - cBase;
  h->windowSentBase = lo; h->windowBndBase = bBase; h->windowNodeBase = nBase;
  const u32 B = (u32)h->cfg.beam, G = (u32)std::max(1, h->cfg.gbeam), S = h->cfg.S, SP = h->hm.ft.nStored;
  const u32 RC = h->cfg.rcheck > 0 ? (u32)h->cfg.rcheck : 1u;
  // work order: longest sentence first; validate shapes on the way
  std::vector<u32> order(NS);
  u32 maxNb = 0;
  size_t rnTotal = 0;
  std::vector<u32> rnOff(NS + 1);
  h->hostStatus.assign(((size_t)NS + 3) & ~(size_t)3, 0);
  auto effNb = [&](u32 s) -> u32 { return h->hostStatus[s] ? 0u : sbo[s + 1] - sbo[s]; };
  for (u32 s = 0; s < NS; ++s) {
    order[s] = s;
    const u32 nb = sbo[s + 1] - sbo[s];
    if (nb < 3 || nb > 65535u) h->hostStatus[s] = JPP_B200_SENT_BAD_SHAPE;  // this sentence fails, the batch goes on
    maxNb = std::max(maxNb, effNb(s));
    rnOff[s] = (u32)rnTotal;
    rnTotal += (size_t)G * (effNb(s) + 1) + 4;
  }
  rnOff[NS] = (u32)rnTotal;
  if (rnTotal > 0xFFFFFFF0ull) return fail(h, JPP_B200_INVALID_PARAMETER, "batch too large for the RNN workspace; split it");
  std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 c) { return effNb(a) > effNb(c); });
  h->maxNb = maxNb;
  h->rnTotal = rnTotal;
  auto& bf = h->buf;
#define ENSURE(idx, bytes) CU_CHECK(h, bf[jpp_b200_handle::idx].ensure(bytes))
  ENSURE(B_SBO, (size_t)(NS + 1) * 4); ENSURE(B_SCO, (size_t)(NS + 1) * 4);
  ENSURE(B_BNO, (size_t)(NB + 1) * 4); ENSURE(B_BEO, (size_t)(NB + 1) * 4);
  ENSURE(B_INFO, (size_t)NN * sizeof(jpp_b200_node_info));
  const bool compactEntries = b->node_entry == nullptr;
  // rows that travel with the batch: all of them (compact form) or the window's extra nodes (resident form)
  const u32 xBase = residentForm ? b->sent_extra_row_off[lo] : 0u;
  const u32 batchRows = residentForm ? b->sent_extra_row_off[hi] - xBase : b->num_entry_rows;
  ENSURE(B_ENTRY, (size_t)std::max(1u, compactEntries ? batchRows : NN) * h->hm.ft.entrySize * 4);
  if (compactEntries && !residentForm) ENSURE(B_ENTRYROW, (size_t)NN * 4);
  if (residentForm) ENSURE(B_EXTRAOFF, (size_t)(NS + 1) * 4);
  ENSURE(B_PH, (size_t)std::max(1u, rowPh ? batchRows : NN) * std::max(1u, h->hm.ft.nPlaceholders) * 4);
  ENSURE(B_RNNID, (size_t)std::max(1u, rowRnn ? batchRows : NN) * 4);
  ENSURE(B_ENDS, (size_t)NE * sizeof(jpp_b200_node_ptr));
  ENSURE(B_CPS, (size_t)std::max(1u, NC) * sizeof(jpp_b200_codepoint));
  ENSURE(B_NODEBND, (size_t)NN * 4); ENSURE(B_BNDSENT, (size_t)NB * 4); ENSURE(B_ENDNODE, (size_t)NE * 4);
  ENSURE(B_ORDER, (size_t)NS * 4);
  ENSURE(B_BEAM, (size_t)NN * B * sizeof(jpp_b200_beam_element));
  ENSURE(B_BSCORES, (size_t)NN * B * S * 4);
  ENSURE(B_PATS, (size_t)NN * SP * 8);
  ENSURE(B_T0, (size_t)NN * 4);
  ENSURE(B_PRE, (size_t)NN * RC * 4);
  ENSURE(B_CUT, (size_t)NN * 4); ENSURE(B_CUTIDX, (size_t)NN * 4); ENSURE(B_SELPOS, (size_t)NN * 4); ENSURE(B_BI0, (size_t)NN * 4);
  ENSURE(B_GBNODE, (size_t)NB * G * 4); ENSURE(B_GBSLOT, (size_t)NB * G * 2); ENSURE(B_GBLBND, (size_t)NB * G * 2);
  ENSURE(B_GBCNT, (size_t)NB);
  ENSURE(B_EOSG, (size_t)NS * G * 8); ENSURE(B_EOSLOC, (size_t)NS * G * 4); ENSURE(B_EOSRNN, (size_t)NS * G * 4);
  ENSURE(B_COUNTERS, 64); ENSURE(B_WORK, 16);
  ENSURE(B_PATHLEN, (size_t)NS * 4); ENSURE(B_PATHNODES, (size_t)NB * 4);
  ENSURE(B_STATUS, h->hostStatus.size());
  // The copies and the small index kernels of a window run on a stream of the highest priority.  With several
  // lanes the device is busy with the walk of the window before (thousands of blocks waiting for ~600 slots):
  // on an equal-priority stream the index kernels of this window would be dispatched only after the last of
  // those blocks, i.e. every upload would wait for most of the previous window's walk.
  if (!h->upStream) {
    int lo = 0, hi = 0;
    CU_CHECK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU_CHECK(h, cudaStreamCreateWithPriority(&h->upStream, cudaStreamNonBlocking, hi));
  }
  static const bool flatPriority = getenv("JPP_B200_FLAT_PRIORITY") != nullptr;  // A/B: uploads on the compute stream
  cudaStream_t st = flatPriority ? h->stream : h->upStream;
#define H2D(idx, src, bytes) CU_CHECK(h, cudaMemcpyAsync(bf[jpp_b200_handle::idx].p, src, bytes, cudaMemcpyHostToDevice, st))
  CU_CHECK(h, cudaEventRecord(h->ev[0], st));
  H2D(B_SBO, sbo, (size_t)(NS + 1) * 4);
  H2D(B_SCO, sco, (size_t)(NS + 1) * 4);
  H2D(B_BNO, bno, (size_t)(NB + 1) * 4);
  H2D(B_BEO, beo, (size_t)(NB + 1) * 4);
  if (bBase | cBase | nBase | eBase) {
    k_rebase<<<(NS + 256) / 256, 256, 0, st>>>(bf[jpp_b200_handle::B_SBO].as<u32>(), NS + 1, bBase);
    k_rebase<<<(NS + 256) / 256, 256, 0, st>>>(bf[jpp_b200_handle::B_SCO].as<u32>(), NS + 1, cBase);
    k_rebase<<<(NB + 256) / 256, 256, 0, st>>>(bf[jpp_b200_handle::B_BNO].as<u32>(), NB + 1, nBase);
    k_rebase<<<(NB + 256) / 256, 256, 0, st>>>(bf[jpp_b200_handle::B_BEO].as<u32>(), NB + 1, eBase);
  }
  H2D(B_INFO, b->node_info + nBase, (size_t)NN * sizeof(jpp_b200_node_info));
  if (residentForm) {
    if (batchRows) H2D(B_ENTRY, b->entry_rows + (size_t)xBase * h->hm.ft.entrySize, (size_t)batchRows * h->hm.ft.entrySize * 4);
    H2D(B_EXTRAOFF, b->sent_extra_row_off + lo, (size_t)(NS + 1) * 4);
    if (xBase) k_rebase<<<(NS + 256) / 256, 256, 0, st>>>(bf[jpp_b200_handle::B_EXTRAOFF].as<u32>(), NS + 1, xBase);
  } else if (compactEntries) {
    H2D(B_ENTRY, b->entry_rows, (size_t)b->num_entry_rows * h->hm.ft.entrySize * 4);
    H2D(B_ENTRYROW, b->node_entry_row + nBase, (size_t)NN * 4);
  } else {
    H2D(B_ENTRY, b->node_entry + (size_t)nBase * h->hm.ft.entrySize, (size_t)NN * h->hm.ft.entrySize * 4);
  }
  if (h->hm.ft.nPlaceholders) {
    if (rowPh) { if (batchRows) H2D(B_PH, b->row_placeholders + (size_t)xBase * h->hm.ft.nPlaceholders, (size_t)batchRows * h->hm.ft.nPlaceholders * 4); }
    else H2D(B_PH, b->node_placeholders + (size_t)nBase * h->hm.ft.nPlaceholders, (size_t)NN * h->hm.ft.nPlaceholders * 4);
  }
  if (h->hasRnn) {
    if (rowRnn) { if (batchRows) H2D(B_RNNID, b->row_rnn_id + xBase, (size_t)batchRows * 4); }
    else H2D(B_RNNID, b->node_rnn_id + nBase, (size_t)NN * 4);
  }
  H2D(B_ENDS, b->ends + eBase, (size_t)NE * sizeof(jpp_b200_node_ptr));
  if (NC) H2D(B_CPS, b->codepoints + cBase, (size_t)NC * sizeof(jpp_b200_codepoint));
  H2D(B_ORDER, order.data(), (size_t)NS * 4);
  H2D(B_STATUS, h->hostStatus.data(), h->hostStatus.size());
  if (copiesQueued) (*copiesQueued)();
  BatchView& bv = h->bv;
  bv = BatchView{};
  bv.numSentences = NS; bv.totalBoundaries = NB; bv.totalNodes = NN; bv.totalEnds = NE;
#define P(idx, T) bf[jpp_b200_handle::idx].as<T>()
  bv.sentBndOff = P(B_SBO, u32); bv.sentCpOff = P(B_SCO, u32); bv.bndNodeOff = P(B_BNO, u32); bv.bndEndOff = P(B_BEO, u32);
  bv.nodeInfo = P(B_INFO, jpp_b200_node_info); bv.nodeEntry = compactEntries ? nullptr : P(B_ENTRY, i32);
  bv.nodeEntryRow = (compactEntries && !residentForm) ? P(B_ENTRYROW, u32) : nullptr; bv.entryRows = compactEntries ? P(B_ENTRY, i32) : nullptr;
  bv.numEntryRows = compactEntries ? batchRows : 0;
  if (residentForm) {
    bv.sentExtraOff = P(B_EXTRAOFF, u32);
    bv.dicRowOfPtr = h->dic.dRowOfPtr; bv.dicPtrLimit = h->dic.ptrLimit; bv.dicRows = h->dic.dRows; bv.dicRnnId = h->hasRnn ? h->dic.dRnnId : nullptr;
  }
  bv.nodePlaceholders = (h->hm.ft.nPlaceholders && !rowPh) ? P(B_PH, i32) : nullptr;
  bv.rowPlaceholders = (h->hm.ft.nPlaceholders && rowPh) ? P(B_PH, i32) : nullptr;
  bv.nodeRnnId = (h->hasRnn && !rowRnn) ? P(B_RNNID, i32) : nullptr;
  bv.rowRnnId = (h->hasRnn && rowRnn) ? P(B_RNNID, i32) : nullptr;
  bv.ends = P(B_ENDS, jpp_b200_node_ptr); bv.cps = P(B_CPS, jpp_b200_codepoint);
  bv.nodeBnd = P(B_NODEBND, u32); bv.bndSent = P(B_BNDSENT, u32); bv.endNode = P(B_ENDNODE, u32); bv.sentOrder = P(B_ORDER, u32);
  bv.beam = P(B_BEAM, jpp_b200_beam_element); bv.beamScores = P(B_BSCORES, float); bv.patterns = P(B_PATS, u64);
  // [nodes][stored] (the ABI layout).  A structure-of-arrays layout was measured and is 29 % slower
  // for the walk kernels (profiles/README.md r1i): rows of left / t2 nodes are read whole, one line each.
  bv.patNodeStride = SP; bv.patPStride = 1;
  bv.t0 = P(B_T0, float); bv.pre = P(B_PRE, float); bv.cut = P(B_CUT, float); bv.cutIdx = P(B_CUTIDX, u32);
  bv.selPos = P(B_SELPOS, u32); bv.bi0 = P(B_BI0, float); bv.gbNode = P(B_GBNODE, u32); bv.gbSlot = P(B_GBSLOT, u16); bv.gbLeftBnd = P(B_GBLBND, u16);
  bv.gbCount = P(B_GBCNT, u8); bv.eosGbeam = P(B_EOSG, u64); bv.eosLocal = P(B_EOSLOC, float); bv.eosRnn = P(B_EOSRNN, float);
  bv.counters = P(B_COUNTERS, unsigned long long);
  bv.sentStatus = P(B_STATUS, u8);
  CU_CHECK(h, cudaMemsetAsync(bv.nodeBnd, 0xFF, (size_t)NN * 4, st));  // nodes of failed sentences stay unowned
  k_boundary_index<<<(NS + 127) / 128, 128, 0, st>>>(bv);
  k_validate<<<(NB + 127) / 128, 128, 0, st>>>(bv, ValidateArgs{NC, h->hasRnn ? (u32)h->rt.vocab : 0u, h->hasRnn ? 1u : 0u, P(B_STATUS, u8)});
  CU_CHECK(h, cudaMemcpyAsync(h->hostStatus.data(), bv.sentStatus, h->hostStatus.size(), cudaMemcpyDeviceToHost, st));
  k_node_index<<<(NB + 127) / 128, 128, 0, st>>>(bv);
  CU_CHECK(h, cudaGetLastError());
  {
    // wave index (nodes grouped by local boundary) + per-sentence wave state
    ENSURE(W_GBKEY, (size_t)NS * G * 8); ENSURE(W_GBLNODE, (size_t)NS * G * 4); ENSURE(W_GBT2, (size_t)NS * G * 4);
    ENSURE(W_GBT1IDX, (size_t)NS * G); ENSURE(W_T1NODE, (size_t)NS * G * 4); ENSURE(W_GCNT, NS); ENSURE(W_UCNT, NS);
    // the node list per boundary number is only read by the wavefront kernels (and full-beam mode)
    const bool needWaveIndex = !fused_walk_ok(h->cfg) || getenv("JPP_B200_PROFILE") != nullptr;
    ENSURE(W_SELCNT, 16); ENSURE(W_NODES, needWaveIndex ? (size_t)NN * 4 + 16 : 16);
    {
      // selected right nodes of one wave: at most right_gbeam_size per active sentence (makeT0cutoffBeam)
      size_t selCap = NN;
      if (h->cfg.rcheck > 0) selCap = std::min<size_t>(NN, (size_t)NS * (u32)h->cfg.rbeam);
      // the fused walk gives every block of 32 sentences a fixed slice
      if (fused_walk_ok(h->cfg)) selCap = (size_t)((NS + 31) / 32) * 32 * (u32)h->cfg.rbeam;  // covers 16 or 32 sentences per block
      if (selCap * G >= 0x7FFFFFFFull) return fail(h, JPP_B200_INVALID_PARAMETER, "batch too large for the item lists; split it");
      ENSURE(W_BIITEMS, (selCap * G + 4) * sizeof(WaveItem)); ENSURE(W_TRIITEMS, (selCap * G + 4) * sizeof(WaveItem));
      ENSURE(W_SCB, (selCap * G + 4) * 4); ENSURE(W_TRI, (selCap * G + 4) * 8);
    }
    ENSURE(W_OFF, (size_t)(maxNb + 4) * 4); ENSURE(W_CURSOR, (size_t)(maxNb + 4) * 4 * 2);
    h->waveNodeOff.assign(maxNb + 2, 0);
    if (needWaveIndex) {
      u32* cnt = P(W_CURSOR, u32);
      u32* cursor = cnt + (maxNb + 4);
      CU_CHECK(h, cudaMemsetAsync(cnt, 0, (size_t)(maxNb + 4) * 4 * 2, st));
      k_wave_count<<<(NB + 255) / 256, 256, 0, st>>>(bv, cnt);
      size_t tmpBytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt, P(W_OFF, u32), (int)(maxNb + 2), st);
      CU_CHECK(h, bf[jpp_b200_handle::R_CUBTMP].ensure(tmpBytes + 16));
      cub::DeviceScan::ExclusiveSum(bf[jpp_b200_handle::R_CUBTMP].p, tmpBytes, cnt, P(W_OFF, u32), (int)(maxNb + 2), st);
      k_wave_fill<<<(NB + 255) / 256, 256, 0, st>>>(bv, P(W_OFF, u32), cursor, P(W_NODES, u32));
      CU_CHECK(h, cudaGetLastError());
      CU_CHECK(h, cudaMemcpyAsync(h->waveNodeOff.data(), P(W_OFF, u32), (size_t)(maxNb + 2) * 4, cudaMemcpyDeviceToHost, st));
    }
    h->waveActive.assign(maxNb + 1, 0);
    for (u32 s = 0; s < NS; ++s) {
      const u32 nb = effNb(s);
      h->waveActive[0] += 1;               // difference array: active for b < nb
      if (nb <= maxNb) h->waveActive[nb] -= 1;
    }
    for (u32 i = 1; i <= maxNb; ++i) h->waveActive[i] += h->waveActive[i - 1];
    WaveView& wv = h->wv;
    wv = WaveView{};
    wv.gbKey = P(W_GBKEY, u64); wv.gbLnode = P(W_GBLNODE, u32); wv.gbT2node = P(W_GBT2, u32); wv.gbT1idx = P(W_GBT1IDX, u8);
    wv.t1node = P(W_T1NODE, u32); wv.gCount = P(W_GCNT, u8); wv.uCount = P(W_UCNT, u8);
    wv.selCount = P(W_SELCNT, u32); wv.biItems = P(W_BIITEMS, WaveItem); wv.triItems = P(W_TRIITEMS, WaveItem); wv.scb = P(W_SCB, float); wv.tri = P(W_TRI, float); wv.waveNodes = P(W_NODES, u32); wv.waveNodeOff = P(W_OFF, u32);
  }
  if (h->hasRnn) {
    ENSURE(R_OFF, (size_t)(NS + 1) * 4);
    H2D(R_OFF, rnOff.data(), (size_t)(NS + 1) * 4);
    ENSURE(R_ID, rnTotal * 4); ENSURE(R_PREV, rnTotal * 4); ENSURE(R_NEXT, rnTotal * 4); ENSURE(R_BND, rnTotal * 4);
    ENSURE(R_LEN, rnTotal * 4); ENSURE(R_DEPTH, rnTotal * 4); ENSURE(R_HASH, rnTotal * 8); ENSURE(R_RNCNT, (size_t)NS * 4);
    ENSURE(R_RSNODE, rnTotal * 4); ENSURE(R_RSSLOT, rnTotal * 2); ENSURE(R_RSEOSG, rnTotal * 2); ENSURE(R_RSRNODE, rnTotal * 4);
    ENSURE(R_RSCNT, (size_t)NS * 4); ENSURE(R_CONN, (size_t)NN * B * 4); ENSURE(R_BNDHEAD, (size_t)NB * 4);
    ENSURE(R_LVLCNT, (size_t)(maxNb + 4) * 4); ENSURE(R_LVLOFF, (size_t)(maxNb + 4) * 4); ENSURE(R_CURSOR, (size_t)(maxNb + 4) * 4);
    ENSURE(R_LVLLIST, rnTotal * 4); if (h->dRnnWImage) ENSURE(R_LVLREC, rnTotal * 16);
    ENSURE(R_CTXOFF, (size_t)(NS + 1) * 4); ENSURE(R_RNSENT, rnTotal * 4);
    ENSURE(R_CHNODE, rnTotal * 4); ENSURE(R_CHSLOT, rnTotal * 2); ENSURE(R_CHBND, rnTotal * 2); ENSURE(R_CHID, rnTotal * 4); ENSURE(R_CHLEN, rnTotal * 4); ENSURE(R_SCLIST, rnTotal * 4);
    ENSURE(R_SCTOTAL, 16);
    RnnWork& rw = h->rw;
    rw = RnnWork{};
    rw.rnOff = P(R_OFF, u32); rw.id = P(R_ID, i32); rw.prev = P(R_PREV, i32); rw.next = P(R_NEXT, i32); rw.bnd = P(R_BND, i32);
    rw.length = P(R_LEN, i32); rw.depth = P(R_DEPTH, i32); rw.hash = P(R_HASH, u64); rw.rnCount = P(R_RNCNT, u32);
    rw.rsNode = P(R_RSNODE, u32); rw.rsSlot = P(R_RSSLOT, u16); rw.rsEosG = P(R_RSEOSG, u16); rw.rsRnode = P(R_RSRNODE, i32);
    rw.rsCount = P(R_RSCNT, u32); rw.connNode = P(R_CONN, i32); rw.bndHead = P(R_BNDHEAD, i32);
    rw.levelCount = P(R_LVLCNT, u32); rw.levelOff = P(R_LVLOFF, u32); rw.levelList = P(R_LVLLIST, u32);
    rw.chainNode = P(R_CHNODE, u32); rw.chainSlot = P(R_CHSLOT, u16); rw.chainBnd = P(R_CHBND, u16);
    rw.chainId = P(R_CHID, i32); rw.chainLen = P(R_CHLEN, i32);
    rw.scoreList = P(R_SCLIST, u32); rw.scoreTotal = P(R_SCTOTAL, u32);
    rw.maxDepth = maxNb + 1;
  }
  CU_CHECK(h, cudaEventRecord(h->ev[1], st));
  CU_CHECK(h, cudaStreamSynchronize(st));
  h->failedSentences = 0;
  for (u32 s = 0; s < NS; ++s) h->failedSentences += h->hostStatus[s] != 0;
  h->resident = true;
  return JPP_B200_OK;
#undef H2D
#undef ENSURE
}


static jpp_b200_status upload_batch_impl(jpp_b200_handle* h, const jpp_b200_batch* b) {
  if (!h || !b) return JPP_B200_INVALID_PARAMETER;
  return upload_window(h, b, 0, b->num_sentences);
}

static jpp_b200_status score_resident_impl(jpp_b200_handle* h) {
  if (!h) return JPP_B200_INVALID_PARAMETER;
  if (!h->resident) return fail(h, JPP_B200_INVALID_STATE, "no batch resident: call jpp_b200_upload_batch first");
  CU_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  BatchView& bv = h->bv;
  auto& bf = h->buf;
  const u32 NS = bv.numSentences, NN = bv.totalNodes, NB = bv.totalBoundaries;
  CU_CHECK(h, cudaMemsetAsync(bf[jpp_b200_handle::B_COUNTERS].p, 0, 64, st));
  CU_CHECK(h, cudaMemsetAsync(bf[jpp_b200_handle::B_WORK].p, 0, 16, st));
  CU_CHECK(h, cudaEventRecord(h->ev[2], st));
  {
    static const bool t0Array = getenv("JPP_B200_T0_ARRAY") != nullptr;  // A/B: all patterns in a local array
#ifdef JPP_T0_PRIM_SMEM
    const size_t primBytes = (size_t)std::max(1u, h->hm.ft.nPrim) * 128 * 8;
#else
    const size_t primBytes = 0;
#endif
    if (h->staticT0 && !t0Array) k_t0_static<<<(NN + 127) / 128, 128, 0, st>>>(h->hm.ft, bv, h->dWeights);
    else if (t0Array) k_t0<false><<<(NN + 127) / 128, 128, primBytes, st>>>(h->dFt, bv, h->dWeights);
    else k_t0<true><<<(NN + 127) / 128, 128, primBytes, st>>>(h->dFt, bv, h->dWeights);
  }
  CU_CHECK(h, cudaGetLastError());
  CU_CHECK(h, cudaEventRecord(h->ev[3], st));
  {
    const ScoreCfg& cfg = h->cfg;
    const FeatTables& ft = h->hm.ft;
    const WaveView& wv = h->wv;
    static const bool profileOn = getenv("JPP_B200_PROFILE") != nullptr;
    static const bool serialGbeam = getenv("JPP_B200_SERIAL_GBEAM") != nullptr;  // A/B switch for W1
    if (!h->prof) h->prof = new WaveProfile();
    WaveProfile& prof = *h->prof;
    prof.on = profileOn;
    enum { K_GBEAM, K_PRESCORE, K_CUTOFF, K_ITEMS, K_FINAL, K_FULL, K_KINDS };
    static const char* const kindNames[K_KINDS] = {"k_w_gbeam", "k_w_prescore", "k_w_cutoff", "k_w_items", "k_w_final", "k_w_full"};
    k_w_init<<<(NS + 127) / 128, 128, 0, st>>>(bv, cfg);
    u32 launches = 1;
    const bool fused = fused_walk_ok(cfg) && !profileOn;
    if (fused) {
      static const int sentEnv = getenv("JPP_B200_FUSED_SENT") ? atoi(getenv("JPP_B200_FUSED_SENT")) : 0;  // A/B: sentences per block
      const bool narrow = cfg.beam <= 8 && cfg.gbeam <= 8;
      const u32 SENT = sentEnv == 16 || sentEnv == 32 ? (u32)sentEnv : kFusedSentDefault;
      const u32 itemCap = SENT * (u32)cfg.rbeam * (u32)cfg.gbeam;
      const u32 blocks = (NS + SENT - 1) / SENT;
      const bool stat = h->staticProgram;
#define JPP_WALK2(GW, KB, SM, ST, SE) k_walk_fused<GW, KB, SM, ST, SE><<<blocks, SE * 8, 0, st>>>(ft, bv, cfg, wv, h->dWeights, itemCap)
#define JPP_WALK(GW, KB, SM, ST) do { if (SENT == 16) JPP_WALK2(GW, KB, SM, ST, 16); else JPP_WALK2(GW, KB, SM, ST, 32); } while (0)
      if (narrow) {
        if (small_gbeam_ok(cfg)) { if (stat) JPP_WALK(8, 8, true, true); else JPP_WALK(8, 8, true, false); }
        else { if (stat) JPP_WALK(8, 8, false, true); else JPP_WALK(8, 8, false, false); }
      } else {
        if (stat) JPP_WALK(32, MAX_BEAM, false, true); else JPP_WALK(32, MAX_BEAM, false, false);
      }
#undef JPP_WALK
#undef JPP_WALK2
      ++launches;
    }
    for (u32 bnd = 2; bnd < h->maxNb && !fused; ++bnd) {
      const u32 nAct = h->waveActive[bnd];
      if (nAct == 0) break;
      const u32 first = h->waveNodeOff[bnd], count = h->waveNodeOff[bnd + 1] - first;
      if (count == 0) continue;
      if (cfg.gbeam <= 0) {  // computeScoresFull
        prof.begin(st, K_FULL);
        k_w_full<<<(count + 127) / 128, 128, 0, st>>>(ft, bv, cfg, wv, h->dWeights, first, count);
        prof.end(st);
        ++launches;
        continue;
      }
      prof.begin(st, K_GBEAM);
      if (serialGbeam) k_w_gbeam<<<(nAct + 127) / 128, 128, 0, st>>>(bv, cfg, wv, nAct, bnd);
      else if (cfg.beam <= 8 && cfg.gbeam <= 8) k_w_gbeam_warp<8, 8><<<(nAct + 15) / 16, 128, 0, st>>>(bv, cfg, wv, nAct, bnd);
      else if (cfg.beam <= 8) k_w_gbeam_warp<8, 32><<<(nAct + 3) / 4, 128, 0, st>>>(bv, cfg, wv, nAct, bnd);
      else k_w_gbeam_warp<MAX_BEAM, 32><<<(nAct + 3) / 4, 128, 0, st>>>(bv, cfg, wv, nAct, bnd);
      prof.end(st);
      if (cfg.rcheck > 0) {
        const u32 pb = std::min<u32>((count + kCacheThreads - 1) / kCacheThreads, (u32)h->smCount);
        prof.begin(st, K_PRESCORE);
        k_w_prescore<<<std::max(1u, pb), kCacheThreads, kCacheSlots * 8, st>>>(ft, bv, cfg, wv, h->dWeights, first, count);
        prof.end(st);
      }
      prof.begin(st, K_CUTOFF);
      k_w_cutoff_group<<<(nAct + 31) / 32, 256, 0, st>>>(ft, bv, cfg, wv, nAct, bnd);
      prof.end(st);
      u32 selBound = count;
      if (cfg.rcheck > 0) selBound = std::min<u64>(count, (u64)nAct * (u32)cfg.rbeam);
      // at most (G - 1) + G items per selected node; a fraction of that in practice
      const u32 iblocks = std::min<u32>((u32)(((u64)selBound * (u32)cfg.gbeam + kCacheThreads - 1) / kCacheThreads), (u32)h->smCount);
      prof.begin(st, K_ITEMS);
      k_w_items<<<std::max(1u, iblocks), kCacheThreads, kCacheSlots * 8, st>>>(ft, bv, wv, h->dWeights);
      prof.end(st);
      prof.begin(st, K_FINAL);
      if (small_gbeam_ok(cfg)) k_w_final<true><<<(count + 127) / 128, 128, 0, st>>>(bv, cfg, wv, first, count);
      else k_w_final<false><<<(count + 127) / 128, 128, 0, st>>>(bv, cfg, wv, first, count);
      prof.end(st);
      launches += 4 + (cfg.rcheck > 0 ? 1 : 0);
    }
    if (prof.on) {
      cudaStreamSynchronize(st);
      prof.report(kindNames, K_KINDS);
    }
    CU_CHECK(h, cudaGetLastError());
    h->walkLaunches = launches;
  }
  CU_CHECK(h, cudaEventRecord(h->ev[4], st));
  if (h->hasRnn) {
    RnnWork& rw = h->rw;
    const u32 B = (u32)h->cfg.beam;
    k_fill_i32<<<h->smCount * 4, 256, 0, st>>>(rw.connNode, (size_t)NN * B, -1);
    k_fill_i32<<<h->smCount, 256, 0, st>>>(rw.bndHead, (size_t)NB, -1);
    CU_CHECK(h, cudaMemsetAsync(rw.levelCount, 0, (size_t)(h->maxNb + 4) * 4, st));
    CU_CHECK(h, cudaMemsetAsync(rw.scoreTotal, 0, 4, st));
    CU_CHECK(h, cudaMemsetAsync(bf[jpp_b200_handle::R_CURSOR].p, 0, (size_t)(h->maxNb + 4) * 4, st));
    k_rnn_tree<<<(NS + 63) / 64, 64, 0, st>>>(bv, h->cfg, rw);
    CU_CHECK(h, cudaGetLastError());
    // compact context rows: exclusive scan of the per-sentence node counts
    u32* ctxOff = bf[jpp_b200_handle::R_CTXOFF].as<u32>();
    size_t tmpBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, rw.rnCount, ctxOff, (int)NS, st);
    CU_CHECK(h, bf[jpp_b200_handle::R_CUBTMP].ensure(tmpBytes + 16));
    cub::DeviceScan::ExclusiveSum(bf[jpp_b200_handle::R_CUBTMP].p, tmpBytes, rw.rnCount, ctxOff, (int)NS, st);
    // worst case rows = rnTotal, but allocate from the actual sum: needs one small sync
    u32 lastOff = 0, lastCnt = 0;
    CU_CHECK(h, cudaMemcpyAsync(&lastOff, ctxOff + (NS - 1), 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(h, cudaMemcpyAsync(&lastCnt, rw.rnCount + (NS - 1), 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(h, cudaStreamSynchronize(st));
    const size_t rows = (size_t)lastOff + lastCnt;
    CU_CHECK(h, bf[jpp_b200_handle::R_CTX].ensure(std::max<size_t>(1, rows) * h->rt.H * 4));
    rw.ctx = bf[jpp_b200_handle::R_CTX].as<float>();
    u32* rnSent = bf[jpp_b200_handle::R_RNSENT].as<u32>();
    k_rnn_mark_sentence<<<(NS + 127) / 128, 128, 0, st>>>(bv, rw, rnSent);
    cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, rw.levelCount, rw.levelOff, (int)(h->maxNb + 2), st);
    CU_CHECK(h, bf[jpp_b200_handle::R_CUBTMP].ensure(tmpBytes + 16));
    cub::DeviceScan::ExclusiveSum(bf[jpp_b200_handle::R_CUBTMP].p, tmpBytes, rw.levelCount, rw.levelOff, (int)(h->maxNb + 2), st);
    // layer size 128: the level step as tcgen05 tensor-core MMAs (3xTF32); JPP_B200_RNN_TC=0 selects the
    // FFMA kernel (an A/B switch for measurements and for bisecting a numeric difference)
    const bool useTc = h->rnnTc && h->rt.H == kTcH && h->dRnnWImage != nullptr;
    LevelRec* levelRec = useTc ? bf[jpp_b200_handle::R_LVLREC].as<LevelRec>() : nullptr;
    k_rnn_level_fill<<<(NS * 8 + 255) / 256, 256, 0, st>>>(bv, rw, bf[jpp_b200_handle::R_CURSOR].as<u32>(), ctxOff, levelRec);
    k_rnn_bos<<<NS, 64, 0, st>>>(bv, rw, h->rt, ctxOff);
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaEventRecord(h->ev[5], st));
    {
      const u32 H = h->rt.H;
      const bool crossLane = (H % 32 == 0) && H / 32 <= 8 && !getenv("JPP_B200_RNN_LEVEL_V1");
      const u32 split = rnn_level_split(H);
      const dim3 grid((u32)h->smCount * 2 / split, split);
      const size_t wsBytes = (size_t)H * H * sizeof(float) / split;
      for (u32 level = 1; level <= h->maxNb; ++level) {
        if (useTc) {
          k_rnn_level_tc<<<h->smCount, kTcThreads, kTcSmem, st>>>(rw, h->rt, h->dRnnWImage, reinterpret_cast<const TcRec*>(levelRec), level);
          continue;
        }
        if (!crossLane) {
          k_rnn_level<8><<<h->smCount * 4, 128, 4 * 8 * H * sizeof(float), st>>>(rw, h->rt, h->dRnnWt, ctxOff, rnSent, level);
          continue;
        }
        switch (H / 32) {
          case 1: k_rnn_level_x<1><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 2: k_rnn_level_x<2><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 3: k_rnn_level_x<3><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 4: k_rnn_level_x<4><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 5: k_rnn_level_x<5><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 6: k_rnn_level_x<6><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          case 7: k_rnn_level_x<7><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
          default: k_rnn_level_x<8><<<grid, 256, wsBytes, st>>>(rw, h->rt, ctxOff, rnSent, level); break;
        }
      }
    }
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaEventRecord(h->ev[6], st));
    {
      const u32 H = h->rt.H;
      const bool coop = (H % 32 == 0) && H / 32 <= 8 && !getenv("JPP_B200_RNN_SCORES_V1");
      const u32 grid = (u32)h->smCount * 16;
#define JPP_SCORES_X(KI) k_rnn_scores_x<KI><<<grid, 128, 0, st>>>(bv, h->cfg, rw, h->rt, ctxOff, rnSent)
      if (!coop) k_rnn_scores<<<grid, 128, 0, st>>>(bv, h->cfg, rw, h->rt, ctxOff, rnSent);
      else switch (H / 32) {
        case 1: JPP_SCORES_X(1); break;
        case 2: JPP_SCORES_X(2); break;
        case 3: JPP_SCORES_X(3); break;
        case 4: JPP_SCORES_X(4); break;
        case 5: JPP_SCORES_X(5); break;
        case 6: JPP_SCORES_X(6); break;
        case 7: JPP_SCORES_X(7); break;
        default: JPP_SCORES_X(8); break;
      }
#undef JPP_SCORES_X
    }
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaEventRecord(h->ev[7], st));
    k_rerank<<<(NS * 8 + 255) / 256, 256, 0, st>>>(bv, h->cfg, h->scoreW[0], h->scoreW[1]);
    CU_CHECK(h, cudaGetLastError());
    // counters: tree nodes and scores
    h->counters.rnn_nodes = rows;
    h->counters.rnn_scores = 0;
  }
  k_best_path<<<(NS + 127) / 128, 128, 0, st>>>(bv, h->cfg, bf[jpp_b200_handle::B_PATHLEN].as<u32>(), bf[jpp_b200_handle::B_PATHNODES].as<u32>(), h->windowNodeBase);
  CU_CHECK(h, cudaGetLastError());
  CU_CHECK(h, cudaEventRecord(h->ev[8], st));
  CU_CHECK(h, cudaStreamSynchronize(st));
  // counters
  unsigned long long dc[8] = {0};
  CU_CHECK(h, cudaMemcpy(dc, bf[jpp_b200_handle::B_COUNTERS].p, 64, cudaMemcpyDeviceToHost));
  if (h->hasRnn) {
    u32 nScores = 0;
    CU_CHECK(h, cudaMemcpy(&nScores, h->rw.scoreTotal, 4, cudaMemcpyDeviceToHost));
    h->counters.rnn_scores = nScores;
  } else {
    h->counters.rnn_nodes = 0;
    h->counters.rnn_scores = 0;
  }
  jpp_b200_counters& c = h->counters;
  c.sentences = NS; c.boundaries = NB; c.nodes = NN;
  c.failed_sentences = h->failedSentences;
  // T0 runs on every node except the two BOS nodes of each sentence
  const unsigned long long t0Nodes = (unsigned long long)NN - 2ull * NS;
  c.gathers = dc[0] + t0Nodes * h->hm.ft.nUni;
  c.walk_gathers = dc[0];
  c.walk_gathers_issued = h->cfg.gbeam > 0 ? dc[2] : dc[0];
  c.walk_bytes = 4ull * dc[0] + dc[1];
  c.launches = 2 + h->walkLaunches + (h->hasRnn ? (12ull + h->maxNb) : 0ull);
  // BASELINE.md roofline accounting: 4 B per gather + streaming terms
  const u32 B = (u32)h->cfg.beam, SP = h->hm.ft.nStored;
  // T0 kernel: entry + NodeInfo + placeholders in, pattern row + score out; walk: step_counters()
  c.algorithmic_bytes = c.walk_bytes + 4ull * t0Nodes * h->hm.ft.nUni +
                        (unsigned long long)NN * (h->hm.ft.entrySize * 4 + 8 + 4 * h->hm.ft.nPlaceholders + 8ull * SP + 4);
  (void)B;
  float ms = 0;
  for (int i = 0; i < 8; ++i) c.kernel_ms[i] = 0;
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); c.kernel_ms[0] = ms;  // T0
  cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); c.kernel_ms[1] = ms;  // beam walk
  if (h->hasRnn) {
    cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); c.kernel_ms[2] = ms;  // rnn tree + lists
    cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]); c.kernel_ms[3] = ms;  // rnn contexts
    cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]); c.kernel_ms[4] = ms;  // rnn scores
    cudaEventElapsedTime(&ms, h->ev[7], h->ev[8]); c.kernel_ms[5] = ms;  // re-rank
  }
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[8]); c.kernel_ms[7] = ms;    // whole scoring
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); c.kernel_ms[6] = ms;    // upload (H2D + index kernels)
  return JPP_B200_OK;
}

static jpp_b200_status download_result_impl(jpp_b200_handle* h, jpp_b200_result* r) {
  if (!h || !r) return JPP_B200_INVALID_PARAMETER;
  if (!h->resident) return fail(h, JPP_B200_INVALID_STATE, "no batch resident");
  CU_CHECK(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const BatchView& bv = h->bv;
  const size_t NN = bv.totalNodes, B = (size_t)h->cfg.beam, S = h->cfg.S, SP = h->hm.ft.nStored;
  // the resident window's slices of the caller's arrays (whole arrays for a plain upload)
  const size_t n0 = h->windowNodeBase, s0 = h->windowSentBase, b0 = h->windowBndBase;
  if (r->node_beam) CU_CHECK(h, cudaMemcpyAsync(r->node_beam + n0 * B, bv.beam, NN * B * sizeof(jpp_b200_beam_element), cudaMemcpyDeviceToHost, st));
  if (r->node_beam_scores) CU_CHECK(h, cudaMemcpyAsync(r->node_beam_scores + n0 * B * S, bv.beamScores, NN * B * S * 4, cudaMemcpyDeviceToHost, st));
  if (r->node_patterns) CU_CHECK(h, cudaMemcpyAsync(r->node_patterns + n0 * SP, bv.patterns, NN * SP * 8, cudaMemcpyDeviceToHost, st));
  if (r->node_t0_score) CU_CHECK(h, cudaMemcpyAsync(r->node_t0_score + n0, bv.t0, NN * 4, cudaMemcpyDeviceToHost, st));
  if (r->best_path_len) CU_CHECK(h, cudaMemcpyAsync(r->best_path_len + s0, h->buf[jpp_b200_handle::B_PATHLEN].p, (size_t)bv.numSentences * 4, cudaMemcpyDeviceToHost, st));
  if (r->best_path_nodes) CU_CHECK(h, cudaMemcpyAsync(r->best_path_nodes + b0, h->buf[jpp_b200_handle::B_PATHNODES].p, (size_t)bv.totalBoundaries * 4, cudaMemcpyDeviceToHost, st));
  CU_CHECK(h, cudaStreamSynchronize(st));
  if (r->sentence_status) {
    memcpy(r->sentence_status + s0, h->hostStatus.data(), bv.numSentences);
  } else if (h->failedSentences) {
    u32 first = 0;
    while (first < bv.numSentences && !h->hostStatus[first]) ++first;
    return fail(h, JPP_B200_INVALID_PARAMETER,
                std::to_string(h->failedSentences) + " sentence(s) failed validation and were not scored (first: sentence " +
                    std::to_string(s0 + first) + ", status " + std::to_string((unsigned)h->hostStatus[first]) +
                    "); pass jpp_b200_result.sentence_status to get them per sentence");
  }
  return JPP_B200_OK;
}

namespace {

// another lane: own stream, events and buffers; the model memory of `owner`
// (defined below) L2 residency policy of the weight table on a stream
jpp_b200_status make_lane(jpp_b200_handle* owner, jpp_b200_handle* prev) {
  if (prev->lane) return JPP_B200_OK;
  jpp_b200_handle* h = owner;
  jpp_b200_handle* l = new (std::nothrow) jpp_b200_handle();
  if (!l) return fail(h, JPP_B200_INVALID_STATE, "out of host memory");
  l->device = h->device; l->hm = h->hm; l->cfg = h->cfg; l->pubCfg = h->pubCfg;
  l->dFt = h->dFt; l->dWeights = h->dWeights; l->rt = h->rt; l->hasRnn = h->hasRnn;
  for (int i = 0; i < 4; ++i) l->dRnnBlocks[i] = h->dRnnBlocks[i];
  l->staticProgram = h->staticProgram; l->staticT0 = h->staticT0; l->rnnTc = h->rnnTc;
  l->dRnnWt = h->dRnnWt; l->dRnnWImage = h->dRnnWImage; l->scoreW[0] = h->scoreW[0]; l->scoreW[1] = h->scoreW[1]; l->smCount = h->smCount;
  l->ownsModel = false;
  cudaError_t e = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking);
  for (auto& ev : l->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
  if (e != cudaSuccess) {
    jpp_b200_destroy(l);
    return fail(h, JPP_B200_CUDA_ERROR, std::string("lane: ") + cudaGetErrorString(e));
  }
  prev->lane = l;
  return JPP_B200_OK;
}

jpp_b200_status run_window(jpp_b200_handle* h, const jpp_b200_batch* batch, u32 lo, u32 hi, jpp_b200_result* result) {
  jpp_b200_status st = upload_window(h, batch, lo, hi);
  if (st != JPP_B200_OK) return st;
  st = score_resident_impl(h);
  if (st != JPP_B200_OK) return st;
  return download_result_impl(h, result);
}

constexpr int kMaxLanes = 8;

}  // namespace

// One call = H2D + all kernels + D2H.  Batches of more than a few thousand sentences are cut
// into windows of equal node count that run as concurrent lanes (a host thread, a stream and a
// set of buffers each, one model): the upload of one window hides behind the kernels of
// another and the latency-bound kernels of one lane fill the gaps of the others.  Sentences are
// independent, so the result is the same as scoring the batch in one piece.
// The lanes upload one after the other (PCIe is one pipe: window 0 complete after a sixth of the
// transfer beats six windows complete at the end) and compute side by side.
// JPP_B200_LANES=n (1..8, default 6: 108.8 / 86.7 / 84.6 / 82.3 / 82.9 ms per 100 000 sentences for
// 1 / 3 / 4 / 6 / 8 lanes, 80 ms with the batch already resident).  The handle holds no resident
// batch afterwards.
static jpp_b200_status score_batch_impl(jpp_b200_handle* h, const jpp_b200_batch* batch, jpp_b200_result* result) {
  if (!h || !batch || !result) return JPP_B200_INVALID_PARAMETER;
  if (batch->num_sentences == 0) {  // nothing to score (e.g. a rank whose shard is empty)
    h->counters = jpp_b200_counters{};
    return JPP_B200_OK;
  }
  const char* lanesEnv = getenv("JPP_B200_LANES");
  const char* minEnv = getenv("JPP_B200_LANE_MIN_SENTENCES");  // tests lower it to reach the multi-lane path
  int nLanes = lanesEnv ? atoi(lanesEnv) : 6;
  nLanes = std::max(1, std::min(nLanes, kMaxLanes));
  const u32 minSentences = minEnv ? (u32)std::max(1, atoi(minEnv)) : 2048u;  // per lane
  const u32 NS = batch->num_sentences;
  if (batch->sent_boundary_off && batch->boundary_node_off)
    while (nLanes > 1 && NS < (u64)minSentences * nLanes) --nLanes;
  else
    nLanes = 1;
  if (nLanes == 1) return run_window(h, batch, 0, NS, result);
  // window k ends where k/n of the nodes lie
  const u32* sbo = batch->sent_boundary_off;
  const u32* bno = batch->boundary_node_off;
  const u64 total = (u64)bno[sbo[NS]] - bno[sbo[0]];
  u32 cut[kMaxLanes + 1];
  cut[0] = 0;
  for (int k = 1; k < nLanes; ++k) {
    const u64 want = bno[sbo[0]] + total * k / nLanes;
    u32 a = cut[k - 1] + 1, b = NS - (u32)(nLanes - k);
    while (a < b) {
      const u32 m = a + (b - a) / 2;
      if ((u64)bno[sbo[m]] < want) a = m + 1; else b = m;
    }
    cut[k] = a;
  }
  cut[nLanes] = NS;
  jpp_b200_handle* lanes[kMaxLanes] = {h};
  for (int k = 1; k < nLanes; ++k) {
    const jpp_b200_status st = make_lane(h, lanes[k - 1]);
    if (st != JPP_B200_OK) return st;
    lanes[k] = lanes[k - 1]->lane;
    lanes[k]->cfg = h->cfg;
    lanes[k]->pubCfg = h->pubCfg;
    lanes[k]->dic = h->dic;
  }
  jpp_b200_status sts[kMaxLanes];
  for (auto& x : sts) x = JPP_B200_OK;
  std::thread workers[kMaxLanes];
  if (!h->computeToken) h->computeToken = new std::mutex();
  const char* serialEnv = getenv("JPP_B200_LANE_TURNS");
  // JPP_B200_LANE_TURNS=1: one window computes at a time.  Measured slower (100 ms against 85 ms
  // per 100 000 sentences with the kernels of the windows side by side), so it is off.
  std::mutex* token = (serialEnv && atoi(serialEnv) == 1) ? h->computeToken : nullptr;
  // windows start in order: a lane begins its upload when the previous one has finished its own
  // (PCIe is one pipe; finishing window 0 early matters more than starting all at once).  Waiting
  // lanes sleep on a condition variable: the host cores belong to the caller's staging threads.
  std::mutex turnMutex;
  std::condition_variable turnCv;
  int uploaded = 0;
  // JPP_B200_TRACE=1: one line per call on stderr with the host-clock timeline of every lane (ms since the call
  // began: upload started / uploaded / scored / downloaded) -- where an end-to-end call spends its time
  static const bool trace = getenv("JPP_B200_TRACE") != nullptr;
  const auto tCall = std::chrono::steady_clock::now();
  float tl[kMaxLanes][4] = {};
  auto since = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - tCall).count(); };
  auto work = [&](int k) noexcept {
    jpp_b200_status st = JPP_B200_OK;
    {
      std::unique_lock<std::mutex> lk(turnMutex);
      turnCv.wait(lk, [&] { return uploaded >= k; });
    }
    tl[k][0] = since();
    bool passed = false;
    const std::function<void()> pass = [&] {
      if (passed) return;
      passed = true;
      { std::lock_guard<std::mutex> lk(turnMutex); uploaded = k + 1; }
      turnCv.notify_all();
    };
    // JPP_B200_UPLOAD_EARLY=1: the next window queues its copies as soon as this one has queued its own.  Measured
    // (r2m): every upload is over by 28 ms instead of 44, but window 0 starts computing at 13 ms instead of 5 and
    // the call ends no earlier (74-79 ms against 72-75), so the turn is passed when the whole upload is done.
    static const bool wholeTurns = getenv("JPP_B200_UPLOAD_EARLY") == nullptr;
    st = guarded(lanes[k], [&] { return upload_window(lanes[k], batch, cut[k], cut[k + 1], wholeTurns ? nullptr : &pass); });
    tl[k][1] = since();
    pass();  // (an upload that failed before its copies still lets the others go on)
    if (st == JPP_B200_OK) {
      if (token) { std::lock_guard<std::mutex> turn(*token); st = guarded(lanes[k], [&] { return score_resident_impl(lanes[k]); }); }
      else st = guarded(lanes[k], [&] { return score_resident_impl(lanes[k]); });
    }
    tl[k][2] = since();
    if (st == JPP_B200_OK) st = guarded(lanes[k], [&] { return download_result_impl(lanes[k], result); });
    tl[k][3] = since();
    sts[k] = st;
  };
  int started = 1;
  try {
    for (int k = 1; k < nLanes; ++k) { workers[k] = std::thread(work, k); started = k + 1; }
  } catch (...) {
    // a worker could not be started: the ones that run finish their windows, the rest is run here
  }
  work(0);
  for (int k = started; k < nLanes; ++k) work(k);
  for (int k = 1; k < started; ++k) workers[k].join();
  if (trace) {
    std::string line = "[jpp_b200 trace] lanes (upload from, to, scored, downloaded; kernels ms):";
    char buf[160];
    for (int k = 0; k < nLanes; ++k) {
      snprintf(buf, sizeof buf, " | %d: %.1f %.1f %.1f %.1f; %.1f", k, tl[k][0], tl[k][1], tl[k][2], tl[k][3], lanes[k]->counters.kernel_ms[7]);
      line += buf;
    }
    fprintf(stderr, "%s\n", line.c_str());
  }
  for (int k = 0; k < nLanes; ++k) lanes[k]->resident = false;
  for (int k = 0; k < nLanes; ++k)
    if (sts[k] != JPP_B200_OK) {
      if (k) h->error = lanes[k]->error;
      return sts[k];
    }
  // counters of the call = all lanes (times: the longest lane)
  jpp_b200_counters& c = h->counters;
  for (int k = 1; k < nLanes; ++k) {
    const jpp_b200_counters& d = lanes[k]->counters;
    c.sentences += d.sentences; c.boundaries += d.boundaries; c.nodes += d.nodes; c.gathers += d.gathers;
    c.algorithmic_bytes += d.algorithmic_bytes; c.rnn_nodes += d.rnn_nodes; c.rnn_scores += d.rnn_scores;
    c.walk_gathers += d.walk_gathers; c.walk_bytes += d.walk_bytes; c.launches += d.launches;
    c.walk_gathers_issued += d.walk_gathers_issued; c.failed_sentences += d.failed_sentences;
    for (int i = 0; i < 8; ++i) c.kernel_ms[i] = std::max(c.kernel_ms[i], d.kernel_ms[i]);
  }
  return JPP_B200_OK;
}

jpp_b200_status jpp_b200_create(const jpp_b200_model* model, const jpp_b200_config* config, int device, jpp_b200_handle** out) {
  return guarded(nullptr, [&] { return create_impl(model, config, device, out); });
}
jpp_b200_status jpp_b200_upload_batch(jpp_b200_handle* h, const jpp_b200_batch* b) {
  return guarded(h, [&] { return upload_batch_impl(h, b); });
}
jpp_b200_status jpp_b200_score_resident(jpp_b200_handle* h) {
  return guarded(h, [&] { return score_resident_impl(h); });
}
jpp_b200_status jpp_b200_download_result(jpp_b200_handle* h, jpp_b200_result* r) {
  return guarded(h, [&] { return download_result_impl(h, r); });
}
jpp_b200_status jpp_b200_score_batch(jpp_b200_handle* h, const jpp_b200_batch* batch, jpp_b200_result* result) {
  return guarded(h, [&] { return score_batch_impl(h, batch, result); });
}

static jpp_b200_status add_entry_rows_impl(jpp_b200_handle* h, const jpp_b200_entry_table* t) {
  if (!h || !t) return JPP_B200_INVALID_PARAMETER;
  if (!h->ownsModel) return fail(h, JPP_B200_INVALID_STATE, "not the owning handle");
  if (t->num_rows == 0) return JPP_B200_OK;
  if (!t->entry_ptr || !t->rows || (h->hasRnn && !t->rnn_id)) return fail(h, JPP_B200_INVALID_PARAMETER, "entry table has null arrays");
  CU_CHECK(h, cudaSetDevice(h->device));
  const u32 es = h->hm.ft.entrySize;
  i32 maxPtr = -1;
  for (u32 i = 0; i < t->num_rows; ++i) {
    if (t->entry_ptr[i] < 0) return fail(h, JPP_B200_INVALID_PARAMETER, "entry table row with a negative entry_ptr (not a dictionary entry)");
    maxPtr = std::max(maxPtr, t->entry_ptr[i]);
  }
  auto& d = h->dic;
  if (!h->tableStream) CU_CHECK(h, cudaStreamCreateWithFlags(&h->tableStream, cudaStreamNonBlocking));
  cudaStream_t st = h->tableStream;
  // rows: new ones are appended, known entry_ptrs overwritten in place
  if (h->hostRowOfPtr.size() < (size_t)maxPtr + 1) h->hostRowOfPtr.resize((size_t)maxPtr + 1, -1);
  std::vector<u32> dst(t->num_rows);
  u32 next = d.numRows;
  for (u32 i = 0; i < t->num_rows; ++i) {
    i32& slot = h->hostRowOfPtr[t->entry_ptr[i]];
    if (slot < 0) slot = (i32)next++;
    dst[i] = (u32)slot;
  }
  CU_CHECK(h, grow_device(&d.dRows, &d.rowCap, next * es, d.numRows * es, st));
  if (h->hasRnn) CU_CHECK(h, grow_device(&d.dRnnId, &d.idCap, next, d.numRows, st));
  // contiguous runs of appended rows go up in one copy each (the common case: everything new)
  u32 i = 0;
  while (i < t->num_rows) {
    u32 j = i + 1;
    while (j < t->num_rows && dst[j] == dst[j - 1] + 1) ++j;
    CU_CHECK(h, cudaMemcpyAsync(d.dRows + (size_t)dst[i] * es, t->rows + (size_t)i * es, (size_t)(j - i) * es * 4, cudaMemcpyHostToDevice, st));
    if (h->hasRnn) CU_CHECK(h, cudaMemcpyAsync(d.dRnnId + dst[i], t->rnn_id + i, (size_t)(j - i) * 4, cudaMemcpyHostToDevice, st));
    i = j;
  }
  d.numRows = next;
  // direct map entry_ptr -> row
  const u32 limit = (u32)h->hostRowOfPtr.size();
  if (limit > d.ptrCap) {
    if (d.dRowOfPtr) cudaFree(d.dRowOfPtr);
    d.dRowOfPtr = nullptr;
    d.ptrCap = limit + limit / 4 + 1024;
    CU_CHECK(h, cudaMalloc(&d.dRowOfPtr, (size_t)d.ptrCap * 4));
    CU_CHECK(h, cudaMemcpyAsync(d.dRowOfPtr, h->hostRowOfPtr.data(), (size_t)limit * 4, cudaMemcpyHostToDevice, st));
  } else {
    // the span of the map that holds the new entries (the others in it are rewritten with the values a
    // batch in flight may be reading: harmless)
    i32 minPtr = maxPtr;
    for (u32 i = 0; i < t->num_rows; ++i) minPtr = std::min(minPtr, t->entry_ptr[i]);
    CU_CHECK(h, cudaMemcpyAsync(d.dRowOfPtr + minPtr, h->hostRowOfPtr.data() + minPtr, (size_t)(maxPtr - minPtr + 1) * 4,
                                cudaMemcpyHostToDevice, st));
  }
  if (limit > d.ptrLimit) d.ptrLimit = limit;
  CU_CHECK(h, cudaStreamSynchronize(st));
  return JPP_B200_OK;
}

static jpp_b200_status reserve_entry_rows_impl(jpp_b200_handle* h, uint32_t maxRows, uint32_t ptrLimit) {
  if (!h) return JPP_B200_INVALID_PARAMETER;
  if (!h->ownsModel) return fail(h, JPP_B200_INVALID_STATE, "not the owning handle");
  CU_CHECK(h, cudaSetDevice(h->device));
  if (!h->tableStream) CU_CHECK(h, cudaStreamCreateWithFlags(&h->tableStream, cudaStreamNonBlocking));
  auto& d = h->dic;
  const u32 es = h->hm.ft.entrySize;
  CU_CHECK(h, grow_device(&d.dRows, &d.rowCap, maxRows * es, d.numRows * es, h->tableStream));
  if (h->hasRnn) CU_CHECK(h, grow_device(&d.dRnnId, &d.idCap, maxRows, d.numRows, h->tableStream));
  if (ptrLimit > d.ptrCap) {
    if (h->hostRowOfPtr.size() < ptrLimit) h->hostRowOfPtr.resize(ptrLimit, -1);
    i32* np = nullptr;
    CU_CHECK(h, cudaMalloc(&np, (size_t)ptrLimit * 4));
    CU_CHECK(h, cudaMemcpy(np, h->hostRowOfPtr.data(), (size_t)ptrLimit * 4, cudaMemcpyHostToDevice));
    if (d.dRowOfPtr) cudaFree(d.dRowOfPtr);
    d.dRowOfPtr = np;
    d.ptrCap = ptrLimit;
    d.ptrLimit = ptrLimit;
  }
  return JPP_B200_OK;
}

jpp_b200_status jpp_b200_reserve_entry_rows(jpp_b200_handle* h, uint32_t max_rows, uint32_t entry_ptr_limit) {
  return guarded(h, [&] { return reserve_entry_rows_impl(h, max_rows, entry_ptr_limit); });
}

jpp_b200_status jpp_b200_add_entry_rows(jpp_b200_handle* h, const jpp_b200_entry_table* rows) {
  return guarded(h, [&] { return add_entry_rows_impl(h, rows); });
}

jpp_b200_status jpp_b200_alloc_pinned(size_t bytes, void** out) {
  if (!out) return JPP_B200_INVALID_PARAMETER;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    try { g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); } catch (...) {}
    return JPP_B200_CUDA_ERROR;
  }
  return JPP_B200_OK;
}

void jpp_b200_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

jpp_b200_status jpp_b200_gather_probe(jpp_b200_handle* h, uint64_t gathers, uint32_t ilp, uint32_t blocks_per_sm,
                                      double* gathers_per_second) {
  if (!h || !gathers_per_second || gathers == 0) return JPP_B200_INVALID_PARAMETER;
  return guarded(h, [&]() -> jpp_b200_status {
    *gathers_per_second = 0;
    CU_CHECK(h, cudaSetDevice(h->device));
    const u32 bps = blocks_per_sm ? std::min<u32>(blocks_per_sm, 8u) : 8u;
    const u32 blocks = (u32)h->smCount * bps;
    const u64 threads = (u64)blocks * 256u;
    const u32 unit = ilp == 4 ? 4u : (ilp == 16 ? 16u : 8u);
    const u32 perThread = (u32)std::max<u64>(unit, (gathers / threads + unit - 1) / unit * unit);
    float* sink = nullptr;
    CU_CHECK(h, cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t a = nullptr, b = nullptr;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 0.f;
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition is the warm-up
      cudaEventRecord(a, h->stream);
      const u64 seed = 0xA5A5A5A5u + (u64)rep * 977u;
      if (unit == 4) k_gather_probe<4><<<blocks, 256, 0, h->stream>>>(h->dWeights, h->hm.ft.mask, seed, perThread, sink);
      else if (unit == 16) k_gather_probe<16><<<blocks, 256, 0, h->stream>>>(h->dWeights, h->hm.ft.mask, seed, perThread, sink);
      else k_gather_probe<8><<<blocks, 256, 0, h->stream>>>(h->dWeights, h->hm.ft.mask, seed, perThread, sink);
      cudaEventRecord(b, h->stream);
      cudaEventSynchronize(b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(sink);
    CU_CHECK(h, cudaGetLastError());
    if (best > 0.f) *gathers_per_second = (double)threads * perThread / (best * 1e-3);
    return JPP_B200_OK;
  });
}

jpp_b200_status jpp_b200_get_counters(const jpp_b200_handle* h, jpp_b200_counters* out) {
  if (!h || !out) return JPP_B200_INVALID_PARAMETER;
  *out = h->counters;
  return JPP_B200_OK;
}

}  // extern "C"
