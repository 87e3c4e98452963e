// jpp_b200_kernels.h -- device code of the B200 lattice scorer.
//
// One warp walks one sentence: boundaries are sequential (the beams of boundary
// b depend on every earlier boundary: reference analyzer_impl.cc:260-284), the
// work inside a boundary (candidate scan, pattern staging, n-gram hash + weight
// gather, beam build) is spread over the 32 lanes.
//
// Lane-region discipline.  Every cross-lane exchange goes through shared
// memory between JPP_LANES / JPP_LANES_END regions (the end of a region is a
// __syncwarp()).  Code outside a region is warp-uniform: every lane computes
// the same values from shared / global data; uniform writes are guarded by
// JPP_LANE0 and followed by JPP_SYNC().  Because of that discipline the very
// same source also compiles as plain C++ with the lanes run one after another
// (tests/emul, TEST ONLY -- it lets the kernel logic be checked against the
// oracle on a machine without a GPU; it is never part of libjpp_b200.so).
//
// Float semantics: every sum replays the accumulator interleaving of the code
// the reference CLI runs (file:line at each site); this file must be compiled
// with FMA contraction off (nvcc -fmad=false) and without fast-math.
#pragma once
#include <stdint.h>
#include <string.h>

#include "jpp_b200.h"

#if defined(__CUDACC__)
#define JPP_HD __host__ __device__ __forceinline__
#define JPP_D __device__ __forceinline__
#else
#define JPP_HD inline
#define JPP_D inline
#endif

#if defined(__CUDA_ARCH__)
#define JPP_LANES { const int lane = (int)(threadIdx.x & 31u);
#define JPP_LANES_END } __syncwarp();
#define JPP_LANE0 ((threadIdx.x & 31u) == 0)
#define JPP_SYNC() __syncwarp()
#define JPP_LDG(p) __ldg(p)
#else
#define JPP_LANES for (int lane = 0; lane < 32; ++lane) {
#define JPP_LANES_END }
#define JPP_LANE0 true
#define JPP_SYNC() ((void)0)
#define JPP_LDG(p) (*(p))
#endif

namespace jppb200 {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;
typedef uint8_t u8;

#ifndef JPP_GATHER_BATCH
#define JPP_GATHER_BATCH 1
#endif
#ifndef JPP_KCHUNK
#define JPP_KCHUNK 16
#endif

// ---------------------------------------------------------------- limits ----
enum {
  MAX_UNI = 64, MAX_BI = 64, MAX_TRI = 16, MAX_PAT = 64, MAX_PRIM = 48, MAX_COMP = 64, MAX_POOL = 256,
  MAX_BEAM = 32,   // beam_size, global_beam_size, right_gbeam_size upper bound on the GPU path
  KCHUNK = JPP_KCHUNK  // right nodes scored per chunk
};

// util/seahash.h:15-17, impl/feature_impl_types.h:21-24
static const u64 kSeed0 = 0x16f11fe89b0d677cULL;
static const u64 kMult = 0x6eed0e9da4d94a4fULL;
static const u64 kPatternSeed = 0x7a11ed00000000ULL;
static const u64 kUnigramSeed = 0x5123a31421fULL;
static const u64 kBigramSeed = 0x5123a68442fULL;
static const u64 kTrigramSeed = 0x51239ab41f1fULL;

// FastHashRot::mix, util/fast_hash_rot.h:38-43
JPP_HD u64 mix(u64 state, u64 data) {
  u64 v = (state ^ data) * kMult;
  return (v << 32) | (v >> 32);
}

// Compiled feature program (host fills it from jpp_b200_model, uploaded once).
struct FeatTables {
  u32 nUni, nBi, nTri, nPat, nStored, nPrim, nComp, entrySize, nPlaceholders;
  u32 uniVars, biVars, triVars;  // accumulator counts of the generated code
  u32 mask;
  u32 pad_;
  u64 uniState[MAX_UNI];  // Hasher{}.mix(3).mix(index).mix(UnigramSeed)
  u64 biState[MAX_BI];
  u64 triState[MAX_TRI];
  u64 patState[MAX_PAT];  // Hasher{}.mix(index).mix(nrefs).mix(PatternSeed)
  i32 primKind[MAX_PRIM], primA[MAX_PRIM], primB[MAX_PRIM], primC[MAX_PRIM];
  i32 compPrim[MAX_COMP];
  u16 compTrueOff[MAX_COMP], compFalseOff[MAX_COMP];
  u8 compTrueLen[MAX_COMP], compFalseLen[MAX_COMP];
  u16 patRefOff[MAX_PAT];
  u8 patRefLen[MAX_PAT], patIndex[MAX_PAT];
  u8 uniT0[MAX_UNI];
  u8 biT0[MAX_BI], biT1[MAX_BI];
  u8 triT0[MAX_TRI], triT1[MAX_TRI], triT2[MAX_TRI];
  u8 branchPool[MAX_POOL];
  u8 patRefPool[MAX_POOL];
};

struct RnnTables {
  u32 H, order;
  u64 maxentSize, vocab;
  float nceConstant, unkConstant, unkLengthPenalty;
  i32 unkId;
  const float* matrix;
  const float* emb;
  const float* nce;
  const float* maxent;
};

// Device view of one staged batch + the result / scratch arrays.
struct BatchView {
  u32 numSentences, totalBoundaries, totalNodes, totalEnds;
  const u32* sentBndOff;
  const u32* sentCpOff;
  const u32* bndNodeOff;
  const u32* bndEndOff;
  const jpp_b200_node_info* nodeInfo;
  const i32* nodeEntry;
  const i32* nodePlaceholders;  // may be null
  const i32* nodeRnnId;         // may be null
  const jpp_b200_node_ptr* ends;
  const jpp_b200_codepoint* cps;
  // derived on upload
  u32* nodeBnd;    // [nodes] global boundary index of the node
  u32* bndSent;    // [boundaries] sentence of the boundary
  u32* endNode;    // [ends] global node index an end entry points at
  u32* sentOrder;  // [sentences] longest first (work queue order)
  // results
  jpp_b200_beam_element* beam;  // [nodes][B]
  float* beamScores;            // [nodes][B][S]
  u64* patterns;                // [nodes][stored]
  float* t0;                    // [nodes]
  // scratch
  float* pre;        // [rcheck][nodes]  T0 prescores
  float* cut;        // [nodes]          cutoff score
  u32* cutIdx;       // [nodes]          index scratch for nth_element emulation
  u16* selK;         // [nodes]          chunk-local index of a selected node / 0xFFFF
  u32* gbNode;       // [boundaries][G]  global beam elements (left node)
  u16* gbSlot;       // [boundaries][G]
  u16* gbLeftBnd;    // [boundaries][G]  sentence-local boundary of the left node
  u8* gbCount;       // [boundaries]
  u64* eosGbeam;     // [sentences][G]
  float* eosLocal;   // [sentences][G]   perceptron score of every EOS connection
  float* eosRnn;     // [sentences][G]   RNN score of every EOS connection
  unsigned long long* counters;  // [8] gathers, ...
};

struct ScoreCfg {
  i32 beam, gbeam, rcheck, rbeam;
  u32 S;  // scorers
};

JPP_HD u64 pack_cand(float score, u32 left, u32 beam) {
  u32 v;
#if defined(__CUDA_ARCH__)
  v = __float_as_uint(score);
#else
  memcpy(&v, &score, 4);
#endif
  v = (v & 0x80000000u) ? ~v : (v ^ 0x80000000u);
  return ((u64)v << 32) | ((u64)(left & 0xFFFFu) << 16) | (u64)(beam & 0xFFFFu);
}

JPP_HD float cand_score(u64 d) {
  u32 v = (u32)(d >> 32);
  v = ((v & 0x80000000u) == 0) ? ~v : (v ^ 0x80000000u);
#if defined(__CUDA_ARCH__)
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

JPP_HD bool is_fake(const jpp_b200_beam_element& e) { return e.left == 0xFFFFu && e.prev_beam == 0xFFFFu; }

JPP_HD jpp_b200_beam_element fake_elem() {
  jpp_b200_beam_element e;
  e.left = 0xFFFFu;
  e.prev_beam = 0xFFFFu;
#if defined(__CUDA_ARCH__)
  e.total_score = __uint_as_float(0xFFFFFFFFu);
#else
  u32 bits = 0xFFFFFFFFu;
  memcpy(&e.total_score, &bits, 4);
#endif
  return e;  // memset 0xff, score_processor.cc:466-468
}

// ------------------------------------------------------------ T0 kernel -----
// One thread per node: primitive features -> 38 pattern hashes -> 32 unigram
// hashes -> weight gather + ordered sum.  Replaces the generated
// patternsAndUnigramsApply (codegen/pattern_feature_codegen.cc:199-313).
JPP_HD u64 prim_value(const FeatTables& ft, u32 p, const i32* entry, const i32* ph, jpp_b200_node_info ni,
                      const jpp_b200_codepoint* cps, u32 ncp) {
  const i32 a = ft.primA[p];
  switch (ft.primKind[p]) {
    case JPP_B200_PRIM_COPY: return (u64)(u32)entry[a];
    case JPP_B200_PRIM_SHIFT_MASK: return (u64)(((u32)entry[a] >> ft.primB[p]) & (u32)ft.primC[p]);
    case JPP_B200_PRIM_PROVIDED: return ph ? (u64)(u32)ph[a] : 0;
    case JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE: return (u64)(i32)((i32)ni.end - (i32)ni.start);
    case JPP_B200_PRIM_CODEPOINT: {
      u64 value = ~(u64)0;
      if (a > 0) {
        i32 pos = (i32)ni.end + (a - 1);
        if ((u32)pos < ncp) value = cps[pos].codepoint;
      } else {
        i32 pos = (i32)ni.start + a;
        if (0 <= pos && (u32)pos < ncp) value = cps[pos].codepoint;
      }
      return value;
    }
    case JPP_B200_PRIM_CODEPOINT_TYPE: {
      u64 value = 0;
      if (a == 0) {
        for (i32 i = ni.start; i < (i32)ni.end; ++i) value |= (u32)cps[i].char_class;
      } else if (a > 0) {
        i32 pos = (i32)ni.end + (a - 1);
        if ((u32)pos < ncp) value = (u32)cps[pos].char_class;
      } else {
        i32 pos = (i32)ni.start + a;
        if (0 <= pos && (u32)pos < ncp) value = (u32)cps[pos].char_class;
      }
      return value;
    }
    default: return 0;
  }
}

JPP_HD void t0_node(const FeatTables& ft, const BatchView& bv, const float* __restrict__ w, u32 node) {
  const u32 gb = bv.nodeBnd[node];
  const u32 sent = bv.bndSent[gb];
  const u32 localBnd = gb - bv.sentBndOff[sent];
  const u32 SP = ft.nStored;
  u64* pout = bv.patterns + (size_t)node * SP;
  const jpp_b200_node_info ni = bv.nodeInfo[node];
  if (localBnd < 2) {  // BOS rows, lattice_builder.cc:171-176
    for (u32 p = 0; p < SP; ++p) pout[p] = (u64)(u32)ni.entry_ptr;
    bv.t0[node] = 0.0f;
    return;
  }
  const u32 cp0 = bv.sentCpOff[sent];
  const u32 ncp = bv.sentCpOff[sent + 1] - cp0;
  const jpp_b200_codepoint* cps = bv.cps + cp0;
  const i32* entry = bv.nodeEntry + (size_t)node * ft.entrySize;
  const i32* ph = bv.nodePlaceholders ? bv.nodePlaceholders + (size_t)node * ft.nPlaceholders : nullptr;
  u64 prim[MAX_PRIM];
  for (u32 i = 0; i < ft.nPrim; ++i) prim[i] = prim_value(ft, i, entry, ph, ni, cps, ncp);
  u64 pat[MAX_PAT];
  for (u32 p = 0; p < ft.nPat; ++p) {
    u64 h = ft.patState[p];
    const u32 off = ft.patRefOff[p];
    for (u32 k = 0; k < ft.patRefLen[p]; ++k) {
      const u32 c = ft.patRefPool[off + k];
      const u32 tl = ft.compTrueLen[c], fl = ft.compFalseLen[c];
      if (tl == 0 && fl == 0) {
        h = mix(h, prim[ft.compPrim[c]]);
      } else if (prim[ft.compPrim[c]] != 0) {
        for (u32 t = 0; t < tl; ++t) h = mix(h, prim[ft.branchPool[ft.compTrueOff[c] + t]]);
      } else {
        for (u32 t = 0; t < fl; ++t) h = mix(h, prim[ft.branchPool[ft.compFalseOff[c] + t]]);
      }
    }
    pat[ft.patIndex[p]] = h;
  }
  for (u32 p = 0; p < SP; ++p) pout[p] = pat[p];
  // unigram score.  Rows 0..R-2 of a boundary: "float score_part_j = w" for the
  // first uniVars terms then "+=" (pattern_feature_codegen.cc:229-236), summed
  // p0 + p1 + ...; the last row goes through computeUnrolled4RawPerceptron
  // (perceptron.h:46-72): four accumulators started at 0.
  const bool lastRow = (node + 1 == bv.bndNodeOff[gb + 1]);
  const u32 nv = lastRow ? 4u : ft.uniVars;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const u32 nU = ft.nUni, mask = ft.mask;
  for (u32 u0 = 0; u0 < nU; u0 += 4) {
    float wv[4];
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
      const u32 u = u0 + j;
      if (u < nU) {
        const u32 idx = (u32)mix(ft.uniState[u], pat[ft.uniT0[u]]) & mask;
        wv[j] = JPP_LDG(w + idx);
      }
    }
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
      const u32 u = u0 + j;
      if (u < nU) {
        const u32 a = u % nv;
        if (!lastRow && u < nv) acc[a] = wv[j]; else acc[a] += wv[j];
      }
    }
  }
  float s = acc[0];
  for (u32 v = 1; v < nv; ++v) s = s + acc[v];
  bv.t0[node] = s;
}

// ------------------------------------------------- exact sort emulation -----
// Equal float keys are common (homograph entries with identical feature
// fields), so the order the reference produces for them matters.  These follow
// libstdc++'s std::sort / std::nth_element and jumanpp's util::partition
// (reference src/util/stl_util.h:50-135) step by step on u32 index arrays with
// comp(a, b) := score[a] > score[b].
struct ScoreCmp {
  const float* s;
  JPP_HD bool operator()(u32 a, u32 b) const { return s[a] > s[b]; }
};

template <typename Cmp>
JPP_HD void em_insertion_sort(u32* first, u32* last, Cmp cmp) {
  if (first == last) return;
  for (u32* i = first + 1; i != last; ++i) {
    u32 val = *i;
    if (cmp(val, *first)) {
      for (u32* p = i; p != first; --p) *p = *(p - 1);
      *first = val;
    } else {
      u32* l = i;
      u32* next = i - 1;
      while (cmp(val, *next)) { *l = *next; l = next; --next; }
      *l = val;
    }
  }
}

template <typename Cmp>
JPP_HD void em_unguarded_insertion_sort(u32* first, u32* last, Cmp cmp) {
  for (u32* i = first; i != last; ++i) {
    u32 val = *i;
    u32* l = i;
    u32* next = i - 1;
    while (cmp(val, *next)) { *l = *next; l = next; --next; }
    *l = val;
  }
}

JPP_HD void em_swap(u32* a, u32* b) { u32 t = *a; *a = *b; *b = t; }

template <typename Cmp>
JPP_HD u32* em_partition_pivot(u32* first, u32* last, Cmp cmp) {
  u32* mid = first + (last - first) / 2;
  u32 *a = first + 1, *b = mid, *c = last - 1;
  if (cmp(*a, *b)) {
    if (cmp(*b, *c)) em_swap(first, b);
    else if (cmp(*a, *c)) em_swap(first, c);
    else em_swap(first, a);
  } else if (cmp(*a, *c)) em_swap(first, a);
  else if (cmp(*b, *c)) em_swap(first, c);
  else em_swap(first, b);
  u32* lo = first + 1;
  u32* hi = last;
  for (;;) {
    while (cmp(*lo, *first)) ++lo;
    --hi;
    while (cmp(*first, *hi)) --hi;
    if (!(lo < hi)) return lo;
    em_swap(lo, hi);
    ++lo;
  }
}

template <typename Cmp>
JPP_HD void em_adjust_heap(u32* first, long hole, long len, u32 value, Cmp cmp) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (cmp(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  long parent = (hole - 1) / 2;
  while (hole > top && cmp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename Cmp>
JPP_HD void em_heap_select(u32* first, u32* middle, u32* last, Cmp cmp) {
  long len = middle - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      u32 value = first[parent];
      em_adjust_heap(first, parent, len, value, cmp);
      if (parent == 0) break;
      parent--;
    }
  }
  for (u32* i = middle; i < last; ++i)
    if (cmp(*i, *first)) {
      u32 value = *i;
      *i = *first;
      em_adjust_heap(first, 0, middle - first, value, cmp);
    }
}

template <typename Cmp>
JPP_HD void em_sort_heap(u32* first, u32* last, Cmp cmp) {
  while (last - first > 1) {
    --last;
    u32 value = *last;
    *last = *first;
    em_adjust_heap(first, 0, last - first, value, cmp);
  }
}

JPP_HD int em_lg(long n) {
  int k = 0;
  while (n > 1) { n >>= 1; ++k; }
  return k;
}

// std::sort: the recursion of __introsort_loop is unrolled with a small stack
template <typename Cmp>
JPP_HD void em_std_sort(u32* first, long n, Cmp cmp) {
  if (n <= 0) return;
  u32* last = first + n;
  if (n > 16) {
    u32* stF[48];
    u32* stL[48];
    long stD[48];
    int sp = 0;
    stF[0] = first; stL[0] = last; stD[0] = 2L * em_lg(n); sp = 1;
    while (sp > 0) {
      --sp;
      u32* f = stF[sp];
      u32* l = stL[sp];
      long d = stD[sp];
      while (l - f > 16) {
        if (d == 0) {
          em_heap_select(f, l, l, cmp);
          em_sort_heap(f, l, cmp);
          break;
        }
        --d;
        u32* cut = em_partition_pivot(f, l, cmp);
        // __introsort_loop(cut, last) runs first, then the loop continues on [first, cut):
        // the two halves are independent, so deferring the left half is equivalent
        if (sp < 48) { stF[sp] = f; stL[sp] = cut; stD[sp] = d; ++sp; }
        f = cut;
      }
    }
    em_insertion_sort(first, first + 16, cmp);
    em_unguarded_insertion_sort(first + 16, last, cmp);
  } else {
    em_insertion_sort(first, last, cmp);
  }
}

template <typename Cmp>
JPP_HD void em_std_nth_element(u32* first, long nth, long n, Cmp cmp) {
  u32* last = first + n;
  u32* nthp = first + nth;
  if (n <= 0 || nthp == last) return;
  long depth = 2L * em_lg(n);
  while (last - first > 3) {
    if (depth == 0) {
      em_heap_select(first, nthp + 1, last, cmp);
      em_swap(first, nthp);
      return;
    }
    --depth;
    u32* cut = em_partition_pivot(first, last, cmp);
    if (cut <= nthp) first = cut; else last = cut;
  }
  em_insertion_sort(first, last, cmp);
}

template <typename Cmp>
JPP_HD u32* em_part_step(u32* start, u32* end, Cmp cmp) {
  long sz = end - start;
  if (sz == 1) return end;
  if (sz == 2) {
    u32* n = start + 1;
    if (cmp(*n, *start)) em_swap(n, start);
    return n;
  }
  if (sz == 3) {
    u32 *n0 = start, *n1 = start + 1, *n2 = start + 2;
    if (cmp(*n1, *n0)) em_swap(n1, n0);
    if (cmp(*n2, *n1)) em_swap(n2, n1);
    if (cmp(*n1, *n0)) em_swap(n1, n0);
    return n1;
  }
  u32* pivot = start + sz / 2;
  --end;
  em_swap(pivot, end);
  pivot = end;
  --end;
  while (start != end) {
    if (cmp(*start, *pivot)) {
      ++start;
    } else {
      em_swap(start, end);
      --end;
    }
  }
  if (cmp(*pivot, *end)) {
    em_swap(pivot, end);
  } else {
    ++end;
    em_swap(pivot, end);
  }
  return end;
}

template <typename Cmp>
JPP_HD long em_ref_partition(u32* first, long n, Cmp cmp, long minSize, long maxSize) {
  u32* start = first;
  u32* end = first + n;
  for (;;) {
    u32* mid = em_part_step(start, end, cmp);
    long sz = mid - start;
    if (minSize <= sz && sz <= maxSize) return mid - first;
    if (sz > maxSize) {
      end = mid;
      continue;
    }
    sz += 1;
    start = mid + 1;
    minSize -= sz;
    maxSize -= sz;
    if (minSize == 0) return start - first;
  }
}

// makeT0Beam ordering (score_processor.cc:426-438); returns usable length
JPP_HD int beam_order(u32* idx, int n, const float* totals, int maxBeam) {
  for (int i = 0; i < n; ++i) idx[i] = (u32)i;
  long len = n;
  const long boundary = (long)maxBeam * 4 / 3;
  ScoreCmp cmp{totals};
  if (n > boundary) len = em_ref_partition(idx, n, cmp, maxBeam, boundary);
  em_std_sort(idx, len, cmp);
  return (int)len;
}

// -------------------------------------------------------- beam walk ---------
// Shared memory of one warp (sizes depend on the scoring configuration).
struct WarpSmem {
  u64* red;        // [32]
  u64* gbeam;      // [G]   packed BeamCandidate, descending
  u32* gbLnode;    // [G]
  u32* gbT2node;   // [G]
  u32* t1node;     // [G]   unique left nodes in first-seen order
  u32* t1idx;      // [G]
  u64* t1pat;      // [G][SP]
  u64* t2pat;      // [G][SP]
  u64* t0pat;      // [KCHUNK][SP]  pattern rows of the right nodes being scored
  u32* selNode;    // [KCHUNK] right-node offsets (in boundary) of the current chunk
  float* scbuf;    // [KCHUNK][G]  bigram score per (selected node, unique t1)
  float* tot;      // [KCHUNK][G]
  float* loc;      // [KCHUNK][G]
  float* wbuf;     // [wcap] gathered weights, summed in reference order afterwards
  u32* misc;       // [8] uniform scalars
  u32 wcap;
};

// gathered-weight staging capacity (floats) for a configuration
JPP_HD u32 warp_wcap(u32 G, u32 rbeam, u32 NF) {
  u32 k = rbeam ? rbeam : 8u;
  if (k > 8u) k = 8u;
  u32 a = 16u * NF, b = k * G * NF;
  u32 c = a > b ? a : b;
  if (c < G * NF) c = G * NF;
  return c;
}

JPP_HD size_t warp_smem_bytes(u32 G, u32 SP, u32 wcap) {
  size_t b = 0;
  b += 32 * 8;                 // red
  b += (size_t)G * 8;          // gbeam
  b += (size_t)G * SP * 8 * 2; // t1pat, t2pat
  b += (size_t)KCHUNK * SP * 8;// t0pat
  b += (size_t)G * 4 * 4;      // gbLnode, gbT2node, t1node, t1idx
  b += KCHUNK * 4;             // selNode
  b += (size_t)KCHUNK * G * 4 * 3;
  b += (size_t)wcap * 4;
  b += 8 * 4;
  return (b + 15) & ~(size_t)15;
}

JPP_HD WarpSmem carve_warp_smem(unsigned char* base, u32 G, u32 SP, u32 wcap) {
  WarpSmem s;
  unsigned char* p = base;
  s.red = (u64*)p; p += 32 * 8;
  s.gbeam = (u64*)p; p += (size_t)G * 8;
  s.t1pat = (u64*)p; p += (size_t)G * SP * 8;
  s.t2pat = (u64*)p; p += (size_t)G * SP * 8;
  s.t0pat = (u64*)p; p += (size_t)KCHUNK * SP * 8;
  s.gbLnode = (u32*)p; p += (size_t)G * 4;
  s.gbT2node = (u32*)p; p += (size_t)G * 4;
  s.t1node = (u32*)p; p += (size_t)G * 4;
  s.t1idx = (u32*)p; p += (size_t)G * 4;
  s.selNode = (u32*)p; p += KCHUNK * 4;
  s.scbuf = (float*)p; p += (size_t)KCHUNK * G * 4;
  s.tot = (float*)p; p += (size_t)KCHUNK * G * 4;
  s.loc = (float*)p; p += (size_t)KCHUNK * G * 4;
  s.wbuf = (float*)p; p += (size_t)wcap * 4;
  s.misc = (u32*)p;
  s.wcap = wcap;
  return s;
}

// Ordered sums over gathered weights x[0..n): NV interleaved accumulators
// started at 0 ("f_j = 0; f_j += w", partial_ngram_feature_codegen.cc:44-66;
// NV = 4 is also computeUnrolled4RawPerceptron, perceptron.h:46-72; NV = 2 is
// the r1/r2 pair of applyBiTriFullKernel, feature_impl_ngram_partial_kernels.h:35-61).
template <int NV>
JPP_HD void osum(const float* x, u32 n, float* acc) {
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = 0.0f;
  u32 f = 0;
  for (; f + NV <= n; f += NV) {
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += x[f + j];
  }
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (f + j < n) acc[j] += x[f + j];
}

JPP_HD float osum_vars(const float* x, u32 n, u32 nv) {
  float a[8];
  switch (nv) {
    case 8: osum<8>(x, n, a); return ((((((a[0] + a[1]) + a[2]) + a[3]) + a[4]) + a[5]) + a[6]) + a[7];
    case 4: osum<4>(x, n, a); return ((a[0] + a[1]) + a[2]) + a[3];
    case 2: osum<2>(x, n, a); return a[0] + a[1];
    default: osum<1>(x, n, a); return a[0];
  }
}

JPP_HD float osum_unrolled4(const float* x, u32 n) {
  float a[4];
  osum<4>(x, n, a);
  return ((a[0] + a[1]) + a[2]) + a[3];
}

// max over the 32 per-lane values a region left in sm.red
JPP_HD u64 uniform_max_u64(const u64* red) {
#if defined(__CUDA_ARCH__)
  u64 v = red[threadIdx.x & 31u];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const u64 t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
#else
  u64 m = 0;
  for (int l = 0; l < 32; ++l) if (red[l] > m) m = red[l];
  return m;
#endif
}

#if defined(__CUDACC__)
#define JPP_NOINLINE __noinline__
#else
#define JPP_NOINLINE
#endif

// Writes the beam row of one right node from `n` candidate totals.
JPP_NOINLINE
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void write_beam(const BatchView& bv, const ScoreCfg& cfg, u32 node, const u64* gbeam, int n,
                       const float* totals, const float* local) {
  u32 idx[MAX_BEAM];
  const int maxBeam = cfg.beam;
  int len;
  if (n <= 1) {
    idx[0] = 0;
    len = n;
  } else {
    len = beam_order(idx, n, totals, maxBeam);
  }
  jpp_b200_beam_element* row = bv.beam + (size_t)node * maxBeam;
  float* srow = bv.beamScores + (size_t)node * maxBeam * cfg.S;
  int k = 0;
  for (; k < len && k < maxBeam; ++k) {
    const u32 g = idx[k];
    jpp_b200_beam_element e;
    e.left = (u16)(gbeam[g] >> 16);
    e.prev_beam = (u16)gbeam[g];
    e.total_score = totals[g];
    row[k] = e;
    srow[(size_t)k * cfg.S] = local[g];
    for (u32 q = 1; q < cfg.S; ++q) srow[(size_t)k * cfg.S + q] = 0.0f;
  }
  for (; k < maxBeam; ++k) {
    row[k] = fake_elem();
    for (u32 q = 0; q < cfg.S; ++q) srow[(size_t)k * cfg.S + q] = 0.0f;
  }
}

// One warp, one sentence: makeGlobalBeam + computeGbeamScores for every
// boundary (analyzer_impl.cc:260-284, score_processor.cc:246-361).
//
// The hot work of a boundary is flattened over the lanes as (row, feature)
// items: each lane hashes one n-gram feature and gathers one weight per
// iteration (all 32 lanes busy even when a boundary has only a few right
// nodes), staging the weights in shared memory; the float sums are then
// replayed from shared memory in the reference's accumulator order.
JPP_HD void walk_sentence(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg, const float* __restrict__ w,
                          WarpSmem sm, u32 sent) {
  const u32 b0 = bv.sentBndOff[sent];
  const u32 nb = bv.sentBndOff[sent + 1] - b0;
  const u32 B = (u32)cfg.beam;
  const u32 SP = ft.nStored;
  const u32 nB = ft.nBi, nT = ft.nTri, mask = ft.mask;
  const u32 NF = nB + nT;
  const u32 GMAX = (u32)cfg.gbeam;
  unsigned long long gathers = 0, streamBytes = 0;

  // BOS beams (analyzer_impl.cc:185-192)
  JPP_LANES
  for (u32 i = (u32)lane; i < 2 * B && nb >= 2; i += 32) {
    const u32 bb = i / B, k = i % B;
    const u32 n0 = bv.bndNodeOff[b0 + bb];
    jpp_b200_beam_element e;
    if (k == 0) { e.left = 0; e.prev_beam = 0; e.total_score = 0.0f; }
    else { e.left = 0xFFFFu; e.prev_beam = 0xFFFFu; e.total_score = -3.402823466e+38f; }  // EntryBeam::fake()
    bv.beam[(size_t)n0 * B + k] = e;
    for (u32 q = 0; q < cfg.S; ++q) bv.beamScores[((size_t)n0 * B + k) * cfg.S + q] = 0.0f;
  }
  JPP_LANES_END
  if (nb <= 3) {  // analyzer_impl.cc:255-258
    if (nb == 3) {
      JPP_LANES
      const u32 n0 = bv.bndNodeOff[b0 + 2];
      for (u32 k = (u32)lane; k < B; k += 32) {
        bv.beam[(size_t)n0 * B + k] = fake_elem();
        for (u32 q = 0; q < cfg.S; ++q) bv.beamScores[((size_t)n0 * B + k) * cfg.S + q] = 0.0f;
      }
      JPP_LANES_END
    }
    return;
  }

  for (u32 bnd = 2; bnd < nb; ++bnd) {
    const u32 gB = b0 + bnd;
    const u32 n0 = bv.bndNodeOff[gB];
    const u32 R = bv.bndNodeOff[gB + 1] - n0;
    const u32 e0 = bv.bndEndOff[gB];
    const u32 L = bv.bndEndOff[gB + 1] - e0;
    if (R == 0) {  // analyzer_impl.cc:263-266
      if (JPP_LANE0) bv.gbCount[gB] = 0;
      continue;
    }

    // ---- makeGlobalBeam (score_processor.cc:246-282): the G largest packed
    // candidates; keys are distinct, so round j takes the largest key below
    // the one taken in round j-1.
    u64 last = ~(u64)0;
    u32 G = 0;
    const u32 nc = L * B;
    // candidates are packed once into shared memory (the gathered-weight buffer is
    // idle here); a boundary with more candidates than fit re-reads the beams
    u64* candS = reinterpret_cast<u64*>(sm.wbuf);
#ifndef JPP_STAGE_CANDS
#define JPP_STAGE_CANDS 1
#endif
    const bool staged = JPP_STAGE_CANDS && nc <= sm.wcap / 2;
    if (staged) {
      JPP_LANES
      for (u32 left = (u32)lane; left < L; left += 32) {
        const u32 lnode = bv.endNode[e0 + left];
        const jpp_b200_beam_element* row = bv.beam + (size_t)lnode * B;
        for (u32 slot = 0; slot < B; ++slot) {
          const jpp_b200_beam_element el = row[slot];
          candS[left * B + slot] = is_fake(el) ? 0 : pack_cand(el.total_score, left, slot);
        }
      }
      JPP_LANES_END
    }
    for (u32 j = 0; j < GMAX; ++j) {
      JPP_LANES
      u64 best = 0;
      if (staged) {
        for (u32 c = (u32)lane; c < nc; c += 32) {
          const u64 key = candS[c];
          if (key < last && key > best) best = key;
        }
      } else {
        for (u32 left = (u32)lane; left < L; left += 32) {
          const u32 lnode = bv.endNode[e0 + left];
          for (u32 slot = 0; slot < B; ++slot) {
            const jpp_b200_beam_element el = bv.beam[(size_t)lnode * B + slot];
            if (!is_fake(el)) {
              const u64 key = pack_cand(el.total_score, left, slot);
              if (key < last && key > best) best = key;
            }
          }
        }
      }
      sm.red[lane] = best;
      JPP_LANES_END
      const u64 m = uniform_max_u64(sm.red);
      JPP_SYNC();
      if (m == 0) break;
      if (JPP_LANE0) sm.gbeam[j] = m;
      last = m;
      G = j + 1;
    }
    JPP_SYNC();

    // ---- global beam bookkeeping, t1 de-duplication, t2 lookup (:363-409)
    JPP_LANES
    if ((u32)lane < G) {
      const u64 key = sm.gbeam[lane];
      const u32 left = (u32)(key >> 16) & 0xFFFFu, slot = (u32)key & 0xFFFFu;
      const u32 lnode = bv.endNode[e0 + left];
      const u32 lbnd = bv.ends[e0 + left].boundary;
      const jpp_b200_beam_element el = bv.beam[(size_t)lnode * B + slot];
      const u32 t2node = bv.endNode[bv.bndEndOff[b0 + lbnd] + el.left];
      sm.gbLnode[lane] = lnode;
      sm.gbT2node[lane] = t2node;
      bv.gbNode[(size_t)gB * GMAX + lane] = lnode;
      bv.gbSlot[(size_t)gB * GMAX + lane] = (u16)slot;
      bv.gbLeftBnd[(size_t)gB * GMAX + lane] = (u16)lbnd;
    }
    if (lane == 0) bv.gbCount[gB] = (u8)G;
    JPP_LANES_END
    u32 U = 0;
    if (JPP_LANE0) {
      for (u32 g = 0; g < G; ++g) {
        const u32 ln = sm.gbLnode[g];
        u32 u = 0;
        for (; u < U; ++u) if (sm.t1node[u] == ln) break;
        if (u == U) { sm.t1node[U] = ln; ++U; }
        sm.t1idx[g] = u;
      }
      sm.misc[1] = U;
    }
    JPP_SYNC();
    U = sm.misc[1];
    // algorithmic streaming bytes of this boundary (DESIGN.md "roofline accounting"):
    // left beams read, pattern rows of right / unique-t1 / t2 nodes, ends, beams + scores written
    streamBytes += 8ull * L * B + 8ull * SP * ((unsigned long long)R + U + G) + 4ull * L +
                   (8ull + 4ull * cfg.S) * R * B;
    JPP_LANES
    for (u32 i = (u32)lane; i < U * SP; i += 32) sm.t1pat[i] = bv.patterns[(size_t)sm.t1node[i / SP] * SP + i % SP];
    for (u32 i = (u32)lane; i < G * SP; i += 32) sm.t2pat[i] = bv.patterns[(size_t)sm.gbT2node[i / SP] * SP + i % SP];
    JPP_LANES_END

    const bool rightCheck = cfg.rcheck > 0;
    u32 full = 0, toKeep = R, remaining = G;
    if (rightCheck) {
      full = (u32)cfg.rcheck;
      if (R < full) full = R;
      if (G < full) full = G;  // fullBeamApplySize, score_processor.cc:299-301
      toKeep = (u32)cfg.rbeam < R ? (u32)cfg.rbeam : R;
      remaining = G - full;
      const u32 theMax = (u32)cfg.rcheck < G ? (u32)cfg.rcheck : G;  // computeT0Prescores :497-511

      // ---- T0 prescores: T0 + bigrams (biVars accumulators; last row Unrolled4)
      // + trigrams (partial_ngram_feature_codegen.cc:13-153)
      u32 RC = sm.wcap / NF;
      if (RC > KCHUNK) RC = KCHUNK;
      for (u32 i = 0; i < theMax; ++i) {
        const u64* t1 = sm.t1pat + (size_t)sm.t1idx[i] * SP;
        const u64* t2 = sm.t2pat + (size_t)i * SP;
        for (u32 r0 = 0; r0 < R; r0 += RC) {
          const u32 rc = (R - r0) < RC ? (R - r0) : RC;
          JPP_LANES
          const u64* src = bv.patterns + (size_t)(n0 + r0) * SP;
          for (u32 x = (u32)lane; x < rc * SP; x += 32) sm.t0pat[x] = src[x];
          JPP_LANES_END
          JPP_LANES
          u32 row = 0, f = (u32)lane;
          while (f >= NF) { f -= NF; ++row; }
          const u32 total = rc * NF;
          for (u32 it = (u32)lane; it < total; it += 32u * JPP_GATHER_BATCH) {
            u32 idx[JPP_GATHER_BATCH];
            float v[JPP_GATHER_BATCH];
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q) {
              idx[q] = 0;
              if (it + 32u * q < total) {
                const u64* P = sm.t0pat + (size_t)row * SP;
                if (f < nB) {
                  idx[q] = (u32)mix(mix(ft.biState[f], P[ft.biT0[f]]), t1[ft.biT1[f]]) & mask;
                } else {
                  const u32 g = f - nB;
                  idx[q] = (u32)mix(mix(mix(ft.triState[g], P[ft.triT0[g]]), t1[ft.triT1[g]]), t2[ft.triT2[g]]) & mask;
                }
              }
              f += 32;
              while (f >= NF) { f -= NF; ++row; }
            }
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q) v[q] = JPP_LDG(w + idx[q]);
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q)
              if (it + 32u * q < total) sm.wbuf[it + 32u * q] = v[q];
          }
          JPP_LANES_END
          JPP_LANES
          for (u32 r = (u32)lane; r < rc; r += 32) {
            const u32 node = n0 + r0 + r;
            const bool lastRow = (r0 + r + 1 == R);
            const float* x = sm.wbuf + (size_t)r * NF;
            float sc = bv.t0[node];
            sc += lastRow ? osum_unrolled4(x, nB) : osum_vars(x, nB, ft.biVars);
            sc += lastRow ? osum_unrolled4(x + nB, nT) : osum_vars(x + nB, nT, ft.triVars);
            bv.pre[(size_t)i * bv.totalNodes + node] = sc;
          }
          JPP_LANES_END
        }
      }
      gathers += (unsigned long long)theMax * R * NF;

      // ---- makeT0cutoffBeam (:471-495): which toKeep right nodes get full beams
      if (R > toKeep) {
        JPP_LANES
        for (u32 r = (u32)lane; r < R; r += 32) {
          float v = 0;
          for (u32 j = 0; j < full; ++j) v += bv.pre[(size_t)j * bv.totalNodes + n0 + r];
          bv.cut[n0 + r] = v;
        }
        JPP_LANES_END
        if (JPP_LANE0) sm.misc[2] = 0;
        JPP_SYNC();
        JPP_LANES
        for (u32 r = (u32)lane; r < R; r += 32) {
          const float mine = bv.cut[n0 + r];
          u32 gt = 0, ge = 0;
          for (u32 j = 0; j < R; ++j) {
            const float o = bv.cut[n0 + j];
            gt += (o > mine);
            ge += (o >= mine);
          }
          // ge <= toKeep: certainly inside the top toKeep; gt >= toKeep: certainly outside;
          // otherwise a run of equal scores straddles the cut -> replay nth_element
          u16 v = 0xFFFFu;
          if (ge <= toKeep) v = 1;
          else if (gt < toKeep) { v = 2; sm.misc[2] = 1; }
          bv.selK[n0 + r] = v;
        }
        JPP_LANES_END
        if (sm.misc[2] != 0) {
          if (JPP_LANE0) {
            u32* idx = bv.cutIdx + n0;
            for (u32 r = 0; r < R; ++r) idx[r] = r;
            em_std_nth_element(idx, (long)toKeep, (long)R, ScoreCmp{bv.cut + n0});
            for (u32 r = 0; r < R; ++r) bv.selK[n0 + r] = 0xFFFFu;
            for (u32 k = 0; k < toKeep; ++k) bv.selK[n0 + idx[k]] = 1;
          }
          JPP_SYNC();
        }
      } else {
        JPP_LANES
        for (u32 r = (u32)lane; r < R; r += 32) bv.selK[n0 + r] = 1;
        JPP_LANES_END
      }
    } else {
      JPP_LANES
      for (u32 r = (u32)lane; r < R; r += 32) bv.selK[n0 + r] = 1;
      JPP_LANES_END
    }

    // ---- selected right nodes, a chunk at a time: full scoring against the
    // global beam (applyBiTri -> applyBiTriFullKernel, feature_impl_ngram_partial_kernels.h:19-110)
    const bool scoreTail = remaining > 0 && G > 0;
    const u32 perNode = scoreTail ? (U * nB + remaining * nT) : 1u;
    u32 KCH = sm.wcap / (perNode ? perNode : 1u);
    if (KCH > KCHUNK) KCH = KCHUNK;
    if (KCH < 1) KCH = 1;
    u32 rScan = 0;
    while (rScan < R) {
      u32 K = 0;
      if (JPP_LANE0) {
        u32 r = rScan;
        for (; r < R && K < KCH; ++r) {
          if (bv.selK[n0 + r] != 0xFFFFu) { sm.selNode[K] = r; bv.selK[n0 + r] = (u16)K; ++K; }
        }
        sm.misc[3] = K;
        sm.misc[4] = r;
      }
      JPP_SYNC();
      K = sm.misc[3];
      rScan = sm.misc[4];
      if (K == 0) break;
      if (scoreTail) {
        JPP_LANES
        for (u32 x = (u32)lane; x < K * SP; x += 32)
          sm.t0pat[x] = bv.patterns[(size_t)(n0 + sm.selNode[x / SP]) * SP + x % SP];
        JPP_LANES_END
        const u32 triBase = K * U * nB;
        JPP_LANES
        {  // bigram items (k, u, f), four gathers in flight per lane
          u32 k = 0, u = 0, f = (u32)lane;
          while (f >= nB) { f -= nB; if (++u == U) { u = 0; ++k; } }
          for (u32 it = (u32)lane; it < triBase; it += 32u * JPP_GATHER_BATCH) {
            u32 idx[JPP_GATHER_BATCH];
            float v[JPP_GATHER_BATCH];
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q) {
              idx[q] = 0;
              if (it + 32u * q < triBase) {
                const u64* P = sm.t0pat + (size_t)k * SP;
                const u64* t1 = sm.t1pat + (size_t)u * SP;
                idx[q] = (u32)mix(mix(ft.biState[f], P[ft.biT0[f]]), t1[ft.biT1[f]]) & mask;
              }
              f += 32;
              while (f >= nB) { f -= nB; if (++u == U) { u = 0; ++k; } }
            }
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q) v[q] = JPP_LDG(w + idx[q]);
#pragma unroll
            for (int q = 0; q < JPP_GATHER_BATCH; ++q)
              if (it + 32u * q < triBase) sm.wbuf[it + 32u * q] = v[q];
          }
        }
        {  // trigram items (k, j, f) of the tail elements
          const u32 nTri = K * remaining * nT;
          u32 k = 0, j = 0, f = (u32)lane;
          while (nT && f >= nT) { f -= nT; if (++j == remaining) { j = 0; ++k; } }
          for (u32 it = (u32)lane; it < nTri; it += 32) {
            const u32 gi = full + j;
            const u64* P = sm.t0pat + (size_t)k * SP;
            const u64* t1 = sm.t1pat + (size_t)sm.t1idx[gi] * SP;
            const u64* t2 = sm.t2pat + (size_t)gi * SP;
            const u32 idx = (u32)mix(mix(mix(ft.triState[f], P[ft.triT0[f]]), t1[ft.triT1[f]]), t2[ft.triT2[f]]) & mask;
            sm.wbuf[triBase + it] = JPP_LDG(w + idx);
            f += 32;
            while (f >= nT) { f -= nT; if (++j == remaining) { j = 0; ++k; } }
          }
        }
        JPP_LANES_END
        JPP_LANES
        for (u32 it = (u32)lane; it < K * U; it += 32) {
          const u32 k = it / U, u = it - k * U;
          const float* x = sm.wbuf + (size_t)it * nB;
          float v;
          if (u + 1 < U) {
            float a[2];
            osum<2>(x, nB, a);
            v = a[0] + a[1];
          } else {
            v = osum_unrolled4(x, nB);
          }
          sm.scbuf[k * GMAX + u] = v;
        }
        JPP_LANES_END
        JPP_LANES
        for (u32 it = (u32)lane; it < K * remaining; it += 32) {
          const u32 k = it / remaining, j = it - k * remaining, gi = full + j;
          const u32 node = n0 + sm.selNode[k];
          const float* x = sm.wbuf + triBase + (size_t)it * nT;
          float v;
          if (j + 1 < remaining) {
            float a[2];
            osum<2>(x, nT, a);
            v = sm.scbuf[k * GMAX + sm.t1idx[gi]] + a[0] + a[1];
          } else {
            v = sm.scbuf[k * GMAX + sm.t1idx[gi]] + osum_unrolled4(x, nT);
          }
          v += bv.t0[node];  // copyT0Scores(tail, t0Score), score_processor.cc:411-424
          sm.loc[k * GMAX + gi] = v;
          v += cand_score(sm.gbeam[gi]);
          sm.tot[k * GMAX + gi] = v;
        }
        JPP_LANES_END
        gathers += (unsigned long long)K * ((unsigned long long)U * nB + (unsigned long long)remaining * nT);
      }
      JPP_LANES
      for (u32 it = (u32)lane; it < K * full; it += 32) {
        const u32 k = it / full, i = it - k * full;
        float v = bv.pre[(size_t)i * bv.totalNodes + n0 + sm.selNode[k]];
        v += 0.0f;  // copyT0Scores(head, 0)
        sm.loc[k * GMAX + i] = v;
        v += cand_score(sm.gbeam[i]);
        sm.tot[k * GMAX + i] = v;
      }
      JPP_LANES_END
      JPP_LANES
      for (u32 k = (u32)lane; k < K; k += 32) {
        const u32 node = n0 + sm.selNode[k];
        write_beam(bv, cfg, node, sm.gbeam, (int)G, sm.tot + k * GMAX, sm.loc + k * GMAX);
        if (bnd == nb - 1) {
          for (u32 g = 0; g < G; ++g) {
            bv.eosGbeam[(size_t)sent * GMAX + g] = sm.gbeam[g];
            bv.eosLocal[(size_t)sent * GMAX + g] = sm.loc[k * GMAX + g];
          }
        }
      }
      JPP_LANES_END
    }

    // ---- right nodes outside the cutoff: beam of the head elements only
    if (rightCheck && R > toKeep) {
      JPP_LANES
      for (u32 r = (u32)lane; r < R; r += 32) {
        const u32 node = n0 + r;
        if (bv.selK[node] == 0xFFFFu) {
          float tot[MAX_BEAM], loc[MAX_BEAM];
          for (u32 i = 0; i < full; ++i) {
            float v = bv.pre[(size_t)i * bv.totalNodes + node];
            v += 0.0f;
            loc[i] = v;
            v += cand_score(sm.gbeam[i]);
            tot[i] = v;
          }
          write_beam(bv, cfg, node, sm.gbeam, (int)full, tot, loc);
        }
      }
      JPP_LANES_END
    }
  }
  if (JPP_LANE0 && bv.counters) {
#if defined(__CUDA_ARCH__)
    atomicAdd(bv.counters + 0, gathers);
    atomicAdd(bv.counters + 1, streamBytes);
#else
    bv.counters[0] += gathers;
    bv.counters[1] += streamBytes;
#endif
  }
}

}  // namespace jppb200
