// jpp_b200_kernels.h -- device code of the B200 lattice scorer.
//
// Every stage is a flat data-parallel function of an index (a node, a sentence,
// a selected node); jpp_b200.cu wraps each in a __global__ kernel.  The beam
// walk is organised as a wavefront over the boundary index (see "wavefront
// walk" below): sentences are independent, so at wave b every sentence that
// has a boundary b is processed at once.
//
// The functions are __host__ __device__: tests/emul compiles this very source
// as plain C++ and calls the functions in loops (TEST ONLY -- it lets the kernel
// logic be checked against the oracle on a machine without a GPU; it is never
// part of libjpp_b200.so).
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
#define JPP_LDG(p) __ldg(p)
#else
#define JPP_LDG(p) (*(p))
#endif

namespace jppb200 {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;
typedef uint8_t u8;

// ---------------------------------------------------------------- limits ----
enum {
  MAX_UNI = 64, MAX_BI = 64, MAX_TRI = 16, MAX_PAT = 64, MAX_PRIM = 48, MAX_COMP = 64, MAX_POOL = 256,
  MAX_BEAM = 32    // beam_size, global_beam_size, right_gbeam_size upper bound on the GPU path
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
  u32 dynamicOrder;  // 1: every sum through computeUnrolled4RawPerceptron (the reference's dynamic feature path)
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
  u8 patOfSlot[MAX_PAT];   // inverse of patIndex (0xFF: no pattern writes the slot)
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
  const i32* nodeEntry;      // [nodes][entrySize], or null: entryRows[nodeEntryRow[node]]
  const u32* nodeEntryRow;
  const i32* entryRows;
  u32 numEntryRows;
  const i32* rowPlaceholders;   // compact form, when nodePlaceholders is null
  const i32* rowRnnId;          // compact form, when nodeRnnId is null
  const i32* nodePlaceholders;  // may be null
  const i32* nodeRnnId;         // may be null
  // resident form (nodeEntry and nodeEntryRow null): dictionary rows live in the handle's table,
  // rows of a sentence's extra nodes in entryRows at sentExtraOff[sent] + ~entry_ptr
  const u32* sentExtraOff;
  const i32* dicRowOfPtr;       // [dicPtrLimit] row of EntryPtr::rawValue(), -1 = not resident
  u32 dicPtrLimit;
  const i32* dicRows;           // [rows][entrySize]
  const i32* dicRnnId;          // [rows] or null
  const jpp_b200_node_ptr* ends;
  const jpp_b200_codepoint* cps;
  // derived on upload
  u32* nodeBnd;    // [nodes] global boundary index of the node
  u32* bndSent;    // [boundaries] sentence of the boundary
  u32* endNode;    // [ends] global node index an end entry points at
  u32* sentOrder;  // [sentences] longest first (work queue order)
  const u8* sentStatus;  // [sentences] 0 = ok; a failed sentence is skipped by every kernel (may be null)
  // results
  jpp_b200_beam_element* beam;  // [nodes][B]
  float* beamScores;            // [nodes][B][S]
  u64* patterns;                // stored pattern p of node n at patterns[n * patNodeStride + p * patPStride]
  size_t patNodeStride, patPStride;  // device: (1, nodes) = structure of arrays, so that
                                // adjacent nodes' values of one pattern share a cache line;
                                // the ABI layout [nodes][stored] is (stored, 1)
  float* t0;                    // [nodes]
  // scratch
  float* pre;        // [rcheck][nodes]  T0 prescores
  float* cut;        // [nodes]          cutoff score
  u32* cutIdx;       // [nodes]          index scratch for nth_element emulation
  u32* selPos;       // [nodes]          position in the wave's list of selected nodes / NOT_SELECTED
  float* bi0;        // [nodes]          applyBiTri bigram sum against unique t1 0 (written by W2)
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
  u32 dyn;  // FeatTables::dynamicOrder (the connection-score sum depends on it too)
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

// stored pattern p of a node
struct PatRow {
  const u64* base;
  size_t pstride;
  JPP_HD u64 operator[](u32 p) const { return base[(size_t)p * pstride]; }
};
JPP_HD PatRow pat_row(const BatchView& bv, u32 node) { return PatRow{bv.patterns + (size_t)node * bv.patNodeStride, bv.patPStride}; }

// first boundary and boundary count of a sentence; a failed sentence (jpp_b200_result::sentence_status != 0)
// counts as having no boundaries, which is what makes every stage skip it
JPP_HD u32 sent_span(const BatchView& bv, u32 sent, u32& b0) {
  b0 = bv.sentBndOff[sent];
  if (bv.sentStatus && bv.sentStatus[sent]) return 0u;
  return bv.sentBndOff[sent + 1] - b0;
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

// ------------------------------------------------------ weight loaders ------
// Every perceptron-table read goes through a loader.  The gathered indices are
// hashes, i.e. uniformly spread over the table (256 MiB at k = 26), but their
// FREQUENCIES are extremely skewed (the 16 384 hottest indices are 73 % of all
// gathers on the benchmark corpus: DESIGN.md section 5), and a warp-wide gather of 32
// different lines costs the L1TEX pipe ~64 cycles however well it caches.
// WLoadCached therefore keeps a direct-mapped (index -> weight) cache of the hot
// slice in shared memory: one 8-byte slot {tag = index, value}, filled on a
// miss.  A slot is written with a single 64-bit store, so concurrent fills of
// the same slot by different threads cannot produce a torn pair; the value is a
// pure function of the index, so whoever wins, a hit returns exactly what the
// table holds -- results are bit-identical to WLoadDirect.
// One read of the perceptron table.  The table (256 MB on the bench model) is read at hash-random indices: a
// line brought into L1 is practically never hit again, while everything else the kernels touch (pattern rows,
// beams, item lists) is.  A/B builds: JPP_GATHER_NA = the read does not allocate in L1; JPP_GATHER_L2_LAST = its
// L2 line is marked evict-last (the streamed lattice data, not the table, should leave L2 first).
JPP_HD float gather_weight(const float* __restrict__ w, u32 id) {
#if defined(__CUDA_ARCH__) && (defined(JPP_GATHER_NA) || defined(JPP_GATHER_L2_LAST))
  float v;
#if defined(JPP_GATHER_L2_LAST)
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#if defined(JPP_GATHER_NA)
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(w + id), "l"(pol));
#else
  asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(w + id), "l"(pol));
#endif
#else
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(w + id));
#endif
  return v;
#else
  return JPP_LDG(w + id);
#endif
}

struct WLoadDirect {
  const float* __restrict__ w;
  template <int N>
  JPP_HD void get(const u32 (&id)[N], float (&x)[N]) const {
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = gather_weight(w, id[j]);
  }
  JPP_HD float one(u32 id) const { return gather_weight(w, id); }
};

#if defined(__CUDACC__)
struct WLoadCached {
  const float* __restrict__ w;
  unsigned long long* slots;  // shared memory, [cmask + 1], tag 0xFFFFFFFF = empty
  u32 cmask;
  template <int N>
  __device__ __forceinline__ void get(const u32 (&id)[N], float (&x)[N]) const {
    unsigned long long e[N];
#pragma unroll
    for (int j = 0; j < N; ++j) e[j] = slots[id[j] & cmask];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const bool hit = (u32)e[j] == id[j];
      x[j] = hit ? __uint_as_float((u32)(e[j] >> 32)) : __ldg(w + id[j]);
      if (!hit) slots[id[j] & cmask] = ((unsigned long long)__float_as_uint(x[j]) << 32) | id[j];
    }
  }
  __device__ __forceinline__ float one(u32 id) const {
    const unsigned long long e = slots[id & cmask];
    if ((u32)e == id) return __uint_as_float((u32)(e >> 32));
    const float v = __ldg(w + id);
    slots[id & cmask] = ((unsigned long long)__float_as_uint(v) << 32) | id;
    return v;
  }
};
#endif

// ------------------------------------------------------- entry of a node -----
// EntryPtr::BOS() / EOS() (core_types.h:43-51)
constexpr i32 kEntryBos = (i32)0x80000000u, kEntryEos = (i32)0x80000002u;

struct NodeEntry {
  const i32* entry;  // null: every field = entry_ptr (BOS / EOS)
  const i32* ph;     // null: placeholders are 0
};

// the three batch forms of include/jpp_b200.h: inline rows, compact rows, resident table
JPP_HD NodeEntry node_entry(const FeatTables& ft, const BatchView& bv, u32 node, u32 sent, i32 entryPtr) {
  NodeEntry e;
  e.ph = bv.nodePlaceholders ? bv.nodePlaceholders + (size_t)node * ft.nPlaceholders : nullptr;
  if (bv.nodeEntry) {
    e.entry = bv.nodeEntry + (size_t)node * ft.entrySize;
  } else if (bv.nodeEntryRow) {
    u32 row = bv.nodeEntryRow[node];
    if (row >= bv.numEntryRows) row = 0;  // cannot happen: the upload's validation fails such a sentence
    e.entry = bv.entryRows + (size_t)row * ft.entrySize;
    if (!e.ph && bv.rowPlaceholders) e.ph = bv.rowPlaceholders + (size_t)row * ft.nPlaceholders;
  } else if (entryPtr >= 0) {
    const i32 row = (u32)entryPtr < bv.dicPtrLimit ? bv.dicRowOfPtr[entryPtr] : -1;
    e.entry = bv.dicRows + (size_t)(row < 0 ? 0 : row) * ft.entrySize;  // row < 0: validation has failed the sentence
    e.ph = nullptr;                                                       // providedFeature() is 0 for dictionary entries
  } else if (entryPtr == kEntryBos || entryPtr == kEntryEos) {
    e.entry = nullptr;
    e.ph = nullptr;
  } else {
    const u32 row = bv.sentExtraOff[sent] + (u32)~entryPtr;
    e.entry = bv.entryRows + (size_t)row * ft.entrySize;
    e.ph = bv.rowPlaceholders ? bv.rowPlaceholders + (size_t)row * ft.nPlaceholders : nullptr;
  }
  return e;
}

// RnnIdContainer::resolveId of a node, wherever the batch form keeps it
JPP_HD i32 node_rnn_id(const BatchView& bv, u32 node, u32 sent, i32 entryPtr) {
  if (bv.nodeRnnId) return bv.nodeRnnId[node];
  if (bv.nodeEntry) return 0;
  if (bv.nodeEntryRow) {
    const u32 row = bv.nodeEntryRow[node];
    return bv.rowRnnId ? bv.rowRnnId[row < bv.numEntryRows ? row : 0u] : 0;
  }
  if (entryPtr >= 0) {
    const i32 row = (u32)entryPtr < bv.dicPtrLimit ? bv.dicRowOfPtr[entryPtr] : -1;
    return (bv.dicRnnId && row >= 0) ? bv.dicRnnId[row] : 0;
  }
  if (entryPtr == kEntryBos || entryPtr == kEntryEos) return 0;
  return bv.rowRnnId ? bv.rowRnnId[bv.sentExtraOff[sent] + (u32)~entryPtr] : 0;
}

// ------------------------------------------------------------ T0 kernel -----
// One thread per node: primitive features -> 38 pattern hashes -> 32 unigram
// hashes -> weight gather + ordered sum.  Replaces the generated
// patternsAndUnigramsApply (codegen/pattern_feature_codegen.cc:199-313).
// (kind, a, b, c) = one row of the primitive table; with literal arguments (the compiled-in program of
// jpp_b200_static_jumandic_t0.h) the switch folds away
JPP_HD u64 prim_value_c(i32 kind, i32 a, i32 pb, i32 pc, const i32* entry, const i32* ph, jpp_b200_node_info ni,
                        const jpp_b200_codepoint* cps, u32 ncp) {
  switch (kind) {
    // entry == nullptr: a BOS / EOS row, every field is the entry pointer itself (feature_impl_types.h:130-133)
    case JPP_B200_PRIM_COPY: return (u64)(u32)(entry ? entry[a] : ni.entry_ptr);
    case JPP_B200_PRIM_SHIFT_MASK: return (u64)(((u32)(entry ? entry[a] : ni.entry_ptr) >> pb) & (u32)pc);
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

JPP_HD u64 prim_value(const FeatTables& ft, u32 p, const i32* entry, const i32* ph, jpp_b200_node_info ni,
                      const jpp_b200_codepoint* cps, u32 ncp) {
  return prim_value_c(ft.primKind[p], ft.primA[p], ft.primB[p], ft.primC[p], entry, ph, ni, cps, ncp);
}

// hash of pattern p over the primitive values (pattern_feature_codegen.cc:199-313)
// primitive i lives at prim[i * ps] (ps = 1: a private array; ps = block size: the block's shared-memory slab)
JPP_HD u64 pattern_value(const FeatTables& ft, u32 p, const u64* prim, u32 ps = 1) {
  u64 h = ft.patState[p];
  const u32 off = ft.patRefOff[p];
  for (u32 k = 0; k < ft.patRefLen[p]; ++k) {
    const u32 c = ft.patRefPool[off + k];
    const u32 tl = ft.compTrueLen[c], fl = ft.compFalseLen[c];
    if (tl == 0 && fl == 0) {
      h = mix(h, prim[ft.compPrim[c] * ps]);
    } else if (prim[ft.compPrim[c] * ps] != 0) {
      for (u32 t = 0; t < tl; ++t) h = mix(h, prim[ft.branchPool[ft.compTrueOff[c] + t] * ps]);
    } else {
      for (u32 t = 0; t < fl; ++t) h = mix(h, prim[ft.branchPool[ft.compFalseOff[c] + t] * ps]);
    }
  }
  return h;
}

// LEAN: a pattern is hashed where it is consumed (stored row, unigram feature) instead of
// keeping all of them in a per-thread array.  With 2 048 threads per SM that array alone was
// 0.6 MB of local memory per SM, evicted through L2 to HBM: ncu showed 28 GB written by a kernel
// whose output is 5 GB (profiles/r2_ncu_summary.md).  The primitive values (8 B x nPrim) stay.
template <typename WL, bool LEAN = true>
JPP_HD void t0_node(const FeatTables& ft, const BatchView& bv, const WL& wl, u32 node, u64* primStore = nullptr, u32 ps = 1) {
  const u32 gb = bv.nodeBnd[node];
  if (gb == 0xFFFFFFFFu) return;  // node of no valid boundary (failed sentence)
  const u32 sent = bv.bndSent[gb];
  const u32 localBnd = gb - bv.sentBndOff[sent];
  const u32 SP = ft.nStored;
  u64* pout = bv.patterns + (size_t)node * bv.patNodeStride;
  const size_t pstr = bv.patPStride;
  // (nodes of a sentence that failed validation never get here: their nodeBnd stays 0xFFFFFFFF)
  const jpp_b200_node_info ni = bv.nodeInfo[node];
  if (localBnd < 2) {  // BOS rows, lattice_builder.cc:171-176
    for (u32 p = 0; p < SP; ++p) pout[p * pstr] = (u64)(u32)ni.entry_ptr;
    bv.t0[node] = 0.0f;
    return;
  }
  const u32 cp0 = bv.sentCpOff[sent];
  const u32 ncp = bv.sentCpOff[sent + 1] - cp0;
  const jpp_b200_codepoint* cps = bv.cps + cp0;
  const NodeEntry ne = node_entry(ft, bv, node, sent, ni.entry_ptr);
  const i32* entry = ne.entry;
  const i32* ph = ne.ph;
  // The primitive values are indexed by the feature program (dynamically): as a private array they are
  // local memory -- ncu r2e: 1.49 G local-load sectors (L1 hits), more than the kernel's 1.31 G sectors of
  // weight gathers.  JPP_T0_PRIM_SMEM (A/B build) keeps them in a slab of shared memory the kernel passes
  // ([primitive][thread]: conflict-free): measured SLOWER, 13.7 -> 18.2 ms (profiles/README.md r2h) -- 8 B x
  // nPrim x 128 threads of shared memory per block cost more occupancy than the local loads cost L1 time.
#if defined(__CUDA_ARCH__) && defined(JPP_T0_PRIM_SMEM)
  u64* prim = primStore;
#else
  u64 primLocal[MAX_PRIM];
  u64* prim = primLocal;
  ps = 1;
  (void)primStore;
#endif
  for (u32 i = 0; i < ft.nPrim; ++i) prim[i * ps] = prim_value(ft, i, entry, ph, ni, cps, ncp);
  u64 pat[LEAN ? 1 : MAX_PAT];
  auto slot_value = [&](u32 slot) -> u64 {
    if (!LEAN) return pat[slot];
    const u32 p = ft.patOfSlot[slot];
    return p == 0xFFu ? 0ull : pattern_value(ft, p, prim, ps);
  };
  if (!LEAN)
    for (u32 p = 0; p < ft.nPat; ++p) pat[ft.patIndex[p]] = pattern_value(ft, p, prim, ps);
#if defined(__CUDA_ARCH__) && defined(JPP_T0_STORE16)
  if (pstr == 1 && (SP & 1u) == 0) {
    // the row is 16-byte aligned (SP even, 8-byte slots): half as many stores, whole 16-byte pieces of a sector
    for (u32 p = 0; p < SP; p += 2) {
      ulonglong2 v;
      v.x = slot_value(p);
      v.y = slot_value(p + 1);
      *reinterpret_cast<ulonglong2*>(pout + p) = v;
    }
  } else
#endif
  for (u32 p = 0; p < SP; ++p) pout[p * pstr] = slot_value(p);
  // unigram score.  Rows 0..R-2 of a boundary: "float score_part_j = w" for the
  // first uniVars terms then "+=" (pattern_feature_codegen.cc:229-236), summed
  // p0 + p1 + ...; the last row goes through computeUnrolled4RawPerceptron
  // (perceptron.h:46-72): four accumulators started at 0.
  const bool lastRow = ft.dynamicOrder || (node + 1 == bv.bndNodeOff[gb + 1]);
  const u32 nv = lastRow ? 4u : ft.uniVars;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const u32 nU = ft.nUni, mask = ft.mask;
  constexpr u32 T0_LB = 4;  // table reads in flight per thread (8 measured 0.5 ms slower here: 32 registers, local arrays)
  for (u32 u0 = 0; u0 < nU; u0 += T0_LB) {
    float wv[T0_LB];
    u32 id[T0_LB];
#pragma unroll
    for (u32 j = 0; j < T0_LB; ++j) {
      const u32 u = u0 + j < nU ? u0 + j : nU - 1;
      id[j] = (u32)mix(ft.uniState[u], slot_value(ft.uniT0[u])) & mask;
    }
    wl.get(id, wv);
#pragma unroll
    for (u32 j = 0; j < T0_LB; ++j) {
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

// ------------------------------------------------------- wavefront walk -----
// The beams of boundary b depend on every earlier boundary of the SAME
// sentence (reference analyzer_impl.cc:260-284) but sentences are independent.
// So all sentences advance one boundary index at a time ("wave" b = local
// boundary b of every sentence that has one), and inside a wave every phase is
// a flat data-parallel kernel:
//   W1 wave_gbeam      thread per sentence      makeGlobalBeam, dedupT1, gatherT2 ids
//   W2 wave_prescore   thread per right node    T0 + bigram + trigram vs the head elements
//   W3 wave_cutoff     thread per sentence      makeT0cutoffBeam -> list of selected nodes
//   W4 wave_item_*     thread per (selected node, unique t1) / (selected node, gbeam element)
//                                               bigram / trigram part of applyBiTri
//   W5 wave_final_node thread per right node    connection scores, makeT0Beam, beam row
// Threads of one warp run the same feature loop in lockstep, so the feature
// tables (kernel parameter -> constant bank) are broadcast reads, the weight
// gathers of a warp are 32 independent requests, and no lane idles because a
// boundary has few right nodes: the parallelism is across sentences.

// per-sentence state of the current wave (device arrays)
struct WaveView {
  u64* gbKey;      // [NS][G]  packed BeamCandidate of the global beam, descending
  u32* gbLnode;    // [NS][G]  left node of each element
  u32* gbT2node;   // [NS][G]  node before the left node
  u8* gbT1idx;     // [NS][G]  index into the unique-t1 list
  u32* t1node;     // [NS][G]  unique left nodes, first-seen order
  u8* gCount;      // [NS]
  u8* uCount;      // [NS]
  u32* selCount;   // [3]      [0] selected nodes, [1] bigram items, [2] trigram items
  struct WaveItem* biItems;   // [cap*G]
  struct WaveItem* triItems;  // [cap*G]
  float* scb;      // [cap*G]     bigram part of applyBiTri per (selected node, unique t1)
  float* tri;      // [cap*G][2]  trigram accumulators per (selected node, global beam element)
  const u32* waveNodes;    // [NN] node ids grouped by local boundary index
  const u32* waveNodeOff;  // [maxNb+1]
};

#if defined(__CUDACC__)
#define JPP_NOINLINE __noinline__
#else
#define JPP_NOINLINE
#endif

// Ordered sum of n gathered weights with NV interleaved accumulators started at
// 0 ("f_j = 0; f_j += w", partial_ngram_feature_codegen.cc:44-66; NV = 4 is also
// computeUnrolled4RawPerceptron, perceptron.h:46-72; NV = 2 is the r1/r2 pair
// of applyBiTriFullKernel, feature_impl_ngram_partial_kernels.h:35-61).
// idxOf(f) yields the table index; NV loads are in flight at a time.
// idxOf(f) yields the table index; LB (a multiple of NV) independent loads are
// issued before their values are added, in feature order, into accumulator f % NV
// -- the order of additions inside each accumulator is unchanged.
template <int NV, int LB, typename WL, typename IdxFn>
JPP_HD void gsum(const WL& wl, u32 n, IdxFn idxOf, float* acc) {
#pragma unroll
  for (int j = 0; j < NV; ++j) acc[j] = 0.0f;
  u32 f = 0;
  for (; f + LB <= n; f += LB) {
    u32 id[LB];
    float x[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) id[j] = idxOf(f + j);
    wl.get(id, x);
#pragma unroll
    for (int j = 0; j < LB; ++j) acc[j % NV] += x[j];
  }
  for (; f + NV <= n; f += NV) {
    u32 id[NV];
    float x[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) id[j] = idxOf(f + j);
    wl.get(id, x);
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] += x[j];
  }
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (f + j < n) acc[j] += wl.one(idxOf(f + j));
}

template <typename WL, typename IdxFn>
JPP_HD float gsum_vars(const WL& wl, u32 n, u32 nv, IdxFn idxOf) {
  float a[8];
  switch (nv) {
    case 8: gsum<8, 8>(wl, n, idxOf, a); return ((((((a[0] + a[1]) + a[2]) + a[3]) + a[4]) + a[5]) + a[6]) + a[7];
    case 4: gsum<4, 8>(wl, n, idxOf, a); return ((a[0] + a[1]) + a[2]) + a[3];
    case 2: gsum<2, 8>(wl, n, idxOf, a); return a[0] + a[1];
    default: gsum<1, 8>(wl, n, idxOf, a); return a[0];
  }
}

// One row of a generated n-gram step: rows 0..R-2 use NV interleaved accumulators, the last
// row goes through computeUnrolled4RawPerceptron (4 accumulators).  Neighbouring threads
// hold rows of both kinds, so both accumulator sets are fed from the same gathered values
// (the additions are free next to the gathers) and the result is selected at the end: no
// divergent gather loops.  Per accumulator the order of additions is that of gsum<NV> /
// gsum<4>, so the result is bit-identical to gsum_vars / gsum_unrolled4.
template <int NV, int LB, typename WL, typename IdxFn>
JPP_HD void gsum_dual(const WL& wl, u32 n, IdxFn idxOf, float (&a)[NV], float& unrolled4) {
  static_assert(LB % 4 == 0 && LB % NV == 0, "batch must cover whole accumulator rounds");
  float b[4];
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 0.0f;
  u32 f = 0;
  for (; f + LB <= n; f += LB) {
    u32 id[LB];
    float x[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) id[j] = idxOf(f + j);
    wl.get(id, x);
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      a[j % NV] += x[j];
      if (NV != 4) b[j % 4] += x[j];
    }
  }
  if (f < n) {  // f is a multiple of LB (hence of NV and of 4): accumulator of feature f + j is j % NV
    u32 id[LB];
    float x[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) id[j] = idxOf(f + j < n ? f + j : n - 1);
    wl.get(id, x);
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      if (f + j < n) {
        a[j % NV] += x[j];
        if (NV != 4) b[j % 4] += x[j];
      }
    }
  }
  if (NV == 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = a[j % NV];
  }
  unrolled4 = ((b[0] + b[1]) + b[2]) + b[3];
}

template <int NV, typename WL, typename IdxFn>
JPP_HD float gsum_row_nv(const WL& wl, u32 n, IdxFn idxOf, bool last) {
  float a[NV], s4;
  gsum_dual<NV, 8>(wl, n, idxOf, a, s4);
  float sa = a[0];
#pragma unroll
  for (int j = 1; j < NV; ++j) sa += a[j];
  return last ? s4 : sa;
}

template <typename WL, typename IdxFn>
JPP_HD float gsum_row(const WL& wl, u32 n, u32 nv, IdxFn idxOf, bool last) {
  switch (nv) {
    case 8: return gsum_row_nv<8>(wl, n, idxOf, last);
    case 4: return gsum_row_nv<4>(wl, n, idxOf, last);
    case 2: return gsum_row_nv<2>(wl, n, idxOf, last);
    default: return gsum_row_nv<1>(wl, n, idxOf, last);
  }
}

template <typename WL, typename IdxFn>
JPP_HD float gsum_unrolled4(const WL& wl, u32 n, IdxFn idxOf) {
  float a[4];
  gsum<4, 8>(wl, n, idxOf, a);
  return ((a[0] + a[1]) + a[2]) + a[3];
}

// Writes the beam row of one right node from `n` candidate totals.
JPP_HD void write_beam(const BatchView& bv, const ScoreCfg& cfg, u32 node, const u64* gbeam, int n,
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

// W0: BOS beams of every sentence (analyzer_impl.cc:185-192) and the degenerate
// "no codepoints" lattice (:255-258)
JPP_HD void wave_init_sentence(const BatchView& bv, const ScoreCfg& cfg, u32 sent) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  const u32 B = (u32)cfg.beam;
  for (u32 bb = 0; bb < 2 && bb < nb; ++bb) {
    const u32 n0 = bv.bndNodeOff[b0 + bb];
    for (u32 k = 0; k < B; ++k) {
      jpp_b200_beam_element e;
      if (k == 0) { e.left = 0; e.prev_beam = 0; e.total_score = 0.0f; }
      else { e.left = 0xFFFFu; e.prev_beam = 0xFFFFu; e.total_score = -3.402823466e+38f; }  // EntryBeam::fake()
      bv.beam[(size_t)n0 * B + k] = e;
      for (u32 q = 0; q < cfg.S; ++q) bv.beamScores[((size_t)n0 * B + k) * cfg.S + q] = 0.0f;
    }
  }
  if (nb == 3) {
    const u32 n0 = bv.bndNodeOff[b0 + 2];
    for (u32 k = 0; k < B; ++k) {
      bv.beam[(size_t)n0 * B + k] = fake_elem();
      for (u32 q = 0; q < cfg.S; ++q) bv.beamScores[((size_t)n0 * B + k) * cfg.S + q] = 0.0f;
    }
    bv.gbCount[b0 + 2] = 0;
  }
}

// W1: makeGlobalBeam (score_processor.cc:246-282) + dedupT1 / gatherT2 ids (:363-409).
// One thread scans the L x B left beam elements and keeps the G largest packed
// keys (distinct by construction: score_processor.h:81-115) in sorted order.
JPP_HD void wave_gbeam(const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, u32 sent, u32 bnd) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  if (bnd >= nb || nb <= 3) return;
  const u32 gB = b0 + bnd;
  const u32 R = bv.bndNodeOff[gB + 1] - bv.bndNodeOff[gB];
  const u32 B = (u32)cfg.beam, GMAX = (u32)cfg.gbeam;
  wv.gCount[sent] = 0;
  wv.uCount[sent] = 0;
  if (R == 0) {  // analyzer_impl.cc:263-266
    bv.gbCount[gB] = 0;
    return;
  }
  const u32 e0 = bv.bndEndOff[gB];
  const u32 L = bv.bndEndOff[gB + 1] - e0;
  u64 top[MAX_BEAM];
  u32 G = 0;
  // beams are filled front to back (score_processor.cc:257-260), so fake slots
  // can simply be skipped: every load below is independent of the others
  for (u32 left0 = 0; left0 < L; left0 += 4) {
    u32 ln[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ln[q] = bv.endNode[e0 + (left0 + q < L ? left0 + q : left0)];
    for (u32 slot = 0; slot < B; ++slot) {
      jpp_b200_beam_element el[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) el[q] = bv.beam[(size_t)ln[q] * B + slot];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (left0 + q >= L || is_fake(el[q])) continue;
        const u64 key = pack_cand(el[q].total_score, left0 + q, slot);
        if (G == GMAX && key < top[G - 1]) continue;
        u32 pos = G < GMAX ? G : GMAX - 1;
        while (pos > 0 && top[pos - 1] < key) { top[pos] = top[pos - 1]; --pos; }
        top[pos] = key;
        if (G < GMAX) ++G;
      }
    }
  }
  u64* gbKey = wv.gbKey + (size_t)sent * GMAX;
  u32* gbLnode = wv.gbLnode + (size_t)sent * GMAX;
  u32* gbT2 = wv.gbT2node + (size_t)sent * GMAX;
  u8* t1idx = wv.gbT1idx + (size_t)sent * GMAX;
  u32* t1node = wv.t1node + (size_t)sent * GMAX;
  u32 U = 0;
  for (u32 g = 0; g < G; ++g) {
    const u64 key = top[g];
    const u32 left = (u32)(key >> 16) & 0xFFFFu, slot = (u32)key & 0xFFFFu;
    const u32 lnode = bv.endNode[e0 + left];
    const u32 lbnd = bv.ends[e0 + left].boundary;
    const jpp_b200_beam_element el = bv.beam[(size_t)lnode * B + slot];
    gbKey[g] = key;
    gbLnode[g] = lnode;
    gbT2[g] = bv.endNode[bv.bndEndOff[b0 + lbnd] + el.left];
    u32 u = 0;
    for (; u < U; ++u) if (t1node[u] == lnode) break;
    if (u == U) { t1node[U] = lnode; ++U; }
    t1idx[g] = (u8)u;
    bv.gbNode[(size_t)gB * GMAX + g] = lnode;
    bv.gbSlot[(size_t)gB * GMAX + g] = (u16)slot;
    bv.gbLeftBnd[(size_t)gB * GMAX + g] = (u16)lbnd;
  }
  bv.gbCount[gB] = (u8)G;
  wv.gCount[sent] = (u8)G;
  wv.uCount[sent] = (u8)U;
}

// Item of the W4 kernel: one (selected right node, left context) pair, ids resolved by W3 so
// that the item thread starts gathering after a single 16-byte load.
//   bigram item:  a = right node, b = unique t1 node,            slot = pos * G + u
//   trigram item: a = right node, b = t1 node,  c = t2 node,     slot = pos * G + g
// bit 31 of `slot` = the row is the last one of its loop (goes through Unrolled4).
struct alignas(16) WaveItem {
  u32 a, b, c, slot;
};
constexpr u32 ITEM_LAST = 0x80000000u;
constexpr u32 NOT_SELECTED = 0xFFFFFFFFu;

// first unique-t1 index that needs a bigram item: with the right check on, W2 has already
// gathered the bigram weights of (node, t1 of the head element) -- unique t1 number 0
JPP_HD u32 first_bigram_item(const ScoreCfg& cfg) { return cfg.rcheck > 0 ? 1u : 0u; }

// Three sums over ONE pass of gathers: NV accumulators (generated applyBiStep2 rows),
// the r1/r2 pair of applyBiTriFullKernel, and computeUnrolled4RawPerceptron.
template <int NV, typename WL, typename IdxFn>
JPP_HD void gsum_triple(const WL& wl, u32 n, IdxFn idxOf, float& sumNV, float& sumPair, float& sum4) {
  constexpr int LB = 8;
  float a[NV], c[2], b[4];
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = 0.0f;
  c[0] = c[1] = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 0.0f;
  for (u32 f = 0; f < n; f += LB) {
    u32 id[LB];
    float x[LB];
#pragma unroll
    for (int j = 0; j < LB; ++j) id[j] = idxOf(f + j < n ? f + j : n - 1);
    wl.get(id, x);
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      if (f + j < n) {  // f is a multiple of 8: accumulator of feature f + j is j % NV / j % 2 / j % 4
        a[j % NV] += x[j];
        c[j % 2] += x[j];
        b[j % 4] += x[j];
      }
    }
  }
  sumNV = a[0];
#pragma unroll
  for (int j = 1; j < NV; ++j) sumNV += a[j];
  sumPair = c[0] + c[1];
  sum4 = ((b[0] + b[1]) + b[2]) + b[3];
}

// W2: computeT0Prescores (score_processor.cc:497-511) for one right node:
// T0 + bigrams against t1 of head element i + trigrams (generated applyBiStep2 /
// applyTriStep3: biVars / triVars accumulators, last row of the boundary through
// computeUnrolled4RawPerceptron; partial_ngram_feature_codegen.cc:13-153).
// The bigram weights of (node, t1 of head element 0) are the ones applyBiTri needs again for
// unique t1 number 0, only summed in another order: that sum is produced from the same
// gathers and kept in bi0[node], so W4 has no item for it.
template <typename WL>
JPP_HD void wave_prescore(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv,
                          const WL& wl, u32 node) {
  const u32 gB = bv.nodeBnd[node];
  const u32 sent = bv.bndSent[gB];
  const u32 G = wv.gCount[sent], U = wv.uCount[sent];
  const u32 theMax = (u32)cfg.rcheck < G ? (u32)cfg.rcheck : G;
  if (theMax == 0) return;
  const u32 GMAX = (u32)cfg.gbeam, mask = ft.mask;
  const bool lastRow = ft.dynamicOrder || (node + 1 == bv.bndNodeOff[gB + 1]);
  const PatRow P = pat_row(bv, node);
  const float t0 = bv.t0[node];
  for (u32 i = 0; i < theMax; ++i) {
    const PatRow t1 = pat_row(bv, wv.gbLnode[(size_t)sent * GMAX + i]);
    const PatRow t2 = pat_row(bv, wv.gbT2node[(size_t)sent * GMAX + i]);
    auto biIdx = [&](u32 f) -> u32 { return (u32)mix(mix(ft.biState[f], P[ft.biT0[f]]), t1[ft.biT1[f]]) & mask; };
    auto triIdx = [&](u32 f) -> u32 {
      return (u32)mix(mix(mix(ft.triState[f], P[ft.triT0[f]]), t1[ft.triT1[f]]), t2[ft.triT2[f]]) & mask;
    };
    float sc = t0;
    if (i == 0) {
      float sNV, sPair, s4;
      switch (ft.biVars) {
        case 8: gsum_triple<8>(wl, ft.nBi, biIdx, sNV, sPair, s4); break;
        case 4: gsum_triple<4>(wl, ft.nBi, biIdx, sNV, sPair, s4); break;
        case 2: gsum_triple<2>(wl, ft.nBi, biIdx, sNV, sPair, s4); break;
        default: gsum_triple<1>(wl, ft.nBi, biIdx, sNV, sPair, s4); break;
      }
      sc += lastRow ? s4 : sNV;
      bv.bi0[node] = (U == 1 || ft.dynamicOrder) ? s4 : sPair;  // applyBiTriFullKernel row of unique t1 0
    } else {
      sc += gsum_row(wl, ft.nBi, ft.biVars, biIdx, lastRow);
    }
    sc += gsum_row(wl, ft.nTri, ft.triVars, triIdx, lastRow);
    bv.pre[(size_t)i * bv.totalNodes + node] = sc;
  }
}

// Sizes of one boundary step (fullBeamApplySize, score_processor.cc:299-301, and makeT0cutoffBeam)
struct StepSizes {
  u32 full, toKeep, remaining, theMax;
};
JPP_HD StepSizes step_sizes(const ScoreCfg& cfg, u32 R, u32 G) {
  StepSizes z;
  z.full = 0; z.toKeep = R; z.remaining = G; z.theMax = 0;
  if (cfg.rcheck > 0) {
    z.full = (u32)cfg.rcheck;
    if (R < z.full) z.full = R;
    if (G < z.full) z.full = G;
    z.toKeep = (u32)cfg.rbeam < R ? (u32)cfg.rbeam : R;
    z.remaining = G - z.full;
    z.theMax = (u32)cfg.rcheck < G ? (u32)cfg.rcheck : G;
  }
  return z;
}

// work counters of one boundary step (DESIGN.md "roofline accounting"): `gathers` is the
// reference algorithm's figure, `issued` what the kernels really read
JPP_HD void step_counters(const FeatTables& ft, const ScoreCfg& cfg, const StepSizes& z, u32 R, u32 L, u32 G, u32 U, u32 nsel,
                          unsigned long long& gathers, unsigned long long& streamBytes, unsigned long long& issued) {
  const bool scoreTail = z.remaining > 0 && G > 0;
  gathers = (unsigned long long)z.theMax * R * (ft.nBi + ft.nTri);
  issued = gathers;
  if (scoreTail) {
    const u32 biRows = U - (first_bigram_item(cfg) < U ? first_bigram_item(cfg) : U);
    gathers += (unsigned long long)nsel * ((unsigned long long)U * ft.nBi + (unsigned long long)z.remaining * ft.nTri);
    issued += (unsigned long long)nsel * ((unsigned long long)biRows * ft.nBi + (unsigned long long)z.remaining * ft.nTri);
  }
  streamBytes = 8ull * L * (u32)cfg.beam + 8ull * ft.nStored * ((unsigned long long)R + U + G) + 4ull * L +
                (8ull + 4ull * cfg.S) * R * (u32)cfg.beam;
}

// items of one selected node (position `pos` in the wave's list)
JPP_HD void emit_items(const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, u32 sent, const StepSizes& z, u32 G, u32 U,
                       u32 node, u32 pos, WaveItem* bi, WaveItem* tri) {
  const u32 GMAX = (u32)cfg.gbeam;
  const u32* t1node = wv.t1node + (size_t)sent * GMAX;
  const u8* t1idx = wv.gbT1idx + (size_t)sent * GMAX;
  const u32* gbT2 = wv.gbT2node + (size_t)sent * GMAX;
  const u32 u0 = first_bigram_item(cfg);
  for (u32 u = u0; u < U; ++u) {
    WaveItem it;
    it.a = node; it.b = t1node[u]; it.c = 0;
    it.slot = (pos * GMAX + u) | (u + 1 == U ? ITEM_LAST : 0u);
    bi[u - u0] = it;
  }
  for (u32 g = z.full; g < G; ++g) {
    WaveItem it;
    it.a = node; it.b = t1node[t1idx[g]]; it.c = gbT2[g];
    it.slot = (pos * GMAX + g) | (g + 1 == G ? ITEM_LAST : 0u);
    tri[g - z.full] = it;
  }
  (void)bv;
}

// W3 (serial statement; the GPU runs k_w_cutoff_group, 8 lanes per sentence, which must
// produce the same selection): makeT0cutoffBeam (score_processor.cc:471-495) -- exact
// std::nth_element replay -- the work counters of the boundary, the list of selected nodes
// and their W4 items.
JPP_HD void wave_cutoff(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, u32 sent, u32 bnd) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  if (bnd >= nb || nb <= 3) return;
  const u32 gB = b0 + bnd;
  const u32 n0 = bv.bndNodeOff[gB];
  const u32 R = bv.bndNodeOff[gB + 1] - n0;
  if (R == 0) return;
  const u32 G = wv.gCount[sent], U = wv.uCount[sent];
  const u32 L = bv.bndEndOff[gB + 1] - bv.bndEndOff[gB];
  const StepSizes z = step_sizes(cfg, R, G);
  if (cfg.rcheck > 0 && R > z.toKeep) {
    u32* idx = bv.cutIdx + n0;
    float* cut = bv.cut + n0;
    for (u32 r = 0; r < R; ++r) {
      float v = 0;
      for (u32 j = 0; j < z.full; ++j) v += bv.pre[(size_t)j * bv.totalNodes + n0 + r];
      cut[r] = v;
      idx[r] = r;
      bv.selPos[n0 + r] = NOT_SELECTED;
    }
    em_std_nth_element(idx, (long)z.toKeep, (long)R, ScoreCmp{cut});
    for (u32 k = 0; k < z.toKeep; ++k) bv.selPos[n0 + idx[k]] = 0;
  } else {
    for (u32 r = 0; r < R; ++r) bv.selPos[n0 + r] = 0;
  }
  u32 nsel = 0;
  for (u32 r = 0; r < R; ++r) nsel += (bv.selPos[n0 + r] != NOT_SELECTED);
  const bool scoreTail = z.remaining > 0 && G > 0;
  const u32 u0 = first_bigram_item(cfg);
  const u32 biPer = scoreTail && U > u0 ? U - u0 : 0u;
  const u32 triPer = scoreTail ? z.remaining : 0u;
  unsigned long long g = 0, sb = 0, gi = 0;
  step_counters(ft, cfg, z, R, L, G, U, nsel, g, sb, gi);
  if (bv.counters) { bv.counters[0] += g; bv.counters[1] += sb; bv.counters[2] += gi; }
  const u32 base = wv.selCount[0], biBase = wv.selCount[1], triBase = wv.selCount[2];
  wv.selCount[0] += nsel;
  wv.selCount[1] += nsel * biPer;
  wv.selCount[2] += nsel * triPer;
  for (u32 r = 0, k = 0; r < R; ++r) {
    if (bv.selPos[n0 + r] == NOT_SELECTED) continue;
    const u32 pos = base + k;
    bv.selPos[n0 + r] = pos;
    if (scoreTail)
      emit_items(bv, cfg, wv, sent, z, G, U, n0 + r, pos, wv.biItems + biBase + (size_t)k * biPer,
                 wv.triItems + triBase + (size_t)k * triPer);
    ++k;
  }
}

// W4, bigram item: applyBiTriFullKernel (feature_impl_ngram_partial_kernels.h:19-110) row of one
// (selected right node, unique t1): the r1/r2 accumulator pair, the last t1 row through
// Unrolled4.  Every item is exactly nBi gathers, so the lanes of a warp stay converged
// whatever the number of unique t1 nodes of their sentences.
template <typename WL>
JPP_HD void wave_item_bigram(const FeatTables& ft, const BatchView& bv, const WaveView& wv, const WL& wl, const WaveItem& it) {
  const u32 mask = ft.mask;
  const PatRow P = pat_row(bv, it.a);
  const PatRow t1 = pat_row(bv, it.b);
  auto biIdx = [&](u32 f) -> u32 { return (u32)mix(mix(ft.biState[f], P[ft.biT0[f]]), t1[ft.biT1[f]]) & mask; };
  float a[2], s4;
  gsum_dual<2, 8>(wl, ft.nBi, biIdx, a, s4);
  wv.scb[it.slot & ~ITEM_LAST] = ((it.slot & ITEM_LAST) || ft.dynamicOrder) ? s4 : a[0] + a[1];
}

// W4, trigram item: the trigram part of the same kernel for one (selected right node, global
// beam element); the two accumulators are kept apart because the reference adds them to the
// bigram sum one after the other (scbuf + r1 + r2).
template <typename WL>
JPP_HD void wave_item_trigram(const FeatTables& ft, const BatchView& bv, const WaveView& wv, const WL& wl, const WaveItem& it) {
  const u32 mask = ft.mask;
  const PatRow P = pat_row(bv, it.a);
  const PatRow t1 = pat_row(bv, it.b);
  const PatRow t2 = pat_row(bv, it.c);
  auto triIdx = [&](u32 f) -> u32 {
    return (u32)mix(mix(mix(ft.triState[f], P[ft.triT0[f]]), t1[ft.triT1[f]]), t2[ft.triT2[f]]) & mask;
  };
  float a[2], s4;
  gsum_dual<2, 4>(wl, ft.nTri, triIdx, a, s4);
  float* out = wv.tri + 2 * (size_t)(it.slot & ~ITEM_LAST);
  if ((it.slot & ITEM_LAST) || ft.dynamicOrder) { out[0] = s4; out[1] = 0.0f; }
  else { out[0] = a[0]; out[1] = a[1]; }
}

// connection score of (selected node at `pos`, global beam element gi): head elements come
// from the prescores, the others from the W4 sums; copyT0Scores (score_processor.cc:411-424)
JPP_HD float step_local_score(const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, const StepSizes& z, u32 G, u32 node,
                              u32 pos, u32 gi, const u8* t1idx, float t0) {
  const u32 GMAX = (u32)cfg.gbeam;
  if (gi < z.full) {
    float v = bv.pre[(size_t)gi * bv.totalNodes + node];
    v += 0.0f;  // copyT0Scores(head, 0)
    return v;
  }
  const u32 ti = t1idx[gi];
  const float sc = (ti < first_bigram_item(cfg)) ? bv.bi0[node] : wv.scb[(size_t)pos * GMAX + ti];
  const float* tr = wv.tri + 2 * ((size_t)pos * GMAX + gi);
  float v = (gi + 1 == G || cfg.dyn) ? sc + tr[0] : sc + tr[0] + tr[1];
  v += t0;  // copyT0Scores(tail, t0Score)
  return v;
}

// With at most 16 candidates and no partition step makeT0Beam's std::sort is libstdc++'s
// insertion sort, i.e. a stable descending sort, so for global beams of up to 8 elements
// (G <= beam*4/3; the CLI default is 6 / 5) the slot of candidate i is the number of
// candidates that are larger or equal-and-earlier: everything stays in registers
// (tests/test_sort_emulation.py checks this equivalence against std::sort).
constexpr int SMALL_G = 8;
JPP_HD bool small_gbeam_ok(const ScoreCfg& cfg) {
  return cfg.gbeam <= SMALL_G && (long)cfg.gbeam <= (long)cfg.beam * 4 / 3;
}

// W5: beam row of one right node of the wave (makeT0Beam, score_processor.cc:426-469).
// Selected nodes see the whole global beam; the others keep a beam of the head elements only
// (:336-346).
template <bool SMALL>
JPP_HD void wave_final_node(const BatchView& bv, const ScoreCfg& cfg, const WaveView& wv, u32 node) {
  const u32 gB = bv.nodeBnd[node];
  const u32 sent = bv.bndSent[gB];
  const u32 G = wv.gCount[sent];
  const u32 GMAX = (u32)cfg.gbeam, B = (u32)cfg.beam;
  const u32 R = bv.bndNodeOff[gB + 1] - bv.bndNodeOff[gB];
  const StepSizes z = step_sizes(cfg, R, G);
  const u32 pos = bv.selPos[node];
  const u32 n = (pos != NOT_SELECTED) ? G : z.full;  // candidates of this node
  const u64* gbKey = wv.gbKey + (size_t)sent * GMAX;
  const u8* t1idx = wv.gbT1idx + (size_t)sent * GMAX;
  const float t0 = bv.t0[node];
  const bool eos = (pos != NOT_SELECTED) && gB + 1 == bv.sentBndOff[sent + 1];
  if (SMALL) {
    float tot[SMALL_G], loc[SMALL_G];
    u64 key[SMALL_G];
#pragma unroll
    for (int gi = 0; gi < SMALL_G; ++gi) {
      tot[gi] = 0.0f; loc[gi] = 0.0f; key[gi] = 0;
      if ((u32)gi >= n) continue;
      key[gi] = gbKey[gi];
      loc[gi] = step_local_score(bv, cfg, wv, z, G, node, pos, (u32)gi, t1idx, t0);
      tot[gi] = loc[gi] + cand_score(key[gi]);
    }
    jpp_b200_beam_element* row = bv.beam + (size_t)node * B;
    float* srow = bv.beamScores + (size_t)node * B * cfg.S;
#pragma unroll
    for (int i = 0; i < SMALL_G; ++i) {
      if ((u32)i >= n) continue;
      u32 rank = 0;
#pragma unroll
      for (int j = 0; j < SMALL_G; ++j) {
        if ((u32)j >= n || j == i) continue;
        rank += ((tot[j] > tot[i]) || (j < i && !(tot[i] > tot[j]))) ? 1u : 0u;
      }
      if (rank < B) {
        jpp_b200_beam_element e;
        e.left = (u16)(key[i] >> 16);
        e.prev_beam = (u16)key[i];
        e.total_score = tot[i];
        row[rank] = e;
        srow[(size_t)rank * cfg.S] = loc[i];
        for (u32 q = 1; q < cfg.S; ++q) srow[(size_t)rank * cfg.S + q] = 0.0f;
      }
    }
    for (u32 k = n; k < B; ++k) {
      row[k] = fake_elem();
      for (u32 q = 0; q < cfg.S; ++q) srow[(size_t)k * cfg.S + q] = 0.0f;
    }
    if (eos) {  // EOS boundary: keep every connection for the re-rank
#pragma unroll
      for (int g = 0; g < SMALL_G; ++g) {
        if ((u32)g >= G) continue;
        bv.eosGbeam[(size_t)sent * GMAX + g] = key[g];
        bv.eosLocal[(size_t)sent * GMAX + g] = loc[g];
      }
    }
  } else {
    float tot[MAX_BEAM], loc[MAX_BEAM];
    for (u32 gi = 0; gi < n; ++gi) {
      loc[gi] = step_local_score(bv, cfg, wv, z, G, node, pos, gi, t1idx, t0);
      tot[gi] = loc[gi] + cand_score(gbKey[gi]);
    }
    write_beam(bv, cfg, node, gbKey, (int)n, tot, loc);
    if (eos) {
      for (u32 g = 0; g < G; ++g) {
        bv.eosGbeam[(size_t)sent * GMAX + g] = gbKey[g];
        bv.eosLocal[(size_t)sent * GMAX + g] = loc[g];
      }
    }
  }
}

// Full-beam scoring (computeScoresFull, analyzer_impl.cc:197-248; selected when the global
// beam is off): every (left node, beam entry of it) is scored against the right node.
// One thread per right node: applyT1 per left node (bigrams, score_processor.cc:136-144),
// applyT2 per beam entry (trigrams, :146-156), then makeBeams (:208-244): the `beam` largest
// packed (total, left, beam) keys.  Returns the number of table reads it made.
template <typename WL>
JPP_HD u32 wave_full_node(const FeatTables& ft, const BatchView& bv, const ScoreCfg& cfg, const WL& wl, u32 node) {
  const u32 gB = bv.nodeBnd[node];
  const u32 sent = bv.bndSent[gB];
  const u32 b0 = bv.sentBndOff[sent];
  const u32 B = (u32)cfg.beam, mask = ft.mask;
  const u32 e0 = bv.bndEndOff[gB];
  const u32 L = bv.bndEndOff[gB + 1] - e0;
  const bool lastRow = ft.dynamicOrder || (node + 1 == bv.bndNodeOff[gB + 1]);
  const PatRow P = pat_row(bv, node);
  const float t0 = bv.t0[node];
  u64 top[MAX_BEAM];
  float topLocal[MAX_BEAM];
  u32 n = 0, gathers = 0;
  for (u32 left = 0; left < L; ++left) {
    const u32 lnode = bv.endNode[e0 + left];
    const u32 lbnd = bv.ends[e0 + left].boundary;
    const jpp_b200_beam_element* lrow = bv.beam + (size_t)lnode * B;
    const PatRow t1 = pat_row(bv, lnode);
    auto biIdx = [&](u32 f) -> u32 { return (u32)mix(mix(ft.biState[f], P[ft.biT0[f]]), t1[ft.biT1[f]]) & mask; };
    float T1 = t0;
    T1 += gsum_row(wl, ft.nBi, ft.biVars, biIdx, lastRow);
    gathers += ft.nBi;
    for (u32 slot = 0; slot < B; ++slot) {
      const jpp_b200_beam_element el = lrow[slot];
      if (is_fake(el)) break;
      const PatRow t2 = pat_row(bv, bv.endNode[bv.bndEndOff[b0 + lbnd] + el.left]);
      auto triIdx = [&](u32 f) -> u32 {
        return (u32)mix(mix(mix(ft.triState[f], P[ft.triT0[f]]), t1[ft.triT1[f]]), t2[ft.triT2[f]]) & mask;
      };
      float T2 = T1;
      T2 += gsum_row(wl, ft.nTri, ft.triVars, triIdx, lastRow);
      gathers += ft.nTri;
      const u64 key = pack_cand(el.total_score + T2, left, slot);  // fillBeamCandidates, :165-191
      if (n == B && key < top[n - 1]) continue;
      u32 pos = n < B ? n : B - 1;
      while (pos > 0 && top[pos - 1] < key) { top[pos] = top[pos - 1]; topLocal[pos] = topLocal[pos - 1]; --pos; }
      top[pos] = key;
      topLocal[pos] = T2;
      if (n < B) ++n;
    }
  }
  jpp_b200_beam_element* row = bv.beam + (size_t)node * B;
  float* srow = bv.beamScores + (size_t)node * B * cfg.S;
  for (u32 k = 0; k < B; ++k) {
    if (k < n) {
      jpp_b200_beam_element e;
      e.left = (u16)(top[k] >> 16);
      e.prev_beam = (u16)top[k];
      e.total_score = cand_score(top[k]);
      row[k] = e;
      srow[(size_t)k * cfg.S] = topLocal[k];
    } else {
      row[k] = fake_elem();
      srow[(size_t)k * cfg.S] = 0.0f;
    }
  }
  return gathers;
}

// 1-best path of a sentence: walk EOS beam[0] back through the (left, prev_beam)
// links (analysis_result.cc:25-76) and store the node indices in sentence order.
JPP_HD void best_path_sentence(const BatchView& bv, const ScoreCfg& cfg, u32* pathLen, u32* pathNodes, u32 sent,
                                u32 nodeBase = 0) {
  u32 b0;
  const u32 nb = sent_span(bv, sent, b0);
  const u32 B = (u32)cfg.beam;
  pathLen[sent] = 0;
  u32 len = 0;
  for (int pass = 0; pass < 2 && nb > 3; ++pass) {
    u32 node = bv.bndNodeOff[b0 + nb - 1], bnd = nb - 1, slot = 0, i = 0;
    for (;;) {
      const jpp_b200_beam_element el = bv.beam[(size_t)node * B + slot];
      if (is_fake(el)) { len = 0; i = 0; break; }  // unreachable EOS: no path
      const u32 e = bv.bndEndOff[b0 + bnd] + el.left;
      slot = el.prev_beam;
      bnd = bv.ends[e].boundary;
      node = bv.endNode[e];
      if (bnd < 2) break;
      if (pass == 1) pathNodes[b0 + (len - 1 - i)] = node + nodeBase;
      ++i;
    }
    if (pass == 0) len = i;
    if (len == 0) break;
  }
  pathLen[sent] = len;
  for (u32 i = len; i < nb; ++i) pathNodes[b0 + i] = 0xFFFFFFFFu;  // unused slots of the sentence: defined
}

}  // namespace jppb200
