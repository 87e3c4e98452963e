/*
 * jpp_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * Plain-C, single-file CPU restatement of the Juman++ lattice-scoring hot path
 *   AnalyzerImpl::bootstrapAnalysis + computeScores
 *   (reference src/core/analysis/analyzer_impl.cc:179-309)
 * over the staged-batch layout of include/jpp_b200.h.  It is the checker the
 * CUDA path is compared with bit for bit (tests/, __graft_entry__.smoke()) and
 * the `port` CPU baseline of bench.py.  Nothing in jumanpp_b200/ may link,
 * load or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file
 * against golden vectors produced by the unmodified reference compiled from
 * /root/reference (oracle/build_ref.py + oracle/ref_harness.cc): stored
 * pattern hashes (u64 ==), every beam element (left, prev_beam ==, total score
 * float bits ==), per-connection scores (float bits ==) and the RNN re-ranked
 * EOS beam.
 *
 * What is restated, with the float summation ORDER of the code the CLI runs
 * (the generated static features, not the dynamic fallback):
 *   hashing            util/fast_hash_rot.h:38-54, seahash.h:15-17
 *   patterns           impl/feature_impl_pattern.h:28-39, feature_impl_prim.h:62-236,
 *                      feature_impl_compute.cc:12-27,59-63
 *   n-gram hashes      impl/feature_impl_ngram_partial.h:29-41,58-100,123-178
 *   T0 (unigrams)      codegen/pattern_feature_codegen.cc:199-313 (4 accumulators,
 *                      last row via perceptron.h:46-72)
 *   bi/tri step 2/3    codegen/partial_ngram_feature_codegen.cc:13-153 (8 / 4
 *                      accumulators, last row via computeUnrolled4RawPerceptron)
 *   applyBiTri         impl/feature_impl_ngram_partial_kernels.h:19-110 (2 accumulators)
 *   global beam        analysis/score_processor.cc:193-206,246-282, score_processor.h:81-115
 *   gbeam scoring      score_processor.cc:284-361,363-424,426-519
 *   full beam          analyzer_impl.cc:197-248, score_processor.cc:97-156,165-244
 *   RNN rescoring      analysis/rnn_scorer_gbeam.cc:37-46,142-286, rnn_id_resolver.cc:173-276,
 *                      rnn/mikolov_rnn_impl.h:98-131,196-257 (Eigen order: see
 *                      oracle/eigen_shim/eigen3/Eigen/Core)
 *   re-rank            score_processor.cc:521-576
 *
 * Tie-breaking: equal float keys are common (homograph entries with identical
 * feature fields), so the sorts the reference links against are emulated
 * exactly: libstdc++ std::sort / std::nth_element and jumanpp's own
 * util::partition (see the "Exact emulation" block below).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "jpp_b200.h"
#include "jpp_oracle.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;

/* ------------------------------------------------------------------ hash --- */
#define SEA_SEED0 0x16f11fe89b0d677cULL
#define SEA_MULT 0x6eed0e9da4d94a4fULL
#define PATTERN_SEED 0x7a11ed00000000ULL
#define UNIGRAM_SEED 0x5123a31421fULL
#define BIGRAM_SEED 0x5123a68442fULL
#define TRIGRAM_SEED 0x51239ab41f1fULL

static inline u64 mix(u64 state, u64 data) {
  u64 v = (state ^ data) * SEA_MULT;
  return (v << 32) | (v >> 32);
}

/* ----------------------------------------------------- counter weights --- */
static inline u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

float jpp_oracle_synth_weight(uint64_t seed, uint64_t i, float scale) {
  u64 h = splitmix64(i + (seed << 40));
  u32 s = (u32)(h & 0xFFFF) + (u32)((h >> 16) & 0xFFFF) + (u32)((h >> 32) & 0xFFFF) + (u32)((h >> 48) & 0xFFFF);
  float v = (float)s * (1.0f / 65536.0f) - 2.0f;
  float f = (float)((double)scale * 1.7320508075688772);
  return v * f;
}

/* ---------------------------------------------------------- compiled model */
typedef struct {
  const jpp_b200_model* m;
  const float* w;
  u32 mask;
  u32 nUni, nBi, nTri;
  u64 uniState[256];   /* Hasher{}.mix(3).mix(index).mix(UnigramSeed) */
  u32 uniT0[256];
  u64 biState[256];
  u32 biT0[256], biT1[256];
  u64 triState[64];
  u32 triT0[64], triT1[64], triT2[64];
  u64 patState[256];
  int uniVars, biVars, triVars;
  int dyn; /* jpp_b200_model.float_order == DYNAMIC: every sum through computeUnrolled4RawPerceptron
              (PartialNgramDynamicFeatureApply, feature_impl_ngram_partial.h:196-371) */
} CModel;

static int compile_model(const jpp_b200_model* m, const float* w, CModel* c) {
  memset(c, 0, sizeof *c);
  c->m = m;
  c->w = w;
  c->mask = (u32)((1ull << m->log2_weights) - 1);
  if (m->num_patterns > 256) return -1;
  for (u32 i = 0; i < m->num_ngrams; ++i) {
    const jpp_b200_ngram* g = &m->ngrams[i];
    u64 s = mix(mix(SEA_SEED0, g->order + 2), g->index);
    if (g->order == 1) {
      if (c->nUni >= 256) return -1;
      c->uniState[c->nUni] = mix(s, UNIGRAM_SEED);
      c->uniT0[c->nUni++] = g->refs[0];
    } else if (g->order == 2) {
      if (c->nBi >= 256) return -1;
      c->biState[c->nBi] = mix(s, BIGRAM_SEED);
      c->biT0[c->nBi] = g->refs[0];
      c->biT1[c->nBi++] = g->refs[1];
    } else if (g->order == 3) {
      if (c->nTri >= 64) return -1;
      c->triState[c->nTri] = mix(s, TRIGRAM_SEED);
      c->triT0[c->nTri] = g->refs[0];
      c->triT1[c->nTri] = g->refs[1];
      c->triT2[c->nTri++] = g->refs[2];
    } else {
      return -1;
    }
  }
  for (u32 p = 0; p < m->num_patterns; ++p) {
    const jpp_b200_pattern* pt = &m->patterns[p];
    c->patState[p] = mix(mix(mix(SEA_SEED0, pt->index), pt->ref_len), PATTERN_SEED);
  }
  /* accumulator counts of the generated code */
  c->dyn = m->float_order == JPP_B200_ORDER_DYNAMIC;
  c->uniVars = c->nUni >= 8 ? 4 : (c->nUni >= 4 ? 2 : 1);            /* pattern_feature_codegen.cc:202-208 */
  c->biVars = c->nBi >= 8 ? 8 : (c->nBi >= 4 ? 4 : (c->nBi >= 2 ? 2 : 1)); /* partial_ngram_feature_codegen.cc:15-22 */
  c->triVars = c->nTri >= 8 ? 8 : (c->nTri >= 4 ? 4 : (c->nTri >= 2 ? 2 : 1)); /* :86-93 */
  return 0;
}

/* optional debugging aid: JPP_ORACLE_DUMP_IDX=<file> appends every gathered table
 * index (u32) -- used once to measure how skewed the gather distribution is */
#include <stdio.h>
static FILE* g_dump;
static int g_dump_init;
static void dump_idx(const u32* idx, int n) {
  if (!g_dump_init) {
    const char* p = getenv("JPP_ORACLE_DUMP_IDX");
    g_dump = p ? fopen(p, "ab") : NULL;
    g_dump_init = 1;
  }
  if (g_dump) fwrite(idx, 4, (size_t)n, g_dump);
}

/* computeUnrolled4RawPerceptron, perceptron.h:46-72 */
static float unrolled4(const float* w, const u32* idx, int n) {
  dump_idx(idx, n);
  float r1 = 0, r2 = 0, r3 = 0, r4 = 0;
  int i;
  for (i = 0; (i + 4) <= n; i += 4) {
    r1 += w[idx[i]];
    r2 += w[idx[i + 1]];
    r3 += w[idx[i + 2]];
    r4 += w[idx[i + 3]];
  }
  switch ((n - i) & 3) {
    case 3: r3 += w[idx[i + 2]]; /* fallthrough */
    case 2: r2 += w[idx[i + 1]]; /* fallthrough */
    case 1: r1 += w[idx[i]];     /* fallthrough */
    default:;
  }
  return r1 + r2 + r3 + r4;
}

/* sum with `nv` interleaved accumulators, generated-code style:
 * first == 1: "float p_j = w" for the first nv terms (pattern_feature_codegen.cc:229-234);
 * first == 0: "float f_j = 0; f_j += w" (partial_ngram_feature_codegen.cc:49-56). */
static float accum_sum(const float* w, const u32* idx, int n, int nv, int assign_first) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  dump_idx(idx, n);
  for (int i = 0; i < n; ++i) {
    float x = w[idx[i]];
    if (assign_first && i < nv) acc[i % nv] = x; else acc[i % nv] += x;
  }
  float s = acc[0];
  for (int v = 1; v < nv; ++v) s = s + acc[v];
  return s;
}

/* 2-accumulator loop of applyBiTriFullKernel (feature_impl_ngram_partial_kernels.h:35-61) */
static void accum2(const float* w, const u32* idx, int n, float* r1o, float* r2o) {
  float r1 = 0, r2 = 0;
  dump_idx(idx, n);
  int f = 0;
  for (; (f + 2) <= n; f += 2) {
    r1 += w[idx[f]];
    r2 += w[idx[f + 1]];
  }
  if (n & 1) r1 += w[idx[f]];
  *r1o = r1;
  *r2o = r2;
}

/* ------------------------------------------------------------- sentence ---- */
typedef struct {
  u16 left, prev;
  float total;
} BeamEl;

typedef struct {
  /* scratch sized for the largest boundary of the batch */
  u32 maxR, maxL;
  u64* pats;       /* [maxR][num_patterns] all patterns of the current right nodes */
  float* t0;       /* [maxR] */
  u64* biSt;       /* [maxR][nBi] */
  u64* triSt;      /* [maxR][nTri] */
  u64* triSt2;     /* [maxR][nTri] */
  float* pre;      /* [rcheck][maxR] */
  float* cutoff;   /* [maxR] */
  u32* cutIdx;     /* [maxR] */
  u64* cand;       /* [maxL*beam] packed BeamCandidate */
  u64* t1pat;      /* [G][stored] */
  u64* t2pat;      /* [G][stored] */
  u32* t1idx;      /* [G] */
  u32* t1node;     /* [G] unique left node index (global node id) */
  float* result;   /* [G] */
  float* scbuf;    /* [maxL] */
  float* T1;       /* [maxR] full-beam */
  float* T2;       /* [maxR] */
  float* fullScores; /* [maxR][beam][maxL] full-beam connection scores */
  u32* idxbuf;     /* [max(G, nUni, nBi)] */
  float* sortScore;
  u32 G;
  /* gbeam record per boundary of the current sentence (for adjustBeamScores) */
  u32* gbOff;      /* [nb+1] */
  u32* gbNode;     /* element -> global node id of the left node */
  u16* gbSlot;     /* element -> beam slot */
  u16* gbBnd;      /* element -> boundary (in sentence) where the left node starts */
  u32 gbCap;
} Scratch;

typedef struct {
  const CModel* cm;
  const jpp_b200_config* cfg;
  const jpp_b200_batch* b;
  jpp_b200_result* r;
  jpp_oracle_stats* stats;
  u32 beam, S, stored;
} Ctx;

static inline u64 pack_cand(float score, u16 left, u16 beam) {
  u32 v;
  memcpy(&v, &score, 4);
  v = (v & 0x80000000u) ? ~v : (v ^ 0x80000000u);
  return ((u64)v << 32) | ((u32)left << 16) | beam;
}
static inline float cand_score(u64 d) {
  u32 v = (u32)(d >> 32);
  v = ((v & 0x80000000u) == 0) ? ~v : (v ^ 0x80000000u);
  float f;
  memcpy(&f, &v, 4);
  return f;
}

static int cmp_u64_desc(const void* a, const void* b) {
  u64 x = *(const u64*)a, y = *(const u64*)b;
  return x < y ? 1 : (x > y ? -1 : 0);
}

static inline int is_fake(const jpp_b200_beam_element* e) { return e->left == 0xFFFF && e->prev_beam == 0xFFFF; }

static void set_fake(jpp_b200_beam_element* e) {
  e->left = 0xFFFF;
  e->prev_beam = 0xFFFF;
  u32 bits = 0xFFFFFFFFu; /* memset 0xff, score_processor.cc:466-468 */
  memcpy(&e->total_score, &bits, 4);
}

/* primitive feature value, feature_impl_prim.h:62-236 */
static u64 prim_value(const jpp_b200_primitive* p, const i32* entry, const i32* ph, const jpp_b200_node_info* ni,
                      const jpp_b200_codepoint* cps, u32 ncp) {
  switch (p->kind) {
    case JPP_B200_PRIM_COPY: return (u64)(u32)entry[p->a];
    case JPP_B200_PRIM_SHIFT_MASK: return (u64)(((u32)entry[p->a] >> p->b) & (u32)p->c);
    case JPP_B200_PRIM_PROVIDED: return ph ? (u64)(u32)ph[p->a] : 0;
    case JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE: return (u64)(i32)((i32)ni->end - (i32)ni->start);
    case JPP_B200_PRIM_CODEPOINT: {
      u64 value = ~(u64)0;
      if (p->a > 0) {
        i32 pos = (i32)ni->end + (p->a - 1);
        if ((u32)pos < ncp) value = cps[pos].codepoint;
      } else {
        i32 pos = (i32)ni->start + p->a;
        if (0 <= pos && (u32)pos < ncp) value = cps[pos].codepoint;
      }
      return value;
    }
    case JPP_B200_PRIM_CODEPOINT_TYPE: {
      u64 value = 0;
      if (p->a == 0) {
        for (i32 i = ni->start; i < (i32)ni->end; ++i) value |= (u32)cps[i].char_class;
      } else if (p->a > 0) {
        i32 pos = (i32)ni->end + (p->a - 1);
        if ((u32)pos < ncp) value = (u32)cps[pos].char_class;
      } else {
        i32 pos = (i32)ni->start + p->a;
        if (0 <= pos && (u32)pos < ncp) value = (u32)cps[pos].char_class;
      }
      return value;
    }
    default: return 0;
  }
}

static void node_patterns(const CModel* cm, const i32* entry, const i32* ph, const jpp_b200_node_info* ni,
                          const jpp_b200_codepoint* cps, u32 ncp, u64* out /* [num_patterns] by index */) {
  const jpp_b200_model* m = cm->m;
  u64 prim[256];
  for (u32 i = 0; i < m->num_primitives; ++i) prim[i] = prim_value(&m->primitives[i], entry, ph, ni, cps, ncp);
  for (u32 p = 0; p < m->num_patterns; ++p) {
    const jpp_b200_pattern* pt = &m->patterns[p];
    u64 h = cm->patState[p];
    for (u32 k = 0; k < pt->ref_len; ++k) {
      const jpp_b200_computation* c = &m->computations[m->pattern_ref_pool[pt->ref_off + k]];
      if (c->true_len == 0 && c->false_len == 0) {
        h = mix(h, prim[c->primitive]);
      } else if (prim[c->primitive] != 0) {
        for (u32 t = 0; t < c->true_len; ++t) h = mix(h, prim[m->branch_pool[c->true_off + t]]);
      } else {
        for (u32 t = 0; t < c->false_len; ++t) h = mix(h, prim[m->branch_pool[c->false_off + t]]);
      }
    }
    out[pt->index] = h;
  }
}

/* ---------------------------------------------------------------------------
 * Exact emulation of the sorting the reference links against, so that EQUAL
 * float keys land where they land in the reference:
 *   - libstdc++ std::sort      (introsort: median-of-3 + unguarded partition
 *     down to 16-element runs, then insertion sort; heap sort when the depth
 *     limit 2*floor(log2 n) is exhausted)          used by makeT0Beam
 *   - libstdc++ std::nth_element (introselect)      used by makeT0cutoffBeam
 *   - jumanpp util::partition / part_step (reference src/util/stl_util.h:50-135)
 * All operate on u32 index arrays with comp(a, b) := score[a] > score[b].
 * tests/test_sort_emulation.py checks these against the real std:: functions.
 * ------------------------------------------------------------------------- */
#define CMP(a, b) (score[(a)] > score[(b)])
#define SWAPU(x, y) do { u32 t_ = (x); (x) = (y); (y) = t_; } while (0)

static void em_insertion_sort(u32* first, u32* last, const float* score) {
  if (first == last) return;
  for (u32* i = first + 1; i != last; ++i) {
    u32 val = *i;
    if (CMP(val, *first)) {
      memmove(first + 1, first, (size_t)(i - first) * sizeof(u32));
      *first = val;
    } else {
      u32* l = i;
      u32* next = i - 1;
      while (CMP(val, *next)) { *l = *next; l = next; --next; }
      *l = val;
    }
  }
}

static void em_unguarded_insertion_sort(u32* first, u32* last, const float* score) {
  for (u32* i = first; i != last; ++i) {
    u32 val = *i;
    u32* l = i;
    u32* next = i - 1;
    while (CMP(val, *next)) { *l = *next; l = next; --next; }
    *l = val;
  }
}

static void em_move_median_to_first(u32* result, u32* a, u32* b, u32* c, const float* score) {
  if (CMP(*a, *b)) {
    if (CMP(*b, *c)) SWAPU(*result, *b);
    else if (CMP(*a, *c)) SWAPU(*result, *c);
    else SWAPU(*result, *a);
  } else if (CMP(*a, *c)) SWAPU(*result, *a);
  else if (CMP(*b, *c)) SWAPU(*result, *c);
  else SWAPU(*result, *b);
}

static u32* em_unguarded_partition(u32* first, u32* last, u32* pivot, const float* score) {
  for (;;) {
    while (CMP(*first, *pivot)) ++first;
    --last;
    while (CMP(*pivot, *last)) --last;
    if (!(first < last)) return first;
    SWAPU(*first, *last);
    ++first;
  }
}

static u32* em_partition_pivot(u32* first, u32* last, const float* score) {
  u32* mid = first + (last - first) / 2;
  em_move_median_to_first(first, first + 1, mid, last - 1, score);
  return em_unguarded_partition(first + 1, last, first, score);
}

static void em_push_heap(u32* first, long hole, long top, u32 value, const float* score) {
  long parent = (hole - 1) / 2;
  while (hole > top && CMP(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void em_adjust_heap(u32* first, long hole, long len, u32 value, const float* score) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (CMP(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  em_push_heap(first, hole, top, value, score);
}

static void em_make_heap(u32* first, u32* last, const float* score) {
  long len = last - first;
  if (len < 2) return;
  long parent = (len - 2) / 2;
  for (;;) {
    u32 value = first[parent];
    em_adjust_heap(first, parent, len, value, score);
    if (parent == 0) return;
    parent--;
  }
}

static void em_pop_heap(u32* first, u32* last, u32* result, const float* score) {
  u32 value = *result;
  *result = *first;
  em_adjust_heap(first, 0, last - first, value, score);
}

static void em_heap_select(u32* first, u32* middle, u32* last, const float* score) {
  em_make_heap(first, middle, score);
  for (u32* i = middle; i < last; ++i)
    if (CMP(*i, *first)) em_pop_heap(first, middle, i, score);
}

static void em_sort_heap(u32* first, u32* last, const float* score) {
  while (last - first > 1) {
    --last;
    em_pop_heap(first, last, last, score);
  }
}

static int em_lg(long n) {
  int k = 0;
  while (n > 1) { n >>= 1; ++k; }
  return k;
}

static void em_introsort_loop(u32* first, u32* last, long depth, const float* score) {
  while (last - first > 16) {
    if (depth == 0) {
      em_heap_select(first, last, last, score);
      em_sort_heap(first, last, score);
      return;
    }
    --depth;
    u32* cut = em_partition_pivot(first, last, score);
    em_introsort_loop(cut, last, depth, score);
    last = cut;
  }
}

/* std::sort(first, last, comp) */
void jpp_oracle_std_sort(u32* first, long n, const float* score) {
  if (n <= 0) return;
  u32* last = first + n;
  em_introsort_loop(first, last, 2L * em_lg(n), score);
  if (n > 16) {
    em_insertion_sort(first, first + 16, score);
    em_unguarded_insertion_sort(first + 16, last, score);
  } else {
    em_insertion_sort(first, last, score);
  }
}

/* std::nth_element(first, first + nth, last, comp) */
void jpp_oracle_std_nth_element(u32* first, long nth, long n, const float* score) {
  u32* last = first + n;
  u32* nthp = first + nth;
  if (n <= 0 || nthp == last) return;
  long depth = 2L * em_lg(n);
  while (last - first > 3) {
    if (depth == 0) {
      em_heap_select(first, nthp + 1, last, score);
      SWAPU(*first, *nthp);
      return;
    }
    --depth;
    u32* cut = em_partition_pivot(first, last, score);
    if (cut <= nthp) first = cut; else last = cut;
  }
  em_insertion_sort(first, last, score);
}

/* jumanpp util::part_step (stl_util.h:50-104) */
static u32* em_part_step(u32* start, u32* end, const float* score) {
  long sz = end - start;
  if (sz == 1) return end;
  if (sz == 2) {
    u32* n = start + 1;
    if (CMP(*n, *start)) SWAPU(*n, *start);
    return n;
  }
  if (sz == 3) {
    u32 *n0 = start, *n1 = start + 1, *n2 = start + 2;
    if (CMP(*n1, *n0)) SWAPU(*n1, *n0);
    if (CMP(*n2, *n1)) SWAPU(*n2, *n1);
    if (CMP(*n1, *n0)) SWAPU(*n1, *n0);
    return n1;
  }
  u32* pivot = start + sz / 2;
  --end;
  SWAPU(*pivot, *end);
  pivot = end;
  --end;
  while (start != end) {
    if (CMP(*start, *pivot)) {
      ++start;
    } else {
      SWAPU(*start, *end);
      --end;
    }
  }
  if (CMP(*pivot, *end)) {
    SWAPU(*pivot, *end);
  } else {
    ++end;
    SWAPU(*pivot, *end);
  }
  return end;
}

/* jumanpp util::partition (stl_util.h:106-135); returns the new length */
long jpp_oracle_ref_partition(u32* first, long n, const float* score, long minSize, long maxSize) {
  u32* start = first;
  u32* end = first + n;
  for (;;) {
    u32* mid = em_part_step(start, end, score);
    long sz = mid - start;
    if (minSize <= sz && sz <= maxSize) return mid - first;
    if (sz > maxSize) {
      end = mid;
      continue;
    }
    sz += 1;
    start = mid;
    ++start;
    minSize -= sz;
    maxSize -= sz;
    if (minSize == 0) return start - first;
  }
}
#undef CMP
#undef SWAPU

/* makeT0Beam ordering (score_processor.cc:426-438): iota, optional
 * util::partition when more than maxBeam*4/3 candidates, then std::sort */
static int sort_for_beam(u32* idx, int n, const float* score, int maxBeam) {
  for (int i = 0; i < n; ++i) idx[i] = (u32)i;
  long len = n;
  long boundary = (long)maxBeam * 4 / 3;
  if (n > boundary) len = jpp_oracle_ref_partition(idx, n, score, maxBeam, boundary);
  jpp_oracle_std_sort(idx, len, score);
  return (int)len;
}

typedef struct {
  u32 boundary0;   /* first boundary (global) */
  u32 nb;
  const u32* bndNodeOff;
  const u32* bndEndOff;
} SentView;

static inline u32 node_of(const Ctx* c, const SentView* sv, jpp_b200_node_ptr p) {
  return c->b->boundary_node_off[sv->boundary0 + p.boundary] + p.position;
}

/* the node that precedes beam element (node, slot): ConnectionPtr::previous -> (boundary,right) */
static u32 prev_node_of(const Ctx* c, const SentView* sv, u32 node, u32 nodeBoundary, u32 slot, jpp_b200_beam_element* beams) {
  const jpp_b200_beam_element* el = &beams[(size_t)node * c->beam + slot];
  u32 e = c->b->boundary_end_off[sv->boundary0 + nodeBoundary] + el->left;
  return node_of(c, sv, c->b->ends[e]);
}

static void make_beam(const Ctx* c, Scratch* s, jpp_b200_beam_element* beamRow, float* scoreRow /* [beam][S] or NULL */,
                      const u64* gbeam, int n, const float* totals, const float* local) {
  int maxBeam = (int)c->beam;
  int len = sort_for_beam(s->idxbuf, n, totals, maxBeam);
  int k = 0;
  for (; k < len && k < maxBeam; ++k) {
    u32 g = s->idxbuf[k];
    beamRow[k].left = (u16)(gbeam[g] >> 16);
    beamRow[k].prev_beam = (u16)gbeam[g];
    beamRow[k].total_score = totals[g];
    if (scoreRow) {
      scoreRow[(size_t)k * c->S] = local[g];
      for (u32 q = 1; q < c->S; ++q) scoreRow[(size_t)k * c->S + q] = 0.0f;
    }
  }
  for (; k < maxBeam; ++k) {
    set_fake(&beamRow[k]);
    if (scoreRow) for (u32 q = 0; q < c->S; ++q) scoreRow[(size_t)k * c->S + q] = 0.0f;
  }
}

/* --------------------------------------------------------------- RNN ------ */
static const u64 PRIMES[] = {108641969, 116049371, 125925907, 133333309, 145678979, 175308587, 197530793, 234567803,
                             251851741, 264197411, 330864029, 399999781, 407407183, 459258997, 479012069, 545678687,
                             560493491, 607407037, 629629243, 656789717, 716048933, 718518067, 725925469, 733332871,
                             753085943, 755555077, 782715551, 790122953, 812345159, 814814293, 893826581, 923456189,
                             940740127, 953085797, 985184539, 990122807};
#define PRIMES_SIZE 36

/* dot product in the order of oracle/eigen_shim (32 interleaved fmaf partials + tree) */
static float dot32(const float* a, const float* b, u32 n) {
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (u32 k = 0; k < n; ++k) {
    part[k & 31] = fmaf(a[k], b[k], part[k & 31]);
  }
  for (int stride = 16; stride >= 1; stride >>= 1)
    for (int l = 0; l < stride; ++l) part[l] = part[l] + part[l + stride];
  return part[0];
}

static void rnn_new_context(const jpp_b200_rnn_model* rm, const float* ctx, i32 id, float* out) {
  u32 H = rm->layer_size;
  if (id == -1) id = 0; /* rnn_scorer_gbeam.cc:126-129 */
  const float* emb = rm->embeddings + (size_t)id * H;
  for (u32 i = 0; i < H; ++i) {
    float v = dot32(rm->matrix + (size_t)i * H, ctx, H); /* mikolov_rnn_impl.h:211 */
    v = v + emb[i];                                      /* :212 */
    out[i] = 1.0f / (1.0f + expf(-v));                   /* :213 */
  }
}

static float rnn_score(const jpp_b200_rnn_model* rm, const float* ctx, i32 prevId, i32 id) {
  u32 H = rm->layer_size;
  i32 eid = id == -1 ? 0 : id;
  float s = dot32(rm->nce_embeddings + (size_t)eid * H, ctx, H); /* :216-223 */
  u64 hashMax = rm->maxent_size - rm->vocab_size;                /* :135-136 */
  u32 order = rm->maxent_order;
  float me = 0;
  /* context slots all hold prev->id (rnn_scorer_gbeam.cc:170-186) */
  for (u32 i = 0; i < order; ++i) {
    u64 x = PRIMES[0] * PRIMES[1];
    for (u32 j = 1; j <= i; ++j) {
      u64 primeIdx = (i * PRIMES[j] + j) % PRIMES_SIZE;
      u64 conv = (u64)(int64_t)prevId + 1;
      x += PRIMES[primeIdx] * conv;
    }
    u64 idx = ((x % hashMax) + (u64)(int64_t)id) % hashMax; /* :104-131: (indices[i] + word) % hashMax */
    float w = rm->maxent[idx];
    me = (i == 0) ? w : me + w;
  }
  if (order > 0) s = s + me; else s = 0; /* MikolovScoreCalculator::addScores case 0 zero-fills, :58-60 */
  s = s - rm->nce_constant; /* :239-243 */
  return s;
}

typedef struct {
  u32 node;
  u16 slot;
} ElemRef;

/* ------------------------------------------------------------ one sentence */
static int score_sentence(const Ctx* c, Scratch* s, u32 sent) {
  const jpp_b200_batch* B = c->b;
  const CModel* cm = c->cm;
  const jpp_b200_model* m = cm->m;
  const float* w = cm->w;
  const u32 mask = cm->mask;
  const u32 nU = cm->nUni, nB = cm->nBi, nT = cm->nTri, NP = m->num_patterns, SP = c->stored;
  SentView sv = {B->sent_boundary_off[sent], B->sent_boundary_off[sent + 1] - B->sent_boundary_off[sent], 0, 0};
  const u32 nb = sv.nb;
  const jpp_b200_codepoint* cps = B->codepoints + B->sent_codepoint_off[sent];
  const u32 ncp = B->sent_codepoint_off[sent + 1] - B->sent_codepoint_off[sent];
  jpp_b200_beam_element* beams = c->r->node_beam;
  float* bscores = c->r->node_beam_scores;
  u64* patsOut = c->r->node_patterns;
  const int maxBeam = (int)c->beam;
  const int gbeamMode = c->cfg->global_beam_size > 0;
  u64 gathers = 0;

  /* BOS (analyzer_impl.cc:185-192, lattice_builder.cc:171-176) */
  for (u32 bb = 0; bb < 2 && bb < nb; ++bb) {
    u32 n0 = B->boundary_node_off[sv.boundary0 + bb];
    for (u32 p = 0; p < SP; ++p) patsOut[(size_t)n0 * SP + p] = (u64)(u32)B->node_info[n0].entry_ptr;
    for (int k = 0; k < maxBeam; ++k) {
      jpp_b200_beam_element* e = &beams[(size_t)n0 * c->beam + k];
      if (k == 0) {
        e->left = 0; e->prev_beam = 0; e->total_score = 0;
      } else {
        e->left = 0xFFFF; e->prev_beam = 0xFFFF; e->total_score = -3.402823466e+38f; /* EntryBeam::fake() */
      }
      if (bscores) for (u32 q = 0; q < c->S; ++q) bscores[((size_t)n0 * c->beam + k) * c->S + q] = 0;
    }
    if (c->r->node_t0_score) c->r->node_t0_score[n0] = 0;
  }
  if (nb <= 3) { /* analyzer_impl.cc:255-258: nothing scored, EOS beam untouched (fake) */
    if (nb == 3) {
      u32 n0 = B->boundary_node_off[sv.boundary0 + 2];
      for (int k = 0; k < maxBeam; ++k) set_fake(&beams[(size_t)n0 * c->beam + k]);
    }
    return 0;
  }
  s->gbOff[0] = s->gbOff[1] = s->gbOff[2] = 0;
  u32 gbCount = 0;
  u32 eosG = 0;
  u64 eosGbeam[64];
  float eosLocal[64]; /* perceptron score of every EOS-gbeam connection (the beam keeps only `beam` of them) */

  for (u32 bnd = 2; bnd < nb; ++bnd) {
    const u32 gB = sv.boundary0 + bnd;
    const u32 n0 = B->boundary_node_off[gB];
    const u32 R = B->boundary_node_off[gB + 1] - n0;
    const u32 e0 = B->boundary_end_off[gB];
    const u32 L = B->boundary_end_off[gB + 1] - e0;
    s->gbOff[bnd] = gbCount;
    if (R == 0) continue; /* analyzer_impl.cc:263-266 */

    /* ---- T0: patterns, unigram scores, bi/tri state (computeT0All) ---- */
    for (u32 r = 0; r < R; ++r) {
      u32 node = n0 + r;
      u64* P = s->pats + (size_t)r * NP;
      node_patterns(cm, B->node_entry ? B->node_entry + (size_t)node * m->entry_size
                                      : B->entry_rows + (size_t)B->node_entry_row[node] * m->entry_size,
                    B->node_placeholders ? B->node_placeholders + (size_t)node * m->num_placeholders
                    : (B->row_placeholders && !B->node_entry ? B->row_placeholders + (size_t)B->node_entry_row[node] * m->num_placeholders : NULL),
                    &B->node_info[node], cps, ncp, P);
      for (u32 p = 0; p < SP; ++p) patsOut[(size_t)node * SP + p] = P[p];
      u32 idx[256];
      for (u32 u = 0; u < nU; ++u) idx[u] = (u32)mix(cm->uniState[u], P[cm->uniT0[u]]) & mask;
      if (!cm->dyn && r + 1 < R) s->t0[r] = accum_sum(w, idx, (int)nU, cm->uniVars, 1);
      else s->t0[r] = unrolled4(w, idx, (int)nU);
      for (u32 f = 0; f < nB; ++f) s->biSt[(size_t)r * nB + f] = mix(cm->biState[f], P[cm->biT0[f]]);
      for (u32 f = 0; f < nT; ++f) s->triSt[(size_t)r * nT + f] = mix(cm->triState[f], P[cm->triT0[f]]);
      if (c->r->node_t0_score) c->r->node_t0_score[node] = s->t0[r];
    }
    gathers += (u64)R * nU;

    if (!gbeamMode) {
      /* ---------------- full beam: analyzer_impl.cc:197-248 ---------------- */
      for (u32 r = 0; r < R; ++r)
        for (int k = 0; k < maxBeam; ++k) set_fake(&beams[(size_t)(n0 + r) * c->beam + k]);
      for (u32 t1 = 0; t1 < L; ++t1) {
        u32 lnode = node_of(c, &sv, B->ends[e0 + t1]);
        const u64* p1 = patsOut + (size_t)lnode * SP;
        /* applyT1 (score_processor.cc:136-144) */
        for (u32 r = 0; r < R; ++r) {
          u32 idx[256];
          for (u32 f = 0; f < nB; ++f) idx[f] = (u32)mix(s->biSt[(size_t)r * nB + f], p1[cm->biT1[f]]) & mask;
          float add = (!cm->dyn && r + 1 < R) ? accum_sum(w, idx, (int)nB, cm->biVars, 0) : unrolled4(w, idx, (int)nB);
          s->T1[r] = s->t0[r] + add;
          for (u32 f = 0; f < nT; ++f) s->triSt2[(size_t)r * nT + f] = mix(s->triSt[(size_t)r * nT + f], p1[cm->triT1[f]]);
        }
        gathers += (u64)R * nB;
        const jpp_b200_beam_element* lb = &beams[(size_t)lnode * c->beam];
        u32 lbnd = B->ends[e0 + t1].boundary;
        for (int bi = 0; bi < maxBeam && !is_fake(&lb[bi]); ++bi) {
          u32 t2node = prev_node_of(c, &sv, lnode, lbnd, (u32)bi, beams);
          const u64* p2 = patsOut + (size_t)t2node * SP;
          for (u32 r = 0; r < R; ++r) {
            u32 idx[64];
            for (u32 f = 0; f < nT; ++f) idx[f] = (u32)mix(s->triSt2[(size_t)r * nT + f], p2[cm->triT2[f]]) & mask;
            float add = (!cm->dyn && r + 1 < R) ? accum_sum(w, idx, (int)nT, cm->triVars, 0) : unrolled4(w, idx, (int)nT);
            s->fullScores[((size_t)r * c->beam + bi) * L + t1] = s->T1[r] + add;
          }
          gathers += (u64)R * nT;
        }
      }
      /* makeBeams (score_processor.cc:208-244) */
      for (u32 r = 0; r < R; ++r) {
        u32 cnt = 0;
        for (u32 left = 0; left < L; ++left) {
          u32 lnode = node_of(c, &sv, B->ends[e0 + left]);
          const jpp_b200_beam_element* lb = &beams[(size_t)lnode * c->beam];
          for (int bi = 0; bi < maxBeam && !is_fake(&lb[bi]); ++bi) {
            float local = s->fullScores[((size_t)r * c->beam + bi) * L + left];
            s->cand[cnt++] = pack_cand(lb[bi].total_score + local, (u16)left, (u16)bi);
          }
        }
        qsort(s->cand, cnt, sizeof(u64), cmp_u64_desc);
        jpp_b200_beam_element* row = &beams[(size_t)(n0 + r) * c->beam];
        int k = 0;
        for (; k < (int)cnt && k < maxBeam; ++k) {
          u16 left = (u16)(s->cand[k] >> 16), bi = (u16)s->cand[k];
          row[k].left = left;
          row[k].prev_beam = bi;
          row[k].total_score = cand_score(s->cand[k]);
          if (bscores) bscores[((size_t)(n0 + r) * c->beam + k) * c->S] = s->fullScores[((size_t)r * c->beam + bi) * L + left];
        }
        for (; k < maxBeam; ++k) {
          set_fake(&row[k]);
          if (bscores) bscores[((size_t)(n0 + r) * c->beam + k) * c->S] = 0;
        }
      }
      continue;
    }

    /* ---------------- makeGlobalBeam (score_processor.cc:246-282) --------- */
    u32 cnt = 0;
    for (u32 left = 0; left < L; ++left) {
      u32 lnode = node_of(c, &sv, B->ends[e0 + left]);
      const jpp_b200_beam_element* lb = &beams[(size_t)lnode * c->beam];
      for (int bi = 0; bi < maxBeam && !is_fake(&lb[bi]); ++bi) s->cand[cnt++] = pack_cand(lb[bi].total_score, (u16)left, (u16)bi);
    }
    qsort(s->cand, cnt, sizeof(u64), cmp_u64_desc);
    u32 G = cnt < (u32)c->cfg->global_beam_size ? cnt : (u32)c->cfg->global_beam_size;
    const u64* gbeam = s->cand;
    if (gbCount + G > s->gbCap) return -2;
    for (u32 i = 0; i < G; ++i) {
      u32 lnode = node_of(c, &sv, B->ends[e0 + (u16)(gbeam[i] >> 16)]);
      s->gbNode[gbCount + i] = lnode;
      s->gbSlot[gbCount + i] = (u16)gbeam[i];
      s->gbBnd[gbCount + i] = B->ends[e0 + (u16)(gbeam[i] >> 16)].boundary;
    }
    gbCount += G;
    if (bnd == nb - 1) {
      eosG = G;
      for (u32 i = 0; i < G; ++i) eosGbeam[i] = gbeam[i];
    }

    /* dedupT1 / gatherT1 / gatherT2 (:363-409) */
    u32 U = 0;
    for (u32 i = 0; i < G; ++i) {
      u32 lnode = s->gbNode[gbCount - G + i];
      u32 u = 0;
      for (; u < U; ++u) if (s->t1node[u] == lnode) break;
      if (u == U) {
        s->t1node[U] = lnode;
        memcpy(s->t1pat + (size_t)U * SP, patsOut + (size_t)lnode * SP, SP * sizeof(u64));
        ++U;
      }
      s->t1idx[i] = u;
      u32 lbnd = B->ends[e0 + (u16)(gbeam[i] >> 16)].boundary;
      u32 t2node = prev_node_of(c, &sv, lnode, lbnd, (u16)gbeam[i], beams);
      memcpy(s->t2pat + (size_t)i * SP, patsOut + (size_t)t2node * SP, SP * sizeof(u64));
    }

    float* result = s->result;
    float local[64];
    if (c->cfg->right_gbeam_check > 0) {
      u32 chk = (u32)c->cfg->right_gbeam_check;
      u32 full = chk;
      if (R < full) full = R;
      if (G < full) full = G;                 /* fullBeamApplySize, :299-301 */
      u32 toKeep = (u32)c->cfg->right_gbeam_size < R ? (u32)c->cfg->right_gbeam_size : R;
      u32 remaining = G - full;
      u32 theMax = chk < G ? chk : G;         /* computeT0Prescores, :497-511 */
      for (u32 i = 0; i < theMax; ++i) {
        const u64* t1 = s->t1pat + (size_t)s->t1idx[i] * SP;
        const u64* t2 = s->t2pat + (size_t)i * SP;
        float* sc = s->pre + (size_t)i * s->maxR;
        for (u32 r = 0; r < R; ++r) {
          u32 idx[256];
          sc[r] = s->t0[r];
          for (u32 f = 0; f < nB; ++f) idx[f] = (u32)mix(s->biSt[(size_t)r * nB + f], t1[cm->biT1[f]]) & mask;
          sc[r] += (!cm->dyn && r + 1 < R) ? accum_sum(w, idx, (int)nB, cm->biVars, 0) : unrolled4(w, idx, (int)nB);
        }
        for (u32 r = 0; r < R; ++r) {
          u32 idx[64];
          for (u32 f = 0; f < nT; ++f)
            idx[f] = (u32)mix(mix(s->triSt[(size_t)r * nT + f], t1[cm->triT1[f]]), t2[cm->triT2[f]]) & mask;
          sc[r] += (!cm->dyn && r + 1 < R) ? accum_sum(w, idx, (int)nT, cm->triVars, 0) : unrolled4(w, idx, (int)nT);
        }
        gathers += (u64)R * (nB + nT);
      }
      /* makeT0cutoffBeam (:471-495) */
      for (u32 r = 0; r < R; ++r) s->cutIdx[r] = r;
      if (R > toKeep) {
        for (u32 r = 0; r < R; ++r) {
          float v = 0;
          for (u32 j = 0; j < full; ++j) v += s->pre[(size_t)j * s->maxR + r];
          s->cutoff[r] = v;
        }
        jpp_oracle_std_nth_element(s->cutIdx, (long)toKeep, (long)R, s->cutoff);
      }
      for (u32 pos = 0; pos < R; ++pos) {
        u32 t0i = s->cutIdx[pos];
        u32 node = n0 + t0i;
        for (u32 i = 0; i < full; ++i) {
          float v = s->pre[(size_t)i * s->maxR + t0i];
          v += 0.0f;               /* copyT0Scores(..., 0), :411-424 */
          local[i] = v;
          v += cand_score(gbeam[i]);
          result[i] = v;
        }
        u32 used = full;
        if (pos < toKeep && remaining > 0) {
          /* applyBiTri -> applyBiTriFullKernel over ALL U t1 rows, tail tri rows */
          const u64* bs = s->biSt + (size_t)t0i * nB;
          const u64* ts = s->triSt + (size_t)t0i * nT;
          for (u32 u = 0; u < U; ++u) {
            u32 idx[256];
            const u64* t1 = s->t1pat + (size_t)u * SP;
            for (u32 f = 0; f < nB; ++f) idx[f] = (u32)mix(bs[f], t1[cm->biT1[f]]) & mask;
            if (!cm->dyn && u + 1 < U) {
              float r1, r2;
              accum2(w, idx, (int)nB, &r1, &r2);
              s->scbuf[u] = r1 + r2;
            } else {
              s->scbuf[u] = unrolled4(w, idx, (int)nB);
            }
          }
          for (u32 j = 0; j < remaining; ++j) {
            u32 gi = full + j;
            u32 idx[64];
            const u64* t1 = s->t1pat + (size_t)s->t1idx[gi] * SP;
            const u64* t2 = s->t2pat + (size_t)gi * SP;
            for (u32 f = 0; f < nT; ++f) idx[f] = (u32)mix(mix(ts[f], t1[cm->triT1[f]]), t2[cm->triT2[f]]) & mask;
            float v;
            if (!cm->dyn && j + 1 < remaining) {
              float r1, r2;
              accum2(w, idx, (int)nT, &r1, &r2);
              v = s->scbuf[s->t1idx[gi]] + r1 + r2;
            } else {
              v = s->scbuf[s->t1idx[gi]] + unrolled4(w, idx, (int)nT);
            }
            v += s->t0[t0i];        /* copyT0Scores(tail, t0Score) */
            local[gi] = v;
            v += cand_score(gbeam[gi]);
            result[gi] = v;
          }
          gathers += (u64)U * nB + (u64)remaining * nT;
          used = G;
        } else if (pos < toKeep) {
          used = G; /* remaining == 0: makeT0Beam over the whole (head-only) gbeam */
        }
        if (bnd == nb - 1) for (u32 i = 0; i < used; ++i) eosLocal[i] = local[i];
        make_beam(c, s, &beams[(size_t)node * c->beam], bscores ? &bscores[(size_t)node * c->beam * c->S] : NULL, gbeam,
                  (int)used, result, local);
      }
    } else {
      /* score every gbeam element against every right node (:348-359) */
      for (u32 t0i = 0; t0i < R; ++t0i) {
        u32 node = n0 + t0i;
        if (G > 0) {
          const u64* bs = s->biSt + (size_t)t0i * nB;
          const u64* ts = s->triSt + (size_t)t0i * nT;
          for (u32 u = 0; u < U; ++u) {
            u32 idx[256];
            const u64* t1 = s->t1pat + (size_t)u * SP;
            for (u32 f = 0; f < nB; ++f) idx[f] = (u32)mix(bs[f], t1[cm->biT1[f]]) & mask;
            if (!cm->dyn && u + 1 < U) {
              float r1, r2;
              accum2(w, idx, (int)nB, &r1, &r2);
              s->scbuf[u] = r1 + r2;
            } else {
              s->scbuf[u] = unrolled4(w, idx, (int)nB);
            }
          }
          for (u32 gi = 0; gi < G; ++gi) {
            u32 idx[64];
            const u64* t1 = s->t1pat + (size_t)s->t1idx[gi] * SP;
            const u64* t2 = s->t2pat + (size_t)gi * SP;
            for (u32 f = 0; f < nT; ++f) idx[f] = (u32)mix(mix(ts[f], t1[cm->triT1[f]]), t2[cm->triT2[f]]) & mask;
            float v;
            if (!cm->dyn && gi + 1 < G) {
              float r1, r2;
              accum2(w, idx, (int)nT, &r1, &r2);
              v = s->scbuf[s->t1idx[gi]] + r1 + r2;
            } else {
              v = s->scbuf[s->t1idx[gi]] + unrolled4(w, idx, (int)nT);
            }
            v += s->t0[t0i];
            local[gi] = v;
            v += cand_score(gbeam[gi]);
            result[gi] = v;
          }
          gathers += (u64)U * nB + (u64)G * nT;
        }
        if (bnd == nb - 1) for (u32 i = 0; i < G; ++i) eosLocal[i] = local[i];
        make_beam(c, s, &beams[(size_t)node * c->beam], bscores ? &bscores[(size_t)node * c->beam * c->S] : NULL, gbeam,
                  (int)G, result, local);
      }
    }
  }
  s->gbOff[nb] = gbCount;

  /* ------------- RNN rescoring + re-rank (analyzer_impl.cc:286-294) ------- */
  if (gbeamMode && c->S == 2 && m->rnn != NULL && bscores != NULL) {
    const jpp_b200_rnn_model* rm = m->rnn;
    const u32 H = rm->layer_size;
    const u32 eosBnd = nb - 1;
    const u32 eosNode = B->boundary_node_off[sv.boundary0 + eosBnd];
    const u32 eosE0 = B->boundary_end_off[sv.boundary0 + eosBnd];
    /* ---- RnnIdResolver::resolveIdsAtGbeam + RnnIdContainer (rnn_id_resolver.cc:173-276) ----
     * The prefix tree is restated literally, INCLUDING its collapse quirk: when
     * a connection's (id,length)-history hash matches some published node of
     * the same coordinate chain, the connection is attached to crdCache_'s
     * entry `it->second` (the NEWEST published node with that coordinate,
     * :224-231), not to the node whose hash matched. */
    typedef struct { i32 id, idx, boundary, length, prev, nextInBnd; u64 hash; } RNode;
    typedef struct { u32 node; u16 slot; u16 isEos; u32 g; i32 rnode; } RScore;
    const u32 sentNode0 = B->boundary_node_off[sv.boundary0];
    const u32 sentNodes = B->boundary_node_off[sv.boundary0 + nb] - sentNode0;
    u32 rcap = eosG * (nb + 1) + 4;
    RNode* rn = (RNode*)malloc(sizeof(RNode) * rcap);
    RScore* rs = (RScore*)malloc(sizeof(RScore) * rcap);
    i32* bndHead = (i32*)malloc(sizeof(i32) * nb);
    i32* bndCnt = (i32*)calloc(nb, sizeof(i32));
    i32* connNode = (i32*)malloc(sizeof(i32) * (size_t)sentNodes * c->beam); /* ptrCache_ */
    i32 eosConn[64];
    u32 nrn = 0, nrs = 0;
    for (u32 i = 0; i < nb; ++i) bndHead[i] = -1;
    for (size_t i = 0; i < (size_t)sentNodes * c->beam; ++i) connNode[i] = -1;
    /* addBos (:311-329) */
    rn[0] = (RNode){0, 0, 0, 0, -1, -1, 0};
    rn[1] = (RNode){0, 0, 1, 0, 0, -1, 0xdeadbeef0000ULL};
    nrn = 2;
    bndHead[1] = 1; bndCnt[1] = 1;
    const u32 bos1 = B->boundary_node_off[sv.boundary0 + 1];
    connNode[(size_t)(bos1 - sentNode0) * c->beam + 0] = 1; /* ptrCache_[&bosBeam.ptr] = bos1 */
    ElemRef* chainp = (ElemRef*)malloc(sizeof(ElemRef) * (nb + 2));
    u16* chainBnd = (u16*)malloc(sizeof(u16) * (nb + 2));
    for (u32 g = 0; g < eosG; ++g) {
      /* addPrevChain: walk back until a connection already in ptrCache_ */
      u32 len = 0;
      u32 curNode = node_of(c, &sv, B->ends[eosE0 + (u16)(eosGbeam[g] >> 16)]);
      u32 curBnd = B->ends[eosE0 + (u16)(eosGbeam[g] >> 16)].boundary;
      u16 curSlot = (u16)eosGbeam[g];
      i32 base;
      for (;;) {
        i32 known = connNode[(size_t)(curNode - sentNode0) * c->beam + curSlot];
        if (known != -1) { base = known; break; }
        chainp[len].node = curNode; chainp[len].slot = curSlot; chainBnd[len] = (u16)curBnd; ++len;
        const jpp_b200_beam_element* el = &beams[(size_t)curNode * c->beam + curSlot];
        jpp_b200_node_ptr pp = B->ends[B->boundary_end_off[sv.boundary0 + curBnd] + el->left];
        curSlot = el->prev_beam;
        curBnd = pp.boundary;
        curNode = node_of(c, &sv, pp);
      }
      /* unwind from the deepest new connection to the EOS fake connection (k == -1) */
      i32 first = base, prev = base;
      i32 newTip = -1;      /* first of the not yet published new nodes (deepest) */
      for (i32 k = (i32)len - 1; k >= -1; --k) {
        i32 id, length; u32 cb;
        if (k >= 0) {
          id = B->node_rnn_id ? B->node_rnn_id[chainp[k].node] : B->row_rnn_id[B->node_entry_row[chainp[k].node]];
          length = (i32)B->node_info[chainp[k].node].end - (i32)B->node_info[chainp[k].node].start;
          cb = chainBnd[k];
        } else {
          id = 0; length = 0; cb = eosBnd; /* nodeCache_[{eos,0}] = {eos,0,0}, :306-308 */
        }
        u64 data = (u64)(u32)id | ((u64)(u16)length << 32);
        u64 hv = (rn[prev].hash ^ data) * SEA_MULT;
        hv = hv ^ (hv >> 32); /* FastHash1::mix, util/fast_hash.h:46-51 */
        /* crdCache_.find(coord): newest published node of (boundary,length,id) */
        i32 it = -1;
        for (i32 q = bndHead[cb]; q != -1; q = rn[q].nextInBnd)
          if (rn[q].length == length && rn[q].id == id) { it = q; break; }
        int collapsed = 0;
        if (it != -1) {
          for (i32 q = it; q != -1; q = rn[q].nextInBnd) {
            if (rn[q].hash == hv) { collapsed = 1; break; }
          }
        }
        i32 mine;
        if (collapsed) {
          mine = it;
          rs[nrs++] = (RScore){k >= 0 ? chainp[k].node : 0, k >= 0 ? chainp[k].slot : 0, (u16)(k < 0), g, mine};
          first = mine;
          newTip = -1;
        } else {
          mine = (i32)nrn;
          rn[nrn++] = (RNode){id, -1, (i32)cb, length, prev, -1, hv};
          if (newTip == -1) newTip = mine;
        }
        if (k >= 0) connNode[(size_t)(chainp[k].node - sentNode0) * c->beam + chainp[k].slot] = mine; else eosConn[g] = mine;
        prev = mine;
      }
      /* addPath publish loop (:247-266): from the tip back to `first` */
      {
        i32 last = prev;
        i32 k = -1;
        while (last != first) {
          u32 cb = (u32)rn[last].boundary;
          rn[last].idx = bndCnt[cb];
          rn[last].nextInBnd = bndHead[cb];
          bndHead[cb] = last;
          bndCnt[cb] += 1;
          rs[nrs++] = (RScore){k >= 0 ? chainp[k].node : 0, k >= 0 ? chainp[k].slot : 0, (u16)(k < 0), g, last};
          last = rn[last].prev;
          ++k;
        }
      }
    }
    /* contexts in boundary order (rnn_scorer_gbeam.cc:142-156), BOS state (:37-46) */
    float* ctxs = (float*)malloc(sizeof(float) * (size_t)nrn * H);
    {
      float* zeros = (float*)calloc(H, sizeof(float));
      rnn_new_context(rm, zeros, 0, ctxs + (size_t)1 * H);
      free(zeros);
    }
    for (u32 bb = 2; bb < eosBnd; ++bb) {
      for (i32 q = bndHead[bb]; q != -1; q = rn[q].nextInBnd) {
        rnn_new_context(rm, ctxs + (size_t)rn[q].prev * H, rn[q].id, ctxs + (size_t)q * H);
        if (c->stats) __sync_fetch_and_add(&c->stats->rnn_nodes, 1);
      }
    }
    /* scores (scoreBoundary + copyScoresToLattice, :226-267) */
    for (u32 i = 0; i < nrs; ++i) {
      const RNode* node = &rn[rs[i].rnode];
      float sc;
      if (node->id == rm->unk_id) {
        sc = rm->unk_constant + rm->unk_length_penalty * (float)node->length;
      } else {
        sc = rnn_score(rm, ctxs + (size_t)node->prev * H, rn[node->prev].id, node->id);
      }
      if (rs[i].isEos) s->sortScore[rs[i].g] = sc;
      else bscores[((size_t)rs[i].node * c->beam + rs[i].slot) * c->S + 1] = sc;
      if (c->stats) __sync_fetch_and_add(&c->stats->rnn_scores, 1);
    }
    free(chainp); free(chainBnd); free(rn); free(rs); free(bndHead); free(bndCnt); free(connNode); free(ctxs);

    /* adjustBeamScores (score_processor.cc:521-549) in boundary order */
    const float w0 = m->score_weights[0], w1 = m->score_weights[1];
    for (u32 bnd = 3; bnd < nb; ++bnd) {
      u32 R = B->boundary_node_off[sv.boundary0 + bnd + 1] - B->boundary_node_off[sv.boundary0 + bnd];
      if (R == 0) continue;
      for (u32 i = s->gbOff[bnd]; i < s->gbOff[bnd + 1]; ++i) {
        u32 node = s->gbNode[i];
        u16 slot = s->gbSlot[i];
        jpp_b200_beam_element* el = &beams[(size_t)node * c->beam + slot];
        const float* sc = &bscores[((size_t)node * c->beam + slot) * c->S];
        float localScore = 0;
        localScore += sc[0] * w0;
        localScore += sc[1] * w1;
        /* previous element */
        u32 nbnd = s->gbBnd[i];
        jpp_b200_node_ptr pp = B->ends[B->boundary_end_off[sv.boundary0 + nbnd] + el->left];
        u32 pnode = node_of(c, &sv, pp);
        localScore += beams[(size_t)pnode * c->beam + el->prev_beam].total_score;
        el->total_score = localScore;
      }
    }
    /* remakeEosBeam (:551-576) */
    float full[64], loc[64];
    u64 gb2[64];
    for (u32 g = 0; g < eosG; ++g) {
      u16 left = (u16)(eosGbeam[g] >> 16), bi = (u16)eosGbeam[g];
      u32 lnode = node_of(c, &sv, B->ends[eosE0 + left]);
      float beamScore = beams[(size_t)lnode * c->beam + bi].total_score;
      gb2[g] = pack_cand(beamScore, left, bi);
      float s0 = eosLocal[g]; /* rawScores.beamLeft(beam, left)[0] */
      float localScore = 0;
      localScore += s0 * w0;
      localScore += s->sortScore[g] * w1;
      loc[g] = s0;
      full[g] = localScore + beamScore;
    }
    /* keep rnn scores aligned with the new order */
    float rnnS[64];
    for (u32 g = 0; g < eosG; ++g) rnnS[g] = s->sortScore[g];
    int elen = sort_for_beam(s->idxbuf, (int)eosG, full, maxBeam);
    jpp_b200_beam_element* row = &beams[(size_t)eosNode * c->beam];
    float* srow = &bscores[(size_t)eosNode * c->beam * c->S];
    int k = 0;
    for (; k < elen && k < maxBeam; ++k) {
      u32 g = s->idxbuf[k];
      row[k].left = (u16)(gb2[g] >> 16);
      row[k].prev_beam = (u16)gb2[g];
      row[k].total_score = full[g];
      srow[(size_t)k * c->S] = loc[g];
      srow[(size_t)k * c->S + 1] = rnnS[g];
    }
    for (; k < maxBeam; ++k) {
      set_fake(&row[k]);
      srow[(size_t)k * c->S] = 0; srow[(size_t)k * c->S + 1] = 0;
    }
  }
  if (c->stats) {
    __sync_fetch_and_add(&c->stats->gathers, gathers);
  }
  return 0;
}

/* -------------------------------------------------------------- driver ---- */
typedef struct {
  Ctx ctx;
  volatile u32* next;
  u32 n;
  u32 maxR, maxL, maxNb;
  int rc;
} Worker;

static void* alloc_or_die(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) abort();
  return p;
}

static void* worker_main(void* arg) {
  Worker* wk = (Worker*)arg;
  const Ctx* c = &wk->ctx;
  const CModel* cm = c->cm;
  Scratch s;
  memset(&s, 0, sizeof s);
  u32 G = c->cfg->global_beam_size > 0 ? (u32)c->cfg->global_beam_size : 1;
  if (G > 64) { wk->rc = -3; return NULL; }
  u32 chk = c->cfg->right_gbeam_check > 0 ? (u32)c->cfg->right_gbeam_check : 1;
  s.maxR = wk->maxR ? wk->maxR : 1;
  s.maxL = wk->maxL ? wk->maxL : 1;
  s.G = G;
  s.pats = alloc_or_die(sizeof(u64) * (size_t)s.maxR * cm->m->num_patterns);
  s.t0 = alloc_or_die(sizeof(float) * s.maxR);
  s.biSt = alloc_or_die(sizeof(u64) * (size_t)s.maxR * (cm->nBi + 1));
  s.triSt = alloc_or_die(sizeof(u64) * (size_t)s.maxR * (cm->nTri + 1));
  s.triSt2 = alloc_or_die(sizeof(u64) * (size_t)s.maxR * (cm->nTri + 1));
  s.pre = alloc_or_die(sizeof(float) * (size_t)chk * s.maxR);
  s.cutoff = alloc_or_die(sizeof(float) * (s.maxR > 64 ? s.maxR : 64));
  s.cutIdx = alloc_or_die(sizeof(u32) * s.maxR);
  s.cand = alloc_or_die(sizeof(u64) * ((size_t)s.maxL * c->beam + 64));
  s.t1pat = alloc_or_die(sizeof(u64) * (size_t)G * c->stored);
  s.t2pat = alloc_or_die(sizeof(u64) * (size_t)G * c->stored);
  s.t1idx = alloc_or_die(sizeof(u32) * G);
  s.t1node = alloc_or_die(sizeof(u32) * G);
  s.result = alloc_or_die(sizeof(float) * 64);
  s.scbuf = alloc_or_die(sizeof(float) * (s.maxL > 64 ? s.maxL : 64));
  s.T1 = alloc_or_die(sizeof(float) * s.maxR);
  s.T2 = alloc_or_die(sizeof(float) * s.maxR);
  s.fullScores = c->cfg->global_beam_size > 0 ? NULL : alloc_or_die(sizeof(float) * (size_t)s.maxR * c->beam * s.maxL);
  s.idxbuf = alloc_or_die(sizeof(u32) * (s.maxR > 256 ? s.maxR : 256));
  s.sortScore = alloc_or_die(sizeof(float) * 64);
  s.gbCap = (wk->maxNb + 2) * G;
  s.gbOff = alloc_or_die(sizeof(u32) * (wk->maxNb + 2));
  s.gbNode = alloc_or_die(sizeof(u32) * s.gbCap);
  s.gbSlot = alloc_or_die(sizeof(u16) * s.gbCap);
  s.gbBnd = alloc_or_die(sizeof(u16) * s.gbCap);
  for (;;) {
    u32 i = __sync_fetch_and_add(wk->next, 1);
    if (i >= wk->n) break;
    int rc = score_sentence(c, &s, i);
    if (rc != 0) { wk->rc = rc; break; }
  }
  free(s.pats); free(s.t0); free(s.biSt); free(s.triSt); free(s.triSt2); free(s.pre); free(s.cutoff); free(s.cutIdx);
  free(s.cand); free(s.t1pat); free(s.t2pat); free(s.t1idx); free(s.t1node); free(s.result); free(s.scbuf); free(s.T1);
  free(s.T2); free(s.fullScores); free(s.idxbuf); free(s.sortScore); free(s.gbOff); free(s.gbNode); free(s.gbSlot); free(s.gbBnd);
  return NULL;
}

int jpp_oracle_score_batch(const jpp_b200_model* model, const jpp_b200_config* cfg, const jpp_b200_batch* batch,
                           jpp_b200_result* result, const float* weights, int num_threads, jpp_oracle_stats* stats) {
  if (!model || !cfg || !batch || !result || !weights) return -1;
  if (!result->node_beam || !result->node_patterns) return -1;
  if (cfg->beam_size <= 0 || cfg->beam_size > 64) return -1;
  if (cfg->global_beam_size > 64) return -1;
  if (cfg->right_gbeam_check > 0 && cfg->right_gbeam_size <= 0) return -1; /* analyzer_impl.cc:58-61 */
  if (model->num_scorers == 2 && cfg->global_beam_size <= 0) return -1;    /* :83-88 */
  for (u32 i = 0; i < model->num_primitives; ++i) {
    int k = model->primitives[i].kind;
    if (k == JPP_B200_PRIM_BYTE_LENGTH || k == JPP_B200_PRIM_CODEPOINT_SIZE) return -4;
  }
  CModel cm;
  if (compile_model(model, weights, &cm) != 0) return -1;
  if (stats) memset(stats, 0, sizeof *stats);
  u32 maxR = 0, maxL = 0, maxNb = 0;
  u32 totalB = batch->sent_boundary_off[batch->num_sentences];
  for (u32 b = 0; b < totalB; ++b) {
    u32 R = batch->boundary_node_off[b + 1] - batch->boundary_node_off[b];
    u32 L = batch->boundary_end_off[b + 1] - batch->boundary_end_off[b];
    if (R > maxR) maxR = R;
    if (L > maxL) maxL = L;
  }
  for (u32 s = 0; s < batch->num_sentences; ++s) {
    u32 nb = batch->sent_boundary_off[s + 1] - batch->sent_boundary_off[s];
    if (nb > maxNb) maxNb = nb;
  }
  if (num_threads < 1) num_threads = 1;
  if (num_threads > 256) num_threads = 256;
  volatile u32 next = 0;
  Worker* wk = (Worker*)calloc((size_t)num_threads, sizeof(Worker));
  pthread_t* th = (pthread_t*)calloc((size_t)num_threads, sizeof(pthread_t));
  for (int t = 0; t < num_threads; ++t) {
    wk[t].ctx.cm = &cm;
    wk[t].ctx.cfg = cfg;
    wk[t].ctx.b = batch;
    wk[t].ctx.r = result;
    wk[t].ctx.stats = stats;
    wk[t].ctx.beam = (u32)cfg->beam_size;
    wk[t].ctx.S = model->num_scorers;
    wk[t].ctx.stored = model->num_stored_patterns;
    wk[t].next = &next;
    wk[t].n = batch->num_sentences;
    wk[t].maxR = maxR;
    wk[t].maxL = maxL;
    wk[t].maxNb = maxNb;
  }
  int rc = 0;
  if (num_threads == 1) {
    worker_main(&wk[0]);
    rc = wk[0].rc;
  } else {
    for (int t = 0; t < num_threads; ++t) pthread_create(&th[t], NULL, worker_main, &wk[t]);
    for (int t = 0; t < num_threads; ++t) {
      pthread_join(th[t], NULL);
      if (wk[t].rc) rc = wk[t].rc;
    }
  }
  free(wk);
  free(th);
  if (stats) {
    stats->sentences = batch->num_sentences;
    stats->boundaries = totalB;
    stats->nodes = batch->boundary_node_off[totalB];
  }
  return rc;
}

/* test hooks: the building blocks the reference's own unit tests pin
 * (score_processor_test.cc:48-113, perceptron_test.cc:11-44) */
uint64_t jpp_oracle_pack_candidate(float score, uint16_t left, uint16_t beam) { return pack_cand(score, left, beam); }
float jpp_oracle_candidate_score(uint64_t packed) { return cand_score(packed); }
float jpp_oracle_unrolled4(const float* weights, const uint32_t* idx, int n) { return unrolled4(weights, idx, n); }
uint64_t jpp_oracle_mix(uint64_t state, uint64_t data) { return mix(state, data); }

void jpp_oracle_fill_weights(float* out, uint32_t log2_size, uint64_t seed, float scale) {
  uint64_t n = 1ull << log2_size;
  for (uint64_t i = 0; i < n; ++i) out[i] = jpp_oracle_synth_weight(seed, i, scale);
}
