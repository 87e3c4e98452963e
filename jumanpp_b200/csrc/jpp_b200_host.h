// jpp_b200_host.h -- host-side compilation of a jpp_b200_model into the
// device feature tables (plain C++, no CUDA).  The constant hash prefixes are
// folded here: Hasher{}.mix(order+2).mix(index).mix(seed) for n-grams
// (reference impl/feature_impl_ngram_partial.h:29-41,58-100,123-178) and
// Hasher{}.mix(index).mix(nrefs).mix(PatternFeatureSeed) for patterns
// (impl/feature_impl_pattern.h:28-33).
#pragma once
#include <string>

#include "jpp_b200.h"
#include "jpp_b200_kernels.h"

#define JPP_STR2(x) #x
#define JPP_STR(x) JPP_STR2(x)

namespace jppb200 {

struct HostModel {
  FeatTables ft;
};

inline jpp_b200_status compile_host_model(const jpp_b200_model* m, HostModel* out, std::string* err) {
  auto bad = [&](jpp_b200_status st, const char* msg) { if (err) *err = msg; return st; };
  if (m->abi_version != JPP_B200_ABI_VERSION) return bad(JPP_B200_INVALID_PARAMETER, "jpp_b200_model.abi_version mismatch");
  if (m->log2_weights < 1 || m->log2_weights > 31) return bad(JPP_B200_INVALID_PARAMETER, "perceptron: size exponent out of range");
  if (m->num_scorers != 1 && m->num_scorers != 2) return bad(JPP_B200_INVALID_PARAMETER, "num_scorers must be 1 or 2");
  if ((m->num_scorers == 2) != (m->rnn != nullptr)) return bad(JPP_B200_INVALID_PARAMETER, "num_scorers == 2 needs an rnn model (and only then)");
  if (m->num_primitives > MAX_PRIM || m->num_computations > MAX_COMP || m->num_patterns > MAX_PAT ||
      m->branch_pool_size > MAX_POOL || m->pattern_ref_pool_size > MAX_POOL)
    return bad(JPP_B200_NOT_IMPLEMENTED, "feature program larger than the device tables");
  if (m->num_stored_patterns > m->num_patterns) return bad(JPP_B200_INVALID_PARAMETER, "num_stored_patterns > num_patterns");
  if (m->entry_size == 0 || m->entry_size > 64) return bad(JPP_B200_INVALID_PARAMETER, "entry_size out of range");
  FeatTables& ft = out->ft;
  memset(&ft, 0, sizeof ft);
  ft.nPrim = m->num_primitives; ft.nComp = m->num_computations; ft.nPat = m->num_patterns;
  ft.nStored = m->num_stored_patterns; ft.entrySize = m->entry_size; ft.nPlaceholders = m->num_placeholders;
  ft.mask = (u32)((1ull << m->log2_weights) - 1);
  if (m->float_order != JPP_B200_ORDER_STATIC_CODEGEN && m->float_order != JPP_B200_ORDER_DYNAMIC)
    return bad(JPP_B200_INVALID_PARAMETER, "unknown float_order");
  ft.dynamicOrder = m->float_order == JPP_B200_ORDER_DYNAMIC ? 1u : 0u;
  for (u32 i = 0; i < m->num_primitives; ++i) {
    const jpp_b200_primitive& p = m->primitives[i];
    switch (p.kind) {
      case JPP_B200_PRIM_COPY: case JPP_B200_PRIM_SHIFT_MASK:
        if (p.a < 0 || (u32)p.a >= m->entry_size) return bad(JPP_B200_INVALID_PARAMETER, "primitive references a field outside the entry");
        break;
      case JPP_B200_PRIM_PROVIDED:
        if (p.a < 0 || (u32)p.a >= m->num_placeholders) return bad(JPP_B200_INVALID_PARAMETER, "primitive references a missing placeholder");
        break;
      case JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE: case JPP_B200_PRIM_CODEPOINT_TYPE: case JPP_B200_PRIM_CODEPOINT: break;
      case JPP_B200_PRIM_BYTE_LENGTH: case JPP_B200_PRIM_CODEPOINT_SIZE:
        return bad(JPP_B200_NOT_IMPLEMENTED, "byte/codepoint length of a dictionary field needs the dictionary string storage; not staged");
      default: return bad(JPP_B200_INVALID_PARAMETER, "unknown primitive kind");
    }
    ft.primKind[i] = p.kind; ft.primA[i] = p.a; ft.primB[i] = p.b; ft.primC[i] = p.c;
  }
  const int32_t* bp = m->branch_pool;
  for (u32 i = 0; i < m->branch_pool_size; ++i) {
    if (bp[i] < 0 || (u32)bp[i] >= m->num_primitives) return bad(JPP_B200_INVALID_PARAMETER, "branch pool entry out of range");
    ft.branchPool[i] = (u8)bp[i];
  }
  for (u32 i = 0; i < m->num_computations; ++i) {
    const jpp_b200_computation& c = m->computations[i];
    if (c.primitive < 0 || (u32)c.primitive >= m->num_primitives || c.true_off + c.true_len > m->branch_pool_size ||
        c.false_off + c.false_len > m->branch_pool_size || c.true_len > 255 || c.false_len > 255)
      return bad(JPP_B200_INVALID_PARAMETER, "computation feature out of range");
    ft.compPrim[i] = c.primitive;
    ft.compTrueOff[i] = (u16)c.true_off; ft.compTrueLen[i] = (u8)c.true_len;
    ft.compFalseOff[i] = (u16)c.false_off; ft.compFalseLen[i] = (u8)c.false_len;
  }
  const int32_t* rp = m->pattern_ref_pool;
  for (u32 i = 0; i < m->pattern_ref_pool_size; ++i) {
    if (rp[i] < 0 || (u32)rp[i] >= m->num_computations) return bad(JPP_B200_INVALID_PARAMETER, "pattern reference out of range");
    ft.patRefPool[i] = (u8)rp[i];
  }
  for (u32 i = 0; i < MAX_PAT; ++i) ft.patOfSlot[i] = 0xFF;
  for (u32 p = 0; p < m->num_patterns; ++p) {
    const jpp_b200_pattern& pt = m->patterns[p];
    if (pt.index >= m->num_patterns || pt.ref_off + pt.ref_len > m->pattern_ref_pool_size || pt.ref_len > 255)
      return bad(JPP_B200_INVALID_PARAMETER, "pattern feature out of range");
    ft.patOfSlot[pt.index] = (u8)p;  // a later pattern with the same slot wins, as in the array form
    ft.patIndex[p] = (u8)pt.index; ft.patRefOff[p] = (u16)pt.ref_off; ft.patRefLen[p] = (u8)pt.ref_len;
    ft.patState[p] = mix(mix(mix(kSeed0, pt.index), pt.ref_len), kPatternSeed);
  }
  for (u32 i = 0; i < m->num_ngrams; ++i) {
    const jpp_b200_ngram& g = m->ngrams[i];
    if (g.order < 1 || g.order > 3) return bad(JPP_B200_INVALID_STATE, "invalid ngram feature: only orders 1-3 are supported");
    for (u32 k = 0; k < g.order; ++k)
      if (g.refs[k] >= m->num_patterns || (k > 0 && g.refs[k] >= m->num_stored_patterns))
        return bad(JPP_B200_INVALID_PARAMETER, "ngram references a pattern that is not stored per node");
    const u64 s = mix(mix(kSeed0, g.order + 2), g.index);
    if (g.order == 1) {
      if (ft.nUni >= MAX_UNI) return bad(JPP_B200_NOT_IMPLEMENTED, "too many unigram features");
      ft.uniState[ft.nUni] = mix(s, kUnigramSeed); ft.uniT0[ft.nUni] = (u8)g.refs[0]; ++ft.nUni;
    } else if (g.order == 2) {
      if (ft.nBi >= MAX_BI) return bad(JPP_B200_NOT_IMPLEMENTED, "too many bigram features");
      if (g.refs[0] >= m->num_stored_patterns) return bad(JPP_B200_INVALID_PARAMETER, "bigram t0 pattern is not stored per node");
      ft.biState[ft.nBi] = mix(s, kBigramSeed); ft.biT0[ft.nBi] = (u8)g.refs[0]; ft.biT1[ft.nBi] = (u8)g.refs[1]; ++ft.nBi;
    } else {
      if (ft.nTri >= MAX_TRI) return bad(JPP_B200_NOT_IMPLEMENTED, "too many trigram features");
      if (g.refs[0] >= m->num_stored_patterns) return bad(JPP_B200_INVALID_PARAMETER, "trigram t0 pattern is not stored per node");
      ft.triState[ft.nTri] = mix(s, kTrigramSeed);
      ft.triT0[ft.nTri] = (u8)g.refs[0]; ft.triT1[ft.nTri] = (u8)g.refs[1]; ft.triT2[ft.nTri] = (u8)g.refs[2]; ++ft.nTri;
    }
  }
  // accumulator counts of the generated code (pattern_feature_codegen.cc:202-208,
  // partial_ngram_feature_codegen.cc:15-22,86-93)
  ft.uniVars = ft.nUni >= 8 ? 4 : (ft.nUni >= 4 ? 2 : 1);
  ft.biVars = ft.nBi >= 8 ? 8 : (ft.nBi >= 4 ? 4 : (ft.nBi >= 2 ? 2 : 1));
  ft.triVars = ft.nTri >= 8 ? 8 : (ft.nTri >= 4 ? 4 : (ft.nTri >= 2 ? 2 : 1));
  if (m->rnn) {
    const jpp_b200_rnn_model* r = m->rnn;
    if (!r->matrix || !r->embeddings || !r->nce_embeddings || !r->maxent) return bad(JPP_B200_INVALID_PARAMETER, "rnn model has null blocks");
    if (r->layer_size == 0 || r->layer_size > 65536) return bad(JPP_B200_NOT_IMPLEMENTED, "we don't support embed sizes > 64k");
    if (r->maxent_order > 4) return bad(JPP_B200_NOT_IMPLEMENTED, "maxent order above 4 is not supported");
    if (r->maxent_size <= r->vocab_size) return bad(JPP_B200_INVALID_PARAMETER, "maxent table smaller than the vocabulary");
  }
  return JPP_B200_OK;
}

}  // namespace jppb200
