/*
 * jpp_b200.h -- C ABI of the B200-native Juman++ lattice scorer.
 *
 * This is the drop-in boundary for ONE path of ku-nlp/jumanpp: the per-sentence
 * scoring loop
 *     core::analysis::AnalyzerImpl::computeScores(const ScorerDef*)
 *         reference src/core/analysis/analyzer_impl.cc:299-309 (-> :250-297 gbeam,
 *         :197-248 full beam)
 * i.e. everything between "lattice built, BOS beams bootstrapped"
 * (analyzer_impl.cc:141-195) and "beams / scores filled, EOS beam[0] = 1-best"
 * (consumed by analysis_result.cc:25-76).  The reference has no FFI for this
 * path (it is in-process C++), so the entry points below are what a binding of
 * that seam needs; each one cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no CUDA / torch types in any signature; every call runs on streams
 *     the handle owns and returns when its work is complete;
 *   - all buffers are caller-owned host memory (pinned memory makes the copies
 *     asynchronous, pageable memory works too); the handle owns device copies;
 *   - every function returns a jpp_b200_status; text via jpp_b200_last_error();
 *     the codes mirror jumanpp::Status (reference src/util/status.hpp:17-23);
 *   - there is NO CPU fallback: if no CUDA device is usable jpp_b200_create
 *     fails with JPP_B200_CUDA_ERROR.
 */
#ifndef JPP_B200_H
#define JPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPP_B200_ABI_VERSION 3u

typedef enum jpp_b200_status {
  JPP_B200_OK = 0,
  JPP_B200_INVALID_PARAMETER = 1, /* Status::InvalidParameter */
  JPP_B200_INVALID_STATE = 2,     /* Status::InvalidState     */
  JPP_B200_NOT_IMPLEMENTED = 3,   /* Status::NotImplemented   */
  JPP_B200_CUDA_ERROR = 4         /* any CUDA runtime failure (maps to InvalidState) */
} jpp_b200_status;

/* ------------------------------------------------------------------ model -- */

/* Primitive features: reference spec::PrimitiveFeatureKind
 * (src/core/spec/spec_types.h:111-121) with the argument meaning of the
 * implementations in src/core/impl/feature_impl_prim.h:62-236. */
typedef enum jpp_b200_prim_kind {
  JPP_B200_PRIM_COPY = 1,          /* a = entry field           (:62-76)   */
  JPP_B200_PRIM_SHIFT_MASK = 2,    /* a = field, b = shift, c = mask (:78-96) */
  JPP_B200_PRIM_PROVIDED = 3,      /* a = placeholder index     (:98-113)  */
  JPP_B200_PRIM_BYTE_LENGTH = 4,   /* needs dictionary strings: rejected     */
  JPP_B200_PRIM_CODEPOINT_SIZE = 5,/* needs dictionary strings: rejected     */
  JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE = 6, /* end - start       (:157-170) */
  JPP_B200_PRIM_CODEPOINT_TYPE = 7,/* a = offset                (:203-236) */
  JPP_B200_PRIM_CODEPOINT = 8      /* a = offset                (:172-201) */
} jpp_b200_prim_kind;

typedef struct jpp_b200_primitive {
  int32_t kind; /* jpp_b200_prim_kind */
  int32_t a, b, c;
} jpp_b200_primitive;

/* Computation feature: reference spec::ComputationFeatureDescriptor
 * (spec_types.h:141-147), semantics of feature_impl_compute.cc:12-27,59-63:
 * both branches empty -> mix(primitive); else primitive != 0 selects the list
 * of primitives to mix. Branch lists index into jpp_b200_model::branch_pool. */
typedef struct jpp_b200_computation {
  int32_t primitive;
  uint32_t true_off, true_len;
  uint32_t false_off, false_len;
} jpp_b200_computation;

/* Pattern feature: spec::PatternFeatureDescriptor (spec_types.h:149-153);
 * value = Hasher{}.mix(index).mix(n_refs).mix(0x7a11ed00000000) then each
 * referenced computation (feature_impl_pattern.h:28-39). `refs` index
 * jpp_b200_model::pattern_ref_pool. */
typedef struct jpp_b200_pattern {
  uint32_t index;
  uint32_t ref_off, ref_len;
} jpp_b200_pattern;

/* N-gram feature: spec::NgramFeatureDescriptor (spec_types.h:155-158) as
 * consumed by feature_impl_ngram_partial.cc:12-37: refs[0] is the t0 (right
 * node) pattern, refs[1] t1 (left node), refs[2] t2 (node before the left). */
typedef struct jpp_b200_ngram {
  uint32_t order; /* 1, 2 or 3 */
  uint32_t index; /* NgramFeatureDescriptor::index (hashed) */
  uint32_t refs[3];
} jpp_b200_ngram;

/* RNN language model: the four float blocks of
 * RnnScorerGbeamFactory::load (rnn_scorer_gbeam.cc:386-396) plus the inference
 * constants of RnnInferenceConfig (rnn_scorer.h:19-28). Optional. */
typedef struct jpp_b200_rnn_model {
  uint32_t layer_size;   /* H   : MikolovRnnModelHeader::layerSize */
  uint32_t maxent_order; /*       ::maxentOrder (<= 4)             */
  uint64_t maxent_size;  /*       ::maxentSize                     */
  uint64_t vocab_size;   /*       ::vocabSize                      */
  float nce_constant;    /* MikolovRnn::nceConstant() after setConfig/load */
  int32_t unk_id;        /* RnnIdResolver::unkId()                 */
  float unk_constant;    /* unkConstantTerm                        */
  float unk_length_penalty; /* unkLengthPenalty                    */
  const float* matrix;        /* [H*H]  column-major, ctx'_i = sum_k W[k + i*H]*ctx_k */
  const float* embeddings;    /* [vocab][H] */
  const float* nce_embeddings;/* [vocab][H] */
  const float* maxent;        /* [maxent_size] */
} jpp_b200_rnn_model;

/* Everything CoreHolder/ScorerDef give the reference's ScoreProcessor:
 * compiled FeaturesSpec (spec_types.h:160-171), perceptron table
 * (perceptron.cc:81-124) and scorer weights (score_api.h:66-72). */
typedef struct jpp_b200_model {
  uint32_t abi_version; /* JPP_B200_ABI_VERSION */
  /* hashed perceptron: float[1 << log2_weights]; if `weights` is NULL the table
   * is synthesised on the device from (weight_seed, weight_scale) with the
   * counter-based generator documented in DESIGN.md (bench / tests only). */
  const float* weights;
  uint32_t log2_weights;
  uint64_t weight_seed;
  float weight_scale;
  /* dictionary entry row width (DictionaryEntries::numFeatures(), 8 for
   * Jumandic) and FeaturesSpec::numPlaceholders */
  uint32_t entry_size;
  uint32_t num_placeholders;
  /* feature program */
  uint32_t num_primitives;
  const jpp_b200_primitive* primitives;
  uint32_t num_computations;
  const jpp_b200_computation* computations;
  uint32_t branch_pool_size;
  const int32_t* branch_pool;
  uint32_t num_patterns;          /* FeaturesSpec::pattern.size()  (38) */
  uint32_t num_stored_patterns;   /* pattern.size() - numUniOnlyPats (14): patterns
                                     with index < this are kept per node */
  const jpp_b200_pattern* patterns;
  uint32_t pattern_ref_pool_size;
  const int32_t* pattern_ref_pool;
  uint32_t num_ngrams;            /* in FeaturesSpec::ngram order */
  const jpp_b200_ngram* ngrams;
  /* scorers: ScorerDef::scoreWeights; num_scorers = 1 (perceptron only) or 2 */
  uint32_t num_scorers;
  float score_weights[2];
  const jpp_b200_rnn_model* rnn;  /* NULL when num_scorers == 1 */
  /* Which of the reference's two feature implementations the float sums replay
   * (jpp_b200_float_order).  The CLI runs the code GENERATED for its spec
   * (StaticFeatureFactory, core.cc:18-22): per-row accumulator interleavings of
   * pattern_feature_codegen.cc:199-313 / partial_ngram_feature_codegen.cc:13-153,
   * last row of a block through computeUnrolled4RawPerceptron.  A model without
   * generated code runs PartialNgramDynamicFeatureApply
   * (feature_impl_ngram_partial.h:196-371): every row through
   * computeUnrolled4RawPerceptron.  The sums differ in the last bits, and with
   * them which of two near-equal candidates wins. */
  uint32_t float_order;
} jpp_b200_model;

typedef enum jpp_b200_float_order {
  JPP_B200_ORDER_STATIC_CODEGEN = 0,
  JPP_B200_ORDER_DYNAMIC = 1
} jpp_b200_float_order;

/* ----------------------------------------------------------------- config -- */

/* AnalyzerConfig + LatticeConfig fields the scoring loop reads
 * (analyzer.h:14-25, lattice_config.h:62-68; CLI defaults jumanpp_args.h:50-54:
 * beam 5, global beam 6, right check 1, right beam 5). */
typedef struct jpp_b200_config {
  int32_t beam_size;          /* LatticeConfig::beamSize            */
  int32_t global_beam_size;   /* AnalyzerConfig::globalBeamSize; <= 0 selects the
                                 full-beam loop (computeScoresFull)  */
  int32_t right_gbeam_check;  /* AnalyzerConfig::rightGbeamCheck    */
  int32_t right_gbeam_size;   /* AnalyzerConfig::rightGbeamSize     */
} jpp_b200_config;

/* ------------------------------------------------------------------ batch -- */

/* One LatticeNodePtr (lattice_config.h:21-24): a node that ENDS at a boundary. */
typedef struct jpp_b200_node_ptr {
  uint16_t boundary;
  uint16_t position;
} jpp_b200_node_ptr;

/* NodeInfo (core_types.h:71-90). */
typedef struct jpp_b200_node_info {
  int32_t entry_ptr;
  uint16_t start;
  uint16_t end;
} jpp_b200_node_info;

/* chars::InputCodepoint as read by the codepoint features. */
typedef struct jpp_b200_codepoint {
  uint32_t codepoint;
  uint32_t char_class;
} jpp_b200_codepoint;

/* A batch of N built lattices in CSR form (the "pre-state" of computeScores).
 * Boundaries are numbered as in the reference: 0 and 1 are the two BOS
 * boundaries (one node each), 2..n-2 the codepoint boundaries, n-1 the EOS
 * boundary (one node) -- Lattice::createdBoundaryCount() = codepoints + 3
 * (analyzer_impl.cc:141-177). */
typedef struct jpp_b200_batch {
  uint32_t num_sentences;
  /* [num_sentences + 1] first boundary of sentence s in the boundary arrays */
  const uint32_t* sent_boundary_off;
  /* [num_sentences + 1] first codepoint of sentence s */
  const uint32_t* sent_codepoint_off;
  /* [total_boundaries + 1] first node starting at the boundary (node index) */
  const uint32_t* boundary_node_off;
  /* [total_boundaries + 1] first entry of the boundary in `ends` */
  const uint32_t* boundary_end_off;
  /* per node (LatticeRightBoundary, lattice_types.h:44-67) */
  const jpp_b200_node_info* node_info;  /* [total_nodes]                     */
  const int32_t* node_entry;            /* [total_nodes][entry_size], filled as
                                           PrimitiveFeatureContext::fillEntryBuffer
                                           does (feature_impl_types.h:128-148) */
  const int32_t* node_placeholders;     /* [total_nodes][num_placeholders] =
                                           providedFeature() (:108-119); may be
                                           NULL when num_placeholders == 0   */
  const int32_t* node_rnn_id;           /* [total_nodes] RnnIdContainer::resolveId
                                           (rnn_id_resolver.cc:278-309); NULL
                                           without an RNN                    */
  /* per boundary, LatticeLeftBoundary::nodePtrs() (lattice_types.h:19-42) */
  const jpp_b200_node_ptr* ends;        /* [total_ends]                      */
  /* per sentence codepoints */
  const jpp_b200_codepoint* codepoints; /* [total_codepoints]                */
  /* Compact alternative to node_entry (ABI 2).  When node_entry is NULL the entry
   * buffer of node n is entry_rows[node_entry_row[n]]: the stager keeps ONE row per
   * distinct entry of the batch -- a dictionary word's EntryPtr identifies its
   * row (DictionaryEntries::entryAtPtr, dic_entries.h), unknown-word nodes add
   * rows of their own -- and sends 4 bytes per node instead of 4*entry_size.
   * The rows hold exactly what fillEntryBuffer would have written. */
  const uint32_t* node_entry_row;       /* [total_nodes] or NULL              */
  const int32_t* entry_rows;            /* [num_entry_rows][entry_size]       */
  uint32_t num_entry_rows;
  /* Everything else that is a function of the entry can ride in the row as well:
   * providedFeature() is 0 for dictionary words and set by the unk makers for
   * unknown-word nodes (which own their rows), and RnnIdContainer::resolveId
   * depends on the entry's fields / the unk node's surface only.  Used, in the
   * compact form, when the per-node arrays above are NULL. */
  const int32_t* row_placeholders;      /* [num_entry_rows][num_placeholders] or NULL */
  const int32_t* row_rnn_id;            /* [num_entry_rows] or NULL                   */
  /* Resident form (ABI 3): node_entry AND node_entry_row are NULL.  The entry
   * buffer of a dictionary node (node_info.entry_ptr >= 0) comes from the table the handle
   * holds on the device (jpp_b200_add_entry_rows): the host stages NOTHING per dictionary
   * node beyond its 8-byte NodeInfo -- no fillEntryBuffer, no RNN id lookup.  Only the nodes
   * the analyzer made up for this sentence (unknown words etc.: entry_ptr < 0, EntryPtr::extPtr()
   * = ~entry_ptr is their index in the sentence's ExtraNodesContext, core_types.h:38-41) bring
   * rows: extra node i of sentence s is entry_rows[sent_extra_row_off[s] + i] (with
   * row_placeholders / row_rnn_id beside it).  BOS / EOS entries need no row: every field of
   * theirs is the entry_ptr itself (feature_impl_types.h:130-133). */
  const uint32_t* sent_extra_row_off;   /* [num_sentences + 1] or NULL                */
} jpp_b200_batch;

/* Pre-decoded dictionary entries: what PrimitiveFeatureContext::fillEntryBuffer
 * (feature_impl_types.h:128-148) and RnnIdContainer::resolveId (rnn_id_resolver.cc:278-309)
 * return for a dictionary EntryPtr depends on the dictionary alone, so it is computed once. */
typedef struct jpp_b200_entry_table {
  uint32_t num_rows;
  const int32_t* entry_ptr; /* [num_rows] EntryPtr::rawValue() (>= 0) of each row's entry      */
  const int32_t* rows;      /* [num_rows][entry_size]                                          */
  const int32_t* rnn_id;    /* [num_rows], NULL for a model without RNN                        */
} jpp_b200_entry_table;

/* ConnectionBeamElement (lattice_config.h:70-73) without the raw pointer: the
 * `previous` link is (boundary/position of ends[left], beam index `prev_beam`);
 * a fake (padding) element has left == prev_beam == 0xFFFF -- test that, not the
 * score: the reference pads with memset 0xFF (score_processor.cc:466-468, a NaN
 * bit pattern) in makeT0Beam and with lowest() in EntryBeam::fake(). */
typedef struct jpp_b200_beam_element {
  uint16_t left;      /* ConnectionPtr::left : index into the boundary's ends  */
  uint16_t prev_beam; /* ConnectionPtr::beam : beam slot of the previous node  */
  float total_score;  /* ConnectionBeamElement::totalScore                     */
} jpp_b200_beam_element;

/* The "post-state" of computeScores. Any pointer may be NULL to skip that
 * output. */
typedef struct jpp_b200_result {
  /* [total_nodes][beam_size] LatticeRightBoundary::beamData(), sorted
   * descending, fake-padded. The EOS node's row is the final (RNN re-ranked)
   * beam; element 0 is the 1-best. */
  jpp_b200_beam_element* node_beam;
  /* [total_nodes][beam_size][num_scorers] LatticeBoundaryScores::forPtr() of
   * the connection held in each beam slot (needed by --format=lattice and by
   * adjustBeamScores); for slots that are fake the values are 0. Only RNN
   * slots of connections on EOS-global-beam paths are defined, as in the
   * reference. */
  float* node_beam_scores;
  /* [total_nodes][num_stored_patterns] LatticeRightBoundary::patternFeatureData() */
  uint64_t* node_patterns;
  /* [total_nodes] ScoreProcessor scores_.bufferT0() (unigram score of the node) */
  float* node_t0_score;
  /* The 1-best path of every sentence, i.e. what AnalysisPath::fillIn
   * (analysis_result.cc:25-76) walks from EOS beam[0] through the `previous`
   * links, already resolved on the device: best_path_len[s] node indices (global
   * node index into the batch, BOS and EOS excluded, in sentence order) stored
   * at best_path_nodes[sent_boundary_off[s] ...].  A caller that only prints the
   * 1-best analysis (the CLI default) downloads these few MB instead of every
   * beam.  best_path_len: [num_sentences]; best_path_nodes: [total_boundaries],
   * unused slots of a sentence hold 0xFFFFFFFF. */
  uint32_t* best_path_len;
  uint32_t* best_path_nodes;
  /* [num_sentences] 0 = scored; otherwise a mask of jpp_b200_sentence_status bits
   * saying why the sentence was NOT scored.  The reference fails one sentence and
   * goes on (`# ERROR` + EOS, jumanpp.cc:166-171); so does the batch: a malformed
   * lattice is found by a validation kernel at upload, skipped by every scoring
   * kernel, and reported here with best_path_len = 0 (its other outputs are
   * undefined).  When this pointer is NULL the caller has not asked for
   * per-sentence status and a batch with a failed sentence returns
   * JPP_B200_INVALID_PARAMETER (the other sentences' outputs are still written). */
  uint8_t* sentence_status;
} jpp_b200_result;

typedef enum jpp_b200_sentence_status {
  JPP_B200_SENT_OK = 0,
  JPP_B200_SENT_BAD_OFFSETS = 1,  /* boundary_node_off / boundary_end_off / codepoint offsets not
                                     monotonic or past the arrays                             */
  JPP_B200_SENT_BAD_SHAPE = 2,    /* fewer than 3 or more than 65535 boundaries, or not exactly
                                     one node on the BOS / BOS / EOS boundaries               */
  JPP_B200_SENT_BAD_ENDS = 4,     /* ends[] entry that names a later boundary or a node the
                                     boundary does not have                                   */
  JPP_B200_SENT_BAD_SPAN = 8,     /* node_info.start/end outside the sentence's codepoints     */
  JPP_B200_SENT_BAD_ENTRY = 16,   /* node_entry_row >= num_entry_rows; resident form: entry_ptr
                                     without a resident row / extra index past the sentence's rows */
  JPP_B200_SENT_BAD_RNN_ID = 32   /* RNN id outside [-1, vocab_size)                           */
} jpp_b200_sentence_status;

/* Work counters of the last scored batch (roofline accounting). */
typedef struct jpp_b200_counters {
  uint64_t sentences;
  uint64_t boundaries;
  uint64_t nodes;
  uint64_t gathers;        /* perceptron-table reads of the reference algorithm
                              (nU*R + c*R*(nB+nT) + K*(U*nB+(G-c)*nT)
                              per boundary; 4 B each)                          */
  uint64_t algorithmic_bytes; /* A of BASELINE.md "Roofline accounting"           */
  uint64_t rnn_nodes;      /* prefix-tree contexts computed (2*H*H flop each)   */
  uint64_t rnn_scores;     /* RNN scores computed (2*H flop + maxent_order reads each) */
  uint64_t walk_gathers;   /* the same for the wave loop alone (without T0)     */
  uint64_t walk_bytes;     /* algorithmic bytes of the beam-walk kernel alone   */
  uint64_t launches;       /* kernels launched by the last score call           */
  float kernel_ms[8];      /* per stage, CUDA events on the scoring stream:
                              0 T0, 1 beam walk, 2 rnn tree, 3 rnn contexts, 4 rnn
                              scores, 5 re-rank, 6 upload (H2D + index), 7 all scoring */
  uint64_t walk_gathers_issued; /* table reads the wave loop really made: the bigram
                              weights of (node, head element's t1) are read once
                              and summed in both orders, so this is smaller   */
  uint64_t failed_sentences;  /* sentences with a non-zero status                  */
} jpp_b200_counters;

typedef struct jpp_b200_handle jpp_b200_handle;

/* Replaces: JumanppEnv::loadModel + initFeatures + makeAnalyzer -> AnalyzerImpl
 * ctor + initScorers (env.cc:11-36, analyzer_impl.cc:27-95): uploads the model
 * once, fixes the scoring configuration.  Device buffers for batches grow on
 * demand and are kept between calls. */
jpp_b200_status jpp_b200_create(const jpp_b200_model* model,
                                const jpp_b200_config* config, int device,
                                jpp_b200_handle** out);

/* Replaces: AnalyzerImpl::setGlobalBeam / env.setBeamSize between sentences
 * (analyzer_impl.cc:311-331). */
jpp_b200_status jpp_b200_set_config(jpp_b200_handle* h,
                                    const jpp_b200_config* config);

/* Replaces: for each sentence of the batch, AnalyzerImpl::bootstrapAnalysis +
 * computeScores (analyzer_impl.cc:179-195, :299-309) including the
 * ScoreComputer (RNN) pass and adjustBeamScores/remakeEosBeam (:286-294).
 * Host buffers in, host buffers out; H2D/D2H copies are issued on the
 * handle's streams and the call returns after they complete.  A batch of zero
 * sentences is a no-op (JPP_B200_OK). */
jpp_b200_status jpp_b200_score_batch(jpp_b200_handle* h,
                                     const jpp_b200_batch* batch,
                                     jpp_b200_result* result);

/* Same work split for callers that keep batches resident on the device (bench,
 * pipelined staging): upload copies `batch` to HBM and returns; score runs the
 * kernels only; download copies the requested outputs back. */
jpp_b200_status jpp_b200_upload_batch(jpp_b200_handle* h,
                                      const jpp_b200_batch* batch);
jpp_b200_status jpp_b200_score_resident(jpp_b200_handle* h);
jpp_b200_status jpp_b200_download_result(jpp_b200_handle* h,
                                         jpp_b200_result* result);

/* Adds rows to the handle's resident entry table (device copy; the whole dictionary at
 * start-up, or the entries of each batch that were not seen before).  A row whose entry_ptr
 * is already resident replaces the old one.  Not to be called while another thread scores
 * on the same handle, unless the table was sized with jpp_b200_reserve_entry_rows. */
jpp_b200_status jpp_b200_add_entry_rows(jpp_b200_handle* h,
                                        const jpp_b200_entry_table* rows);
/* Sizes the resident table for `max_rows` rows and entry_ptr values below `entry_ptr_limit` once.
 * Within that reservation jpp_b200_add_entry_rows never moves the device arrays and runs on a
 * stream of its own, so a stager MAY add the rows of the next batch while jpp_b200_score_batch of
 * the previous one is still running on another thread (that batch only names rows added before
 * it started). */
jpp_b200_status jpp_b200_reserve_entry_rows(jpp_b200_handle* h, uint32_t max_rows,
                                            uint32_t entry_ptr_limit);

jpp_b200_status jpp_b200_get_counters(const jpp_b200_handle* h,
                                      jpp_b200_counters* out);

/* Diagnostic, not part of the scored path: the rate at which THIS device serves independent 4-byte reads
 * at random indices of the handle's perceptron table (about `gathers` of them, `ilp` = 4 / 8 / 16 in flight
 * per thread, `blocks_per_sm` blocks of 256 threads; 0 = defaults).  bench.py reports the scoring kernels'
 * table reads per second against it ("HBM-gather roofline"). */
jpp_b200_status jpp_b200_gather_probe(jpp_b200_handle* h, uint64_t gathers, uint32_t ilp,
                                      uint32_t blocks_per_sm, double* gathers_per_second);

/* Page-locked host memory for the staged batch / result arrays (the H2D / D2H copies of a
 * batch in pinned memory run at full PCIe rate and asynchronously).  The host side needs no
 * CUDA header: this library is the only thing it links. */
jpp_b200_status jpp_b200_alloc_pinned(size_t bytes, void** out);
void jpp_b200_free_pinned(void* p);

/* Replaces: Status::message(); valid until the next call on the handle.
 * With h == NULL returns the message of the last failed jpp_b200_create on
 * this thread. */
const char* jpp_b200_last_error(const jpp_b200_handle* h);

void jpp_b200_destroy(jpp_b200_handle* h);

/* Library self-description (tests check the loaded .so is the sm_100a build). */
uint32_t jpp_b200_abi_version(void);
const char* jpp_b200_build_info(void);

#ifdef __cplusplus
}
#endif

#endif /* JPP_B200_H */
