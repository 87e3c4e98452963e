// jpp_gpu_batch.cc -- see jpp_gpu_batch.h.  Reference files are cited where their behaviour is
// relied on; nothing here scores: the only scoring call is jpp_b200_score_batch.
#include "jpp_gpu_batch.h"

#include <algorithm>
#include <chrono>
#include <cstring>

#include "core/analysis/lattice_types.h"
#include "core/analysis/score_processor.h"
#include "core/impl/feature_impl_types.h"
#include "core/impl/model_io.h"
#include "core/impl/perceptron_io.h"
#include "rnn/mikolov_rnn.h"
#include "util/serialization.h"

namespace jumanpp {
namespace b200 {

using namespace core;
using namespace core::analysis;

namespace {

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Status gpuError(jpp_b200_handle* h, jpp_b200_status st, const char* what) {
  return JPPS_INVALID_STATE << what << ": " << jpp_b200_last_error(h) << " (jpp_b200 status " << (int)st << ")";
}

// the header RnnScorerGbeamFactory::makeInfo / load serialise into the RNN part of a model file
// (rnn_scorer_gbeam.cc:348-373: a struct private to that file; this is its field order)
struct RnnPartHeader {
  core::analysis::rnn::RnnInferenceConfig config;
  i32 unkIdx = 0;
  std::vector<u32> fields;
  jumanpp::rnn::mikolov::MikolovRnnModelHeader rnnHeader{};
};

template <typename Arch>
void Serialize(Arch& a, RnnPartHeader& o) {
  a& o.config.nceBias;
  a& o.config.unkConstantTerm;
  a& o.config.unkLengthPenalty;
  a& o.config.perceptronWeight;
  a& o.config.rnnWeight;
  a& o.config.eosSymbol;
  a& o.config.unkSymbol;
  a& o.config.rnnFields;
  a& o.config.fieldSeparator;
  a& o.unkIdx;
  a& o.fields;
  a& o.rnnHeader.layerSize;
  a& o.rnnHeader.maxentOrder;
  a& o.rnnHeader.maxentSize;
  a& o.rnnHeader.vocabSize;
  a& o.rnnHeader.nceLnz;
}

const EntryPtr kBos = EntryPtr::BOS(), kEos = EntryPtr::EOS();

}  // namespace

// ------------------------------------------------------------ ParallelRunner ----
ParallelRunner::ParallelRunner(u32 numThreads) {
  for (u32 t = 1; t < std::max(1u, numThreads); ++t) threads_.emplace_back([this, t] { worker(t); });
}

ParallelRunner::~ParallelRunner() {
  {
    std::lock_guard<std::mutex> lk(m_);
    stop_ = true;
  }
  cvStart_.notify_all();
  for (auto& t : threads_) t.join();
}

void ParallelRunner::drain(u32 id) {
  for (;;) {
    const size_t b = next_.fetch_add(chunk_);
    if (b >= count_) break;
    const size_t e = std::min(count_, b + chunk_);
    for (size_t i = b; i < e; ++i) (*fn_)(i, id);
  }
}

void ParallelRunner::worker(u32 id) {
  u64 seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(m_);
      cvStart_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
    }
    drain(id);
    {
      std::lock_guard<std::mutex> lk(m_);
      running_ -= 1;
    }
    cvDone_.notify_one();
  }
}

void ParallelRunner::run(size_t count, const std::function<void(size_t, u32)>& fn, size_t chunk) {
  if (count == 0) return;
  {
    std::lock_guard<std::mutex> lk(m_);
    fn_ = &fn;
    count_ = count;
    chunk_ = chunk ? chunk : std::max<size_t>(1, count / (size() * 8));
    next_.store(0);
    running_ = static_cast<u32>(threads_.size());
    generation_ += 1;
  }
  cvStart_.notify_all();
  drain(0);
  std::unique_lock<std::mutex> lk(m_);
  cvDone_.wait(lk, [&] { return running_ == 0; });
}

// ------------------------------------------------------------- BatchAnalyzer ----
BatchAnalyzer::BatchAnalyzer() = default;

BatchAnalyzer::~BatchAnalyzer() {
  if (handle_) jpp_b200_destroy(handle_);
}

Status BatchAnalyzer::initialize(const JumanppEnv& env, const AnalyzerConfig& acfg, const ScoringConfig& sconf,
                                 StringPiece rnnFile, const BatchOptions& opts) {
  env_ = &env;
  opts_ = opts;
  acfg_ = acfg;
  sconf_ = sconf;
  if (opts_.threads == 0) opts_.threads = std::max(1u, std::thread::hardware_concurrency());
  if (opts_.maxBatch == 0) opts_.maxBatch = 1;
  if (acfg.autoBeamStep > 0) return JPPS_NOT_IMPLEMENTED << "auto beam is not available on the GPU path";
  runner_.reset(new ParallelRunner(opts_.threads));
  missingPerWorker_.resize(runner_->size());
  // what Analyzer::initialize gives every analyzer: the scorer weights, but no ScoreComputer
  // instances -- nothing on this path runs them
  scorerDef_.feature = env.scorers()->feature;
  scorerDef_.scoreWeights = env.scorers()->scoreWeights;
  JPP_RETURN_IF_ERROR(buildModel(env, rnnFile));
  JPP_RETURN_IF_ERROR(makeAnalyzers());
  return Status::Ok();
}

Status BatchAnalyzer::makeAnalyzers() {
  for (auto& sl : slots_) {
    sl.impls.resize(opts_.maxBatch);
    sl.analyzers.resize(opts_.maxBatch);
    sl.statuses.clear();
    std::vector<Status> errs;
    for (u32 i = 0; i < opts_.maxBatch; ++i) {
      sl.statuses.emplace_back(Status::Ok());
      errs.emplace_back(Status::Ok());
    }
    sl.shapes.resize(opts_.maxBatch);
    runner_->run(opts_.maxBatch, [&](size_t i, u32) {
      sl.impls[i].reset(new AnalyzerImpl{env_->coreHolder(), sconf_, acfg_});
      sl.analyzers[i].reset(new Analyzer{});
      errs[i] = sl.analyzers[i]->initialize(sl.impls[i].get(), &scorerDef_);
    });
    for (auto& s : errs) JPP_RETURN_IF_ERROR(std::move(s));
  }
  return Status::Ok();
}

// jpp_b200_model from the loaded reference model: the compiled feature program
// (spec::FeaturesSpec), the perceptron table and the RNN blocks as they lie in the model file
Status BatchAnalyzer::buildModel(const JumanppEnv& env, StringPiece rnnFile) {
  auto& fs = env.spec().features;
  std::vector<jpp_b200_primitive> prims;
  for (auto& p : fs.primitive) {
    jpp_b200_primitive o{0, 0, 0, 0};
    using K = spec::PrimitiveFeatureKind;
    switch (p.kind) {
      case K::Copy: o.kind = JPP_B200_PRIM_COPY; o.a = p.references[0]; break;
      case K::SingleBit: o.kind = JPP_B200_PRIM_SHIFT_MASK; o.a = p.references[0]; o.b = p.references[1]; o.c = 1; break;
      case K::Provided: o.kind = JPP_B200_PRIM_PROVIDED; o.a = p.references[0]; break;
      case K::ByteLength: o.kind = JPP_B200_PRIM_BYTE_LENGTH; o.a = p.references[0]; break;
      case K::CodepointSize: o.kind = JPP_B200_PRIM_CODEPOINT_SIZE; o.a = p.references[0]; break;
      case K::SurfaceCodepointSize: o.kind = JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE; break;
      case K::CodepointType: o.kind = JPP_B200_PRIM_CODEPOINT_TYPE; o.a = p.references[0]; break;
      case K::Codepoint: o.kind = JPP_B200_PRIM_CODEPOINT; o.a = p.references[0]; break;
      default: return JPPS_INVALID_PARAMETER << "unknown primitive feature kind";
    }
    prims.push_back(o);
  }
  std::vector<jpp_b200_computation> comps;
  std::vector<int32_t> branchPool, patRefPool;
  for (auto& c : fs.computation) {
    jpp_b200_computation o{};
    o.primitive = c.primitiveFeature;
    o.true_off = (u32)branchPool.size();
    for (auto x : c.trueBranch) branchPool.push_back(x);
    o.true_len = (u32)c.trueBranch.size();
    o.false_off = (u32)branchPool.size();
    for (auto x : c.falseBranch) branchPool.push_back(x);
    o.false_len = (u32)c.falseBranch.size();
    comps.push_back(o);
  }
  std::vector<jpp_b200_pattern> pats;
  for (auto& p : fs.pattern) {
    jpp_b200_pattern o{};
    o.index = (u32)p.index;
    o.ref_off = (u32)patRefPool.size();
    for (auto x : p.references) patRefPool.push_back(x);
    o.ref_len = (u32)p.references.size();
    pats.push_back(o);
  }
  std::vector<jpp_b200_ngram> ngrams;
  for (auto& n : fs.ngram) {
    jpp_b200_ngram o{};
    o.order = (u32)n.references.size();
    o.index = (u32)n.index;
    for (size_t i = 0; i < n.references.size() && i < 3; ++i) o.refs[i] = (u32)n.references[i];
    ngrams.push_back(o);
  }

  auto info = env.modelInfoCopy();
  auto perc = info.firstPartOf(model::ModelPartKind::Perceprton);
  if (perc == nullptr || perc->data.size() != 2) return JPPS_INVALID_STATE << "model has no perceptron part";
  PerceptronInfo pi{};
  {
    util::serialization::Loader l{perc->data[0]};
    if (!l.load(&pi)) return JPPS_INVALID_STATE << "failed to read the perceptron header";
  }

  jpp_b200_model m{};
  m.abi_version = JPP_B200_ABI_VERSION;
  m.weights = reinterpret_cast<const float*>(perc->data[1].begin());
  m.log2_weights = pi.modelSizeExponent;
  if (perc->data[1].size() != (sizeof(float) << m.log2_weights)) return JPPS_INVALID_STATE << "perceptron table size mismatch";
  entrySize_ = (u32)env.coreHolder()->dic().entries().numFeatures();
  numPlaceholders_ = (u32)fs.numPlaceholders;
  m.entry_size = entrySize_;
  m.num_placeholders = numPlaceholders_;
  m.num_primitives = (u32)prims.size(); m.primitives = prims.data();
  m.num_computations = (u32)comps.size(); m.computations = comps.data();
  m.branch_pool_size = (u32)branchPool.size(); m.branch_pool = branchPool.data();
  m.num_patterns = (u32)pats.size(); m.patterns = pats.data();
  m.num_stored_patterns = (u32)(fs.pattern.size() - fs.numUniOnlyPats);
  storedPatterns_ = m.num_stored_patterns;
  m.pattern_ref_pool_size = (u32)patRefPool.size(); m.pattern_ref_pool = patRefPool.data();
  m.num_ngrams = (u32)ngrams.size(); m.ngrams = ngrams.data();
  // generated code linked and accepted by the spec hash (core.cc:18-22, features_api.cc:38-47)?
  // otherwise the reference runs its dynamic implementation, whose float sums associate differently
  m.float_order = env.coreHolder()->features().ngramPartialStatic ? JPP_B200_ORDER_STATIC_CODEGEN : JPP_B200_ORDER_DYNAMIC;
  auto& sw = env.scorers()->scoreWeights;
  numScorers_ = (u32)sw.size();
  m.num_scorers = numScorers_;
  m.score_weights[0] = sw[0];
  m.score_weights[1] = sw.size() > 1 ? sw[1] : 0.0f;

  jpp_b200_rnn_model rm{};
  jumanpp::rnn::mikolov::MikolovModelReader reader;
  if (numScorers_ == 2) {
    hasRnn_ = true;
    auto factory = static_cast<const RnnScorerGbeamFactory*>(env.scorers()->others.at(0));
    auto& cfg = factory->config();  // the configuration in force (model header merged with the command line)
    jumanpp::rnn::mikolov::MikolovRnnModelHeader hdr{};
    float nce;
    if (!rnnFile.empty()) {
      // --rnn-model: RnnScorerGbeamFactory::make (rnn_scorer_gbeam.cc:305-335)
      JPP_RETURN_IF_ERROR(reader.open(rnnFile));
      JPP_RETURN_IF_ERROR(reader.parse());
      hdr = reader.header();
      JPP_RETURN_IF_ERROR(resolver_.build(env.coreHolder()->dic(), cfg, reader.words()));
      rm.matrix = reader.rnnMatrix().data();
      rm.embeddings = reader.embeddings().data();
      rm.nce_embeddings = reader.nceEmbeddings().data();
      rm.maxent = reader.maxentWeights().data();
      nce = cfg.nceBias.isDefault() ? hdr.nceLnz : cfg.nceBias.value();  // MikolovRnn::init + setConfig (:340-346)
    } else {
      // RNN part of the model file: RnnScorerGbeamFactory::load (rnn_scorer_gbeam.cc:419-470)
      auto part = info.firstPartOf(model::ModelPartKind::Rnn);
      if (part == nullptr || part->data.size() != 7) return JPPS_INVALID_STATE << "model has two scorers but no RNN part";
      RnnPartHeader ph;
      util::serialization::Loader l{part->data[0]};
      if (!l.load(&ph)) return JPPS_INVALID_PARAMETER << "failed to read RNN header";
      hdr = ph.rnnHeader;
      JPP_RETURN_IF_ERROR(resolver_.setState(ph.fields, part->data[1], part->data[2], ph.unkIdx));
      rm.matrix = reinterpret_cast<const float*>(part->data[3].begin());
      rm.embeddings = reinterpret_cast<const float*>(part->data[4].begin());
      rm.nce_embeddings = reinterpret_cast<const float*>(part->data[5].begin());
      rm.maxent = reinterpret_cast<const float*>(part->data[6].begin());
      // load() ends with `if (config().rnnWeight.defined()) rnn.setNceConstant(config().rnnWeight)` (:465-467,
      // as written there), and a later setConfig -- JumanppExec::init calls it when the command line
      // carries RNN options -- replaces it by the NCE bias if that is not default (:340-346)
      nce = ph.config.rnnWeight.defined() ? ph.config.rnnWeight.value() : hdr.nceLnz;
      if (opts_.rnnCliConfigApplied && !cfg.nceBias.isDefault()) nce = cfg.nceBias.value();
    }
    JPP_RETURN_IF_ERROR(knownIdx_.loadFromMemory(resolver_.knownIndex()));
    JPP_RETURN_IF_ERROR(unkIdx_.loadFromMemory(resolver_.unkIndex()));
    rm.layer_size = hdr.layerSize;
    rm.maxent_order = hdr.maxentOrder;
    rm.maxent_size = hdr.maxentSize;
    rm.vocab_size = hdr.vocabSize;
    rm.nce_constant = nce;
    rm.unk_id = resolver_.unkId();
    rm.unk_constant = cfg.unkConstantTerm.value();
    rm.unk_length_penalty = cfg.unkLengthPenalty.value();
    m.rnn = &rm;
  }

  jpp_b200_config c{(i32)sconf_.beamSize, acfg_.globalBeamSize, acfg_.rightGbeamCheck, acfg_.rightGbeamSize};
  beam_ = sconf_.beamSize;
  jpp_b200_status st = jpp_b200_create(&m, &c, opts_.device, &handle_);
  if (st != JPP_B200_OK) {
    return JPPS_INVALID_STATE << "jpp_b200_create: " << jpp_b200_last_error(nullptr) << " (status " << (int)st << ")";
  }
  // EntryPtr::rawValue() of a dictionary entry is (offset into the entry storage << 1) | alias bit
  auto dicPart = info.firstPartOf(model::ModelPartKind::Dictionary);
  size_t limit = 0;
  if (dicPart && dicPart->data.size() > 3) limit = dicPart->data[3].size() * 2 + 2;
  resident_.assign(std::max<size_t>(limit, 1024), 0);
  // size the device table once (an entry takes at least one byte per field): rows can then be added
  // while another batch is being scored
  st = jpp_b200_reserve_entry_rows(handle_, (u32)(resident_.size() / (2 * std::max(1u, entrySize_)) + 1024), (u32)resident_.size());
  if (st != JPP_B200_OK) return gpuError(handle_, st, "jpp_b200_reserve_entry_rows");
  return Status::Ok();
}

i32 BatchAnalyzer::resolveRnnId(core::analysis::rnn::RnnReprBuilder* bldr, EntryPtr eptr, util::ArraySlice<i32> entry,
                                const ExtraNodesContext* xtra) const {
  // RnnIdContainer::resolveId (rnn_id_resolver.cc:278-309), which is private to the container
  auto repr = resolver_.reprOf(bldr, eptr, entry, xtra);
  auto t = eptr.isDic() ? knownIdx_.traversal() : unkIdx_.traversal();
  if (t.step(repr) == dic::TraverseStatus::Ok) return t.value();
  return resolver_.unkId();
}

// host pipeline of one sentence, unmodified (Analyzer::analyze, analyzer.cc:45-53, up to the scoring call)
void BatchAnalyzer::buildOne(Slot& sl, size_t i, StringPiece input) {
  auto* impl = sl.impls[i].get();
  SentenceShape& sh = sl.shapes[i];
  sh = SentenceShape{};
  Status s = impl->resetForInput(input);
  if (s) s = impl->prepareNodeSeeds();
  if (s) s = impl->buildLattice();
  if (!s) {
    sl.statuses[i] = std::move(s);
    return;
  }
  sl.statuses[i] = Status::Ok();
  auto lat = impl->lattice();
  // BOS beams: the part of bootstrapAnalysis (analyzer_impl.cc:185-192) that the output side reads
  // (the ScoreProcessor it also allocates is the CPU scorer's scratch: not needed)
  auto beam0 = lat->boundary(0)->starts()->beamData().data();
  EntryBeam::initializeBlock(beam0);
  auto& bosRef = beam0.at(0);
  bosRef = ConnectionBeamElement{{0, 0, 0, 0, nullptr}, 0};
  auto beam1 = lat->boundary(1)->starts()->beamData().data();
  EntryBeam::initializeBlock(beam1);
  beam1.at(0) = ConnectionBeamElement{{1, 0, 0, 0, &bosRef.ptr}, 0};
  const u32 nb = lat->createdBoundaryCount();
  sh.boundaries = nb;
  i32 maxExtra = -1;
  for (u32 b = 0; b < nb; ++b) {
    auto bnd = lat->boundary(b);
    sh.nodes += (u32)bnd->starts()->numEntries();
    sh.ends += (u32)bnd->ends()->nodePtrs().size();
    for (auto& ni : bnd->starts()->nodeInfo()) {
      const EntryPtr e = ni.entryPtr();
      if (e.isSpecial() && e != kBos && e != kEos) maxExtra = std::max(maxExtra, e.extPtr());
    }
  }
  sh.extras = (u32)(maxExtra + 1);
  sh.codepoints = (u32)impl->input().codepoints().size();
}

// the sentence's slices of the staged arrays (resident form of jpp_b200_batch)
void BatchAnalyzer::stageOne(Slot& sl, size_t i, u32 worker) {
  if (!sl.statuses[i]) return;
  auto* impl = sl.impls[i].get();
  auto lat = impl->lattice();
  auto xtra = impl->extraNodesContext();
  const u32 nb = sl.shapes[i].boundaries;
  const u32 bBase = sl.sentBndOff[i], xBase = sl.sentExtraOff[i], cBase = sl.sentCpOff[i];
  u32 n = sl.sentNodeOff[i], e = sl.sentEndOff[i];
  features::impl::PrimitiveFeatureContext pfc{xtra, impl->dic().fields(), impl->dic().entries(), impl->input().codepoints()};
  core::analysis::rnn::RnnReprBuilder reprBldr;
  auto& missing = missingPerWorker_[worker];
  for (u32 b = 0; b < nb; ++b) {
    auto bnd = lat->boundary(b);
    sl.hBndNodeOff[bBase + b] = n;
    sl.hBndEndOff[bBase + b] = e;
    for (auto& ni : bnd->starts()->nodeInfo()) {
      const EntryPtr ep = ni.entryPtr();
      sl.hNodeInfo[n++] = jpp_b200_node_info{ep.rawValue(), ni.start(), ni.end()};
      if (ep.isDic()) {
        // the row of a dictionary entry lives on the device: note the ones it has not seen yet
        const size_t raw = (size_t)ep.rawValue();
        if (raw >= resident_.size() || !resident_[raw]) missing.push_back(ep.rawValue());
      } else if (ep != kBos && ep != kEos) {
        // a node the analyzer made up for this sentence (unknown word, alias): its row travels
        const size_t row = (size_t)xBase + (size_t)ep.extPtr();
        util::MutableArraySlice<i32> buf{sl.hExtraRows.data() + row * entrySize_, entrySize_};
        pfc.fillEntryBuffer(ep, buf);
        for (u32 p = 0; p < numPlaceholders_; ++p) sl.hExtraPh[row * numPlaceholders_ + p] = pfc.providedFeature(ep, p);
        if (hasRnn_) sl.hExtraRnn[row] = (b >= 2 && b + 1 != nb) ? resolveRnnId(&reprBldr, ep, buf, xtra) : 0;
      }
    }
    for (auto& p : bnd->ends()->nodePtrs()) sl.hEnds[e++] = jpp_b200_node_ptr{p.boundary, p.position};
  }
  u32 c = cBase;
  for (auto& cp : impl->input().codepoints()) sl.hCps[c++] = jpp_b200_codepoint{(u32)cp.codepoint, (u32)cp.charClass};
}

// decode the dictionary entries this batch met for the first time and make them resident
Status BatchAnalyzer::addMissingRows() {
  std::vector<i32> all;
  for (auto& v : missingPerWorker_) {
    all.insert(all.end(), v.begin(), v.end());
    v.clear();
  }
  if (all.empty()) return Status::Ok();
  std::sort(all.begin(), all.end());
  all.erase(std::unique(all.begin(), all.end()), all.end());
  if ((size_t)all.back() >= resident_.size()) resident_.resize((size_t)all.back() + 1024, 0);
  std::vector<i32> rows(all.size() * entrySize_), ids(hasRnn_ ? all.size() : 0);
  std::vector<u8> bad(all.size(), 0);
  auto entries = env_->coreHolder()->dic().entries();
  runner_->run(all.size(), [&](size_t k, u32) {
    const EntryPtr ep{all[k]};
    util::MutableArraySlice<i32> buf{rows.data() + k * entrySize_, entrySize_};
    if (entries.entryAtPtr(ep).fill(buf, entrySize_) != entrySize_) bad[k] = 1;  // fillEntryBuffer, feature_impl_types.h:141-142
    if (hasRnn_) {
      core::analysis::rnn::RnnReprBuilder bldr;
      ids[k] = resolveRnnId(&bldr, ep, buf, nullptr);
    }
  }, 256);
  for (auto b : bad)
    if (b) return JPPS_INVALID_STATE << "failed to decode a dictionary entry";
  jpp_b200_entry_table t{(u32)all.size(), all.data(), rows.data(), hasRnn_ ? ids.data() : nullptr};
  jpp_b200_status st = jpp_b200_add_entry_rows(handle_, &t);
  if (st != JPP_B200_OK) return gpuError(handle_, st, "jpp_b200_add_entry_rows");
  for (auto p : all) resident_[(size_t)p] = 1;
  timings_.tableRowsAdded += all.size();
  return Status::Ok();
}

// GPU result -> the sentence's Lattice, in the form the output side reads (analysis_result.cc:25-76,
// lattice_format.cc): beam rows with their `previous` links, per-connection scores
void BatchAnalyzer::unstageOne(Slot& sl, size_t i) {
  if (!sl.statuses[i]) return;
  if (sl.hStatus[i] != 0) {
    sl.statuses[i] = JPPS_INVALID_STATE << "the lattice failed validation on the device (status " << (int)sl.hStatus[i] << ")";
    return;
  }
  auto lat = sl.impls[i]->lattice();
  const u32 nb = sl.shapes[i].boundaries;
  const u32 bBase = sl.sentBndOff[i];
  const u32 B = beam_, S = numScorers_;
  auto nodeOf = [&](u32 bnd, u32 pos) { return sl.hBndNodeOff[bBase + bnd] + pos; };
  if (opts_.fullBeams) {
    for (u32 b = 2; b < nb; ++b) {
      auto bnd = lat->boundary(b);
      auto starts = bnd->starts();
      auto ends = bnd->ends()->nodePtrs();
      const u32 R = (u32)starts->numEntries();
      for (u32 r = 0; r < R; ++r) {
        const u32 node = nodeOf(b, r);
        auto row = starts->beamData().row(r);
        for (u32 k = 0; k < B; ++k) {
          const jpp_b200_beam_element& el = sl.hBeam[(size_t)node * B + k];
          if (el.left == 0xFFFFu && el.prev_beam == 0xFFFFu) {
            row.at(k) = EntryBeam::fake();
            continue;
          }
          const LatticeNodePtr lp = ends.at(el.left);
          const ConnectionBeamElement& prev = lat->boundary(lp.boundary)->starts()->beamData().row(lp.position).at(el.prev_beam);
          ConnectionBeamElement& out = row.at(k);
          out.ptr = ConnectionPtr{(u16)b, el.left, (u16)r, el.prev_beam, &prev.ptr};
          out.totalScore = el.total_score;
          auto sc = bnd->scores()->nodeScores(r).beamLeft(el.prev_beam, el.left);
          for (u32 s = 0; s < S; ++s) sc.at(s) = sl.hBeamScores[((size_t)node * B + k) * S + s];
        }
      }
    }
    return;
  }
  // 1-best chain only: the element of every node on the path goes to slot 0 of its row, so that the
  // walk of AnalysisPath::fillIn (beam index of the previous node = `beam`) sees exactly this chain
  const u32 len = sl.hPathLen[i];
  const ConnectionPtr* prevPtr = &lat->boundary(1)->starts()->beamData().row(0).at(0).ptr;
  u32 prevBnd = 1, prevPos = 0;
  auto link = [&](u32 b, u32 r) {
    auto bnd = lat->boundary(b);
    auto ends = bnd->ends()->nodePtrs();
    u16 left = 0;
    for (u32 l = 0; l < ends.size(); ++l)
      if (ends.at(l).boundary == prevBnd && ends.at(l).position == prevPos) { left = (u16)l; break; }
    ConnectionBeamElement& out = bnd->starts()->beamData().row(r).at(0);
    out.ptr = ConnectionPtr{(u16)b, left, (u16)r, 0, prevPtr};
    out.totalScore = 0;
    prevPtr = &out.ptr;
    prevBnd = b;
    prevPos = r;
  };
  u32 bCursor = 2;
  for (u32 k = 0; k < len; ++k) {
    const u32 node = sl.hPathNodes[bBase + k];
    while (bCursor + 1 < nb && sl.hBndNodeOff[bBase + bCursor + 1] <= node) ++bCursor;  // path nodes come in boundary order
    link(bCursor, node - sl.hBndNodeOff[bBase + bCursor]);
  }
  if (len > 0 || nb == 3) {
    if (nb > 3) link(nb - 1, 0);
  } else {
    sl.statuses[i] = JPPS_INVALID_STATE << "no path reaches the end of the sentence";
  }
}

Status BatchAnalyzer::prepare(int slot, const std::vector<StringPiece>& inputs) {
  Slot& sl = slots_[slot];
  const size_t N = inputs.size();
  if (N > opts_.maxBatch) return JPPS_INVALID_PARAMETER << "batch of " << N << " sentences, at most " << opts_.maxBatch;
  sl.batchSize = N;
  if (N == 0) return Status::Ok();
  double t0 = now();
  runner_->run(N, [&](size_t i, u32) { buildOne(sl, i, inputs[i]); }, 16);
  double t1 = now();
  // CSR offsets: a failed sentence takes no room (the device sees it as a sentence without boundaries
  // and reports it, which un-staging maps back to the status the host already has)
  sl.sentBndOff.assign(N + 1, 0); sl.sentCpOff.assign(N + 1, 0); sl.sentExtraOff.assign(N + 1, 0);
  sl.sentNodeOff.assign(N + 1, 0); sl.sentEndOff.assign(N + 1, 0);
  for (size_t i = 0; i < N; ++i) {
    const SentenceShape z = sl.statuses[i] ? sl.shapes[i] : SentenceShape{};
    sl.sentBndOff[i + 1] = sl.sentBndOff[i] + z.boundaries;
    sl.sentCpOff[i + 1] = sl.sentCpOff[i] + z.codepoints;
    sl.sentExtraOff[i + 1] = sl.sentExtraOff[i] + z.extras;
    sl.sentNodeOff[i + 1] = sl.sentNodeOff[i] + z.nodes;
    sl.sentEndOff[i + 1] = sl.sentEndOff[i] + z.ends;
  }
  const size_t NB = sl.sentBndOff[N], NN = sl.sentNodeOff[N], NE = sl.sentEndOff[N], NC = sl.sentCpOff[N], NX = sl.sentExtraOff[N];
  sl.NB = NB; sl.NN = NN; sl.NE = NE; sl.NC = NC; sl.NX = NX;
  JPP_RETURN_IF_ERROR(sl.hSentBndOff.ensure(N + 1)); JPP_RETURN_IF_ERROR(sl.hSentCpOff.ensure(N + 1));
  JPP_RETURN_IF_ERROR(sl.hSentExtraOff.ensure(N + 1));
  JPP_RETURN_IF_ERROR(sl.hBndNodeOff.ensure(NB + 1)); JPP_RETURN_IF_ERROR(sl.hBndEndOff.ensure(NB + 1));
  JPP_RETURN_IF_ERROR(sl.hNodeInfo.ensure(NN)); JPP_RETURN_IF_ERROR(sl.hEnds.ensure(NE)); JPP_RETURN_IF_ERROR(sl.hCps.ensure(NC));
  JPP_RETURN_IF_ERROR(sl.hExtraRows.ensure(std::max<size_t>(1, NX) * entrySize_));
  JPP_RETURN_IF_ERROR(sl.hExtraPh.ensure(std::max<size_t>(1, NX) * std::max(1u, numPlaceholders_)));
  JPP_RETURN_IF_ERROR(sl.hExtraRnn.ensure(std::max<size_t>(1, NX)));
  JPP_RETURN_IF_ERROR(sl.hPathLen.ensure(N)); JPP_RETURN_IF_ERROR(sl.hPathNodes.ensure(NB)); JPP_RETURN_IF_ERROR(sl.hStatus.ensure(N));
  if (opts_.fullBeams) {
    JPP_RETURN_IF_ERROR(sl.hBeam.ensure(NN * beam_));
    JPP_RETURN_IF_ERROR(sl.hBeamScores.ensure(NN * beam_ * numScorers_));
  }
  std::memcpy(sl.hSentBndOff.data(), sl.sentBndOff.data(), (N + 1) * 4);
  std::memcpy(sl.hSentCpOff.data(), sl.sentCpOff.data(), (N + 1) * 4);
  std::memcpy(sl.hSentExtraOff.data(), sl.sentExtraOff.data(), (N + 1) * 4);
  sl.hBndNodeOff[NB] = (u32)NN;
  sl.hBndEndOff[NB] = (u32)NE;
  // rows of extra indices that no lattice node uses stay defined
  std::memset(sl.hExtraRows.data(), 0, std::max<size_t>(1, NX) * entrySize_ * 4);
  std::memset(sl.hExtraPh.data(), 0, std::max<size_t>(1, NX) * std::max(1u, numPlaceholders_) * 4);
  std::memset(sl.hExtraRnn.data(), 0, std::max<size_t>(1, NX) * 4);
  runner_->run(N, [&](size_t i, u32 w) { stageOne(sl, i, w); }, 16);
  double t2 = now();
  JPP_RETURN_IF_ERROR(addMissingRows());
  double t3 = now();
  std::lock_guard<std::mutex> lk(timingsMutex_);
  timings_.buildSec += t1 - t0; timings_.stageSec += t2 - t1; timings_.tableSec += t3 - t2;
  timings_.sentences += N; timings_.nodes += NN;
  timings_.h2dBytes += (N + 1) * 12 + (NB + 1) * 8 + NN * 8 + NE * 4 + NC * 8 + NX * (entrySize_ + numPlaceholders_ + 1) * 4;
  timings_.d2hBytes += N * 5 + NB * 4 + (opts_.fullBeams ? NN * beam_ * (8 + 4 * numScorers_) : 0);
  return Status::Ok();
}

Status BatchAnalyzer::score(int slot) {
  Slot& sl = slots_[slot];
  const size_t N = sl.batchSize;
  if (N == 0) return Status::Ok();
  double t0 = now();
  jpp_b200_batch b{};
  b.num_sentences = (u32)N;
  b.sent_boundary_off = sl.hSentBndOff.data(); b.sent_codepoint_off = sl.hSentCpOff.data();
  b.boundary_node_off = sl.hBndNodeOff.data(); b.boundary_end_off = sl.hBndEndOff.data();
  b.node_info = sl.hNodeInfo.data(); b.ends = sl.hEnds.data(); b.codepoints = sl.hCps.data();
  b.entry_rows = sl.hExtraRows.data(); b.num_entry_rows = (u32)sl.NX;
  b.row_placeholders = numPlaceholders_ ? sl.hExtraPh.data() : nullptr;
  b.row_rnn_id = hasRnn_ ? sl.hExtraRnn.data() : nullptr;
  b.sent_extra_row_off = sl.hSentExtraOff.data();
  jpp_b200_result r{};
  r.best_path_len = sl.hPathLen.data();
  r.best_path_nodes = sl.hPathNodes.data();
  r.sentence_status = sl.hStatus.data();
  if (opts_.fullBeams) {
    r.node_beam = sl.hBeam.data();
    r.node_beam_scores = sl.hBeamScores.data();
  }
  if (sl.NB > 0) {
    jpp_b200_status st = jpp_b200_score_batch(handle_, &b, &r);  // <- the scoring
    if (st != JPP_B200_OK) return gpuError(handle_, st, "jpp_b200_score_batch");
  } else {
    std::memset(sl.hStatus.data(), 0, N);
  }
  std::lock_guard<std::mutex> lk(timingsMutex_);
  timings_.scoreSec += now() - t0;
  return Status::Ok();
}

Status BatchAnalyzer::finish(int slot) {
  Slot& sl = slots_[slot];
  const size_t N = sl.batchSize;
  if (N == 0) return Status::Ok();
  double t0 = now();
  runner_->run(N, [&](size_t i, u32) { unstageOne(sl, i); }, 16);
  std::lock_guard<std::mutex> lk(timingsMutex_);
  timings_.unstageSec += now() - t0;
  for (size_t i = 0; i < N; ++i) timings_.failed += !sl.statuses[i];
  return Status::Ok();
}

}  // namespace b200
}  // namespace jumanpp
