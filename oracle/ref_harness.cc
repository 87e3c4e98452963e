// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the shipped product.
//
// Harness around the UNMODIFIED ku-nlp/jumanpp reference (compiled in place by
// oracle/build_ref.py into oracle/_ref/libjpp_ref.a).  It does the three
// things the parity / measurement plan needs from the real reference:
//
//   stage : run the reference's own host pipeline
//           (AnalyzerImpl::resetForInput -> prepareNodeSeeds -> buildLattice,
//           reference src/core/analysis/analyzer_impl.cc:19-177) on a corpus,
//           pack every lattice into the jpp_b200_batch layout of
//           include/jpp_b200.h, and -- with expected=1 -- also run the
//           reference's bootstrapAnalysis + computeScores (:179-309) and dump
//           the post-state (beams, per-connection scores, stored patterns).
//           These files are the golden vectors that pin oracle/jpp_oracle.c
//           and the CUDA path.
//   bench : time the reference's computeScores alone (the hot path) and the
//           whole Analyzer::analyze, with N worker threads each owning an
//           AnalyzerImpl over one shared CoreHolder (the pattern of reference
//           src/jumandic/main/path_diff.cc:174-205).  This is bench.py's
//           `--impl reference` arm and its cpu_baseline.
//   spec  : dump the compiled Jumandic feature program only.
//
// Arguments are key=value; see usage().  Output container: "JPPB0001" -- a
// flat list of named n-d arrays (reader: jumanpp_b200/jppb.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "core/analysis/analyzer_impl.h"
#include "core/analysis/perceptron.h"
#include "core/analysis/rnn_id_resolver.h"
#include "core/analysis/rnn_scorer_gbeam.h"
#include "core/analysis/score_processor.h"
#include "core/core.h"
#include "core/dic/darts_trie.h"
#include "core/dic/dic_builder.h"
#include "core/dic/dictionary.h"
#include "core/impl/feature_impl_types.h"
#include "core/impl/model_io.h"
#include "core/impl/perceptron_io.h"
#include "util/serialization.h"
#include <unistd.h>
#include "jpp_jumandic_cg.h"
#include "jumandic/shared/juman_format.h"
#include "jumandic/shared/jumandic_spec.h"
#include "core/analysis/analyzer.h"
#include "rnn/mikolov_rnn.h"
#include "util/mmap.h"

#include "jpp_b200.h"

using namespace jumanpp;
using namespace jumanpp::core;
using namespace jumanpp::core::analysis;

namespace {

// ------------------------------------------------------------------ args ----
struct Args {
  std::map<std::string, std::string> kv;
  std::string str(const std::string& k, const std::string& d = "") const {
    auto it = kv.find(k);
    return it == kv.end() ? d : it->second;
  }
  long num(const std::string& k, long d) const {
    auto it = kv.find(k);
    return it == kv.end() ? d : std::strtol(it->second.c_str(), nullptr, 10);
  }
  double real(const std::string& k, double d) const {
    auto it = kv.find(k);
    return it == kv.end() ? d : std::strtod(it->second.c_str(), nullptr);
  }
  bool has(const std::string& k) const { return kv.count(k) != 0; }
};

[[noreturn]] void die(const std::string& msg) {
  std::cerr << "jpp_ref_harness: " << msg << "\n";
  std::exit(2);
}

void check(const Status& s, const char* what) {
  if (getenv("JPP_HARNESS_TRACE")) std::cerr << "[trace] " << what << "\n";
  if (!s) {
    std::cerr << "jpp_ref_harness: " << what << ": " << s << "\n";
    std::exit(2);
  }
}

// --------------------------------------------------- counter-based weights ---
// twin of jumanpp_b200/synth.py:make_weights
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

inline float synth_weight(uint64_t seed, uint64_t i, float scale) {
  uint64_t h = splitmix64(i + (seed << 40));
  uint32_t s = (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) +
               (uint32_t)((h >> 32) & 0xFFFF) + (uint32_t)((h >> 48) & 0xFFFF);
  float v = (float)s * (1.0f / 65536.0f) - 2.0f;
  float f = (float)((double)scale * 1.7320508075688772);
  return v * f;
}

// ------------------------------------------------------------ JPPB writer ---
enum DType : uint32_t { U8 = 0, I32 = 1, U32 = 2, F32 = 3, U64 = 4, U16 = 5 };

struct JppbWriter {
  struct Arr {
    std::string name;
    uint32_t dtype;
    std::vector<uint64_t> dims;
    const void* data;
    uint64_t nbytes;
  };
  std::vector<Arr> arrs;
  template <typename T>
  void add(const std::string& name, DType dt, const std::vector<T>& v,
           std::vector<uint64_t> dims = {}) {
    if (dims.empty()) dims = {v.size()};
    arrs.push_back({name, dt, dims, v.data(), v.size() * sizeof(T)});
  }
  void write(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) die("cannot open " + path);
    std::fwrite("JPPB0001", 1, 8, f);
    uint32_t n = (uint32_t)arrs.size();
    std::fwrite(&n, 4, 1, f);
    for (auto& a : arrs) {
      uint32_t nl = (uint32_t)a.name.size();
      std::fwrite(&nl, 4, 1, f);
      std::fwrite(a.name.data(), 1, nl, f);
      std::fwrite(&a.dtype, 4, 1, f);
      uint32_t nd = (uint32_t)a.dims.size();
      std::fwrite(&nd, 4, 1, f);
      std::fwrite(a.dims.data(), 8, nd, f);
      std::fwrite(&a.nbytes, 8, 1, f);
      // pad so that data starts 8-aligned
      long pos = std::ftell(f);
      static const char zeros[8] = {0};
      std::fwrite(zeros, 1, (8 - pos % 8) % 8, f);
      if (a.nbytes) std::fwrite(a.data, 1, a.nbytes, f);
    }
    std::fclose(f);
  }
};

// ------------------------------------------------------------------- env ----
struct RefEnv {
  spec::AnalysisSpec spec;
  dic::DictionaryBuilder builder;
  model::FilesystemModel fsModel;
  model::ModelInfo modelInfo;
  model::ModelInfo dicOnlyInfo;
  dic::BuiltDictionary restored;
  dic::DictionaryHolder dic;
  std::unique_ptr<CoreHolder> core;
  jumanpp_generated::JumandicStatic staticFeatures;
  util::FullyMappedFile dicFile;
  float* weights = nullptr;
  size_t numWeights = 0;
  bool explicitWeights = false;
  std::unique_ptr<HashedFeaturePerceptron> perceptron;
  RnnScorerGbeamFactory rnnFactory;
  bool hasRnn = false;
  ScorerDef sdef;
  AnalyzerConfig acfg;
  ScoringConfig sconf{5, 1};
  // rnn id resolution (RnnIdContainer::resolveId is private: rebuild the two
  // tries from the resolver's public serialised form)
  dic::DoubleArray knownIdx, unkIdx;
  jumanpp::core::analysis::rnn::RnnIdResolver resolver;
  jumanpp::core::analysis::rnn::RnnInferenceConfig rnnCfg;
  ::jumanpp::rnn::mikolov::MikolovModelReader rnnReader;

  void init(const Args& a) {
    check(jumandic::SpecFactory::makeSpec(&spec), "spec");
    check(dicFile.open(a.str("dict"), util::MMapType::ReadOnly), "open dict");
    check(builder.importSpec(&spec), "importSpec");
    check(builder.importCsv(a.str("dict"), dicFile.contents()), "importCsv");
    // same round trip as the reference's own fixtures (src/testing/test_analyzer.h:59-91)
    // and jpp_jumandic_bootstrap (src/jumandic/main/bootstrap.cc:57-68): save the
    // built dictionary as a model file, map it back, restore.
    std::string tmpModel = a.str("tmp_model", "/tmp/jpp_ref_harness_" + std::to_string((long)getpid()) + ".model");
    {
      dic::DictionaryHolder probe;
      check(probe.load(builder.result()), "dic probe load");
      model::ModelInfo& nfo = dicOnlyInfo;  // fillModelPart may run once per builder: cmdModel reuses this part
      nfo.parts.emplace_back();
      check(builder.fillModelPart(&nfo.parts.back()), "fillModelPart");
      model::ModelSaver saver;
      check(saver.open(tmpModel), "saver open");
      check(saver.save(nfo), "saver save");
    }
    check(fsModel.open(tmpModel), "model open");
    check(fsModel.load(&modelInfo), "model load");
    check(restored.restoreDictionary(modelInfo), "restoreDictionary");
    check(dic.load(restored), "dic load");
    std::remove(tmpModel.c_str());
    core.reset(new CoreHolder{restored.spec, dic});
    bool dynamic = a.num("dynamic", 0) != 0;
    check(core->initialize(dynamic ? nullptr : &staticFeatures), "core init");

    int log2w = (int)a.num("log2w", 20);
    uint64_t wseed = (uint64_t)a.num("wseed", 1);
    float wscale = (float)a.real("wscale", 0.5);
    numWeights = size_t{1} << log2w;
    weights = static_cast<float*>(aligned_alloc(64, std::max<size_t>(64, numWeights * sizeof(float))));
    if (a.has("weights_file")) {
      // explicit table (raw little-endian float32, exactly 2^log2w values)
      std::ifstream wf(a.str("weights_file"), std::ios::binary);
      if (!wf) die("cannot open weights_file");
      wf.read(reinterpret_cast<char*>(weights), numWeights * sizeof(float));
      if ((size_t)wf.gcount() != numWeights * sizeof(float)) die("weights_file too short");
      explicitWeights = true;
    } else {
      unsigned nt = std::max(1u, std::thread::hardware_concurrency());
      std::vector<std::thread> ts;
      for (unsigned t = 0; t < nt; ++t) {
        ts.emplace_back([=] {
          for (size_t i = t; i < numWeights; i += nt) weights[i] = synth_weight(wseed, i, wscale);
        });
      }
      for (auto& t : ts) t.join();
    }
    perceptron.reset(new HashedFeaturePerceptron{util::ArraySlice<float>{weights, numWeights}});
    sdef.feature = perceptron.get();
    sdef.scoreWeights.push_back((float)a.real("perc_weight", 1.0));

    sconf.beamSize = (i32)a.num("beam", 5);
    acfg.globalBeamSize = (i32)a.num("gbeam", 6);
    acfg.rightGbeamCheck = (i32)a.num("rcheck", 1);
    acfg.rightGbeamSize = (i32)a.num("rbeam", 5);
    acfg.maxInputBytes = 64 * 1024;

    if (a.has("rnn")) {
      hasRnn = true;
      rnnCfg.nceBias = (float)a.real("nce_bias", 5.62844432562);
      rnnCfg.unkConstantTerm = (float)a.real("unk_const", -3.4748115191);
      rnnCfg.unkLengthPenalty = (float)a.real("unk_len", -2.92994951022);
      rnnCfg.perceptronWeight = (float)a.real("perc_weight", 1.0);
      rnnCfg.rnnWeight = (float)a.real("rnn_weight", 0.0176);
      rnnCfg.rnnFields = std::vector<std::string>{"baseform", "pos"};
      check(rnnFactory.make(a.str("rnn"), dic, rnnCfg), "rnn make");
      sdef.others.push_back(&rnnFactory);
      sdef.scoreWeights.push_back(rnnFactory.config().rnnWeight);
      sdef.scoreWeights[0] = rnnFactory.config().perceptronWeight;
      sconf.numScorers = 2;
      // second resolver instance built the same way, to reach the tries
      auto& rd = rnnReader;
      check(rd.open(a.str("rnn")), "rnn open");
      check(rd.parse(), "rnn parse");
      check(resolver.build(dic, rnnFactory.config(), rd.words()), "resolver");
      check(knownIdx.loadFromMemory(resolver.knownIndex()), "known idx");
      check(unkIdx.loadFromMemory(resolver.unkIndex()), "unk idx");
    }
  }

  std::unique_ptr<AnalyzerImpl> makeAnalyzer() const {
    std::unique_ptr<AnalyzerImpl> p{new AnalyzerImpl{core.get(), sconf, acfg}};
    check(p->initScorers(sdef), "initScorers");
    return p;
  }

  i32 resolveRnnId(jumanpp::core::analysis::rnn::RnnReprBuilder* bldr, EntryPtr eptr, util::ArraySlice<i32> entry,
                   const ExtraNodesContext* xtra) const {
    // rnn_id_resolver.cc:278-309
    auto repr = resolver.reprOf(bldr, eptr, entry, xtra);
    if (eptr.isDic()) {
      auto t = knownIdx.traversal();
      if (t.step(repr) == dic::TraverseStatus::Ok) return t.value();
      return resolver.unkId();
    }
    auto t = unkIdx.traversal();
    if (t.step(repr) == dic::TraverseStatus::Ok) return t.value();
    return resolver.unkId();
  }
};

// --------------------------------------------------------- spec export ------
struct SpecArrays {
  std::vector<int32_t> prims, comps, branchPool, patterns, patRefPool, ngrams;
  std::vector<int32_t> counts;  // n_prim, n_comp, n_pat, n_stored, n_ngram, entry_size, n_placeholders

  void fill(const RefEnv& env) {
    auto& fs = env.restored.spec.features;
    for (auto& p : fs.primitive) {
      int32_t kind = 0, a = 0, b = 0, c = 0;
      using K = spec::PrimitiveFeatureKind;
      switch (p.kind) {
        case K::Copy: kind = JPP_B200_PRIM_COPY; a = p.references[0]; break;
        case K::SingleBit: kind = JPP_B200_PRIM_SHIFT_MASK; a = p.references[0]; b = p.references[1]; c = 1; break;
        case K::Provided: kind = JPP_B200_PRIM_PROVIDED; a = p.references[0]; break;
        case K::ByteLength: kind = JPP_B200_PRIM_BYTE_LENGTH; a = p.references[0]; break;
        case K::CodepointSize: kind = JPP_B200_PRIM_CODEPOINT_SIZE; a = p.references[0]; break;
        case K::SurfaceCodepointSize: kind = JPP_B200_PRIM_SURFACE_CODEPOINT_SIZE; break;
        case K::CodepointType: kind = JPP_B200_PRIM_CODEPOINT_TYPE; a = p.references[0]; break;
        case K::Codepoint: kind = JPP_B200_PRIM_CODEPOINT; a = p.references[0]; break;
        default: die("invalid primitive kind");
      }
      prims.insert(prims.end(), {kind, a, b, c});
    }
    for (auto& c : fs.computation) {
      int32_t toff = (int32_t)branchPool.size();
      for (auto x : c.trueBranch) branchPool.push_back(x);
      int32_t foff = (int32_t)branchPool.size();
      for (auto x : c.falseBranch) branchPool.push_back(x);
      comps.insert(comps.end(), {c.primitiveFeature, toff, (int32_t)c.trueBranch.size(), foff,
                                 (int32_t)c.falseBranch.size()});
    }
    for (auto& p : fs.pattern) {
      int32_t off = (int32_t)patRefPool.size();
      for (auto x : p.references) patRefPool.push_back(x);
      patterns.insert(patterns.end(), {p.index, off, (int32_t)p.references.size()});
    }
    for (auto& n : fs.ngram) {
      int32_t r[3] = {0, 0, 0};
      for (size_t i = 0; i < n.references.size(); ++i) r[i] = n.references[i];
      ngrams.insert(ngrams.end(), {(int32_t)n.references.size(), n.index, r[0], r[1], r[2]});
    }
    counts = {(int32_t)fs.primitive.size(), (int32_t)fs.computation.size(), (int32_t)fs.pattern.size(),
              (int32_t)(fs.pattern.size() - fs.numUniOnlyPats), (int32_t)fs.ngram.size(),
              (int32_t)env.dic.entries().numFeatures(), (int32_t)fs.numPlaceholders};
  }
  void addTo(JppbWriter& w) {
    w.add("spec_counts", I32, counts);
    w.add("spec_prims", I32, prims, {prims.size() / 4, 4});
    w.add("spec_comps", I32, comps, {comps.size() / 5, 5});
    w.add("spec_branch_pool", I32, branchPool);
    w.add("spec_patterns", I32, patterns, {patterns.size() / 3, 3});
    w.add("spec_pattern_ref_pool", I32, patRefPool);
    w.add("spec_ngrams", I32, ngrams, {ngrams.size() / 5, 5});
  }
};

// --------------------------------------------------------------- staging ----
struct Staged {
  std::vector<uint32_t> sentBndOff{0}, sentCpOff{0}, bndNodeOff{0}, bndEndOff{0};
  std::vector<int32_t> nodeEntryPtr;
  std::vector<uint16_t> nodeStartEnd;  // [n][2]
  std::vector<int32_t> nodeEntry, nodePlaceholders, nodeRnnId;
  std::vector<uint16_t> ends;           // [n][2] boundary, position
  std::vector<uint32_t> codepoints;     // [n][2]
  // expected
  std::vector<uint16_t> beamLP;   // [nodes][beam][2]  left, prev_beam
  std::vector<float> beamTotal;   // [nodes][beam]
  std::vector<float> beamScores;  // [nodes][beam][S]
  std::vector<uint64_t> patterns; // [nodes][stored]
  std::vector<uint8_t> sentOk;
  std::vector<uint32_t> sentIndex;  // index of the sentence in the corpus
};

void stageSentence(const RefEnv& env, AnalyzerImpl* ana, bool expected, Staged* st,
                   jumanpp::core::analysis::rnn::RnnReprBuilder* reprBldr) {
  auto lat = ana->lattice();
  auto xtra = ana->extraNodesContext();
  auto entries = env.dic.entries();
  features::impl::PrimitiveFeatureContext pfc{xtra, env.dic.fields(), entries, ana->input().codepoints()};
  u32 nb = lat->createdBoundaryCount();
  u32 esz = (u32)entries.numFeatures();
  u32 nph = (u32)env.restored.spec.features.numPlaceholders;
  u32 beam = lat->config().beamSize;
  u32 S = lat->config().scoreCnt;
  u32 stored = lat->config().numFeaturePatterns;
  std::vector<i32> ebuf(esz);
  for (u32 b = 0; b < nb; ++b) {
    auto bnd = lat->boundary(b);
    auto starts = bnd->starts();
    auto infos = starts->nodeInfo();
    for (u32 n = 0; n < infos.size(); ++n) {
      auto& ni = infos.at(n);
      st->nodeEntryPtr.push_back(ni.entryPtr().rawValue());
      st->nodeStartEnd.push_back(ni.start());
      st->nodeStartEnd.push_back(ni.end());
      util::MutableArraySlice<i32> es{ebuf.data(), esz};
      if (b < 2) {
        // BOS rows: the reference never fills entryData for them; the rnn /
        // pattern code only needs the constant pattern fill (lattice_builder.cc:171-176)
        util::fill(es, ni.entryPtr().rawValue());
      } else if (!pfc.fillEntryBuffer(ni.entryPtr(), es)) {
        die("fillEntryBuffer failed");
      }
      st->nodeEntry.insert(st->nodeEntry.end(), ebuf.begin(), ebuf.end());
      for (u32 p = 0; p < nph; ++p) {
        st->nodePlaceholders.push_back(pfc.providedFeature(ni.entryPtr(), p));
      }
      if (env.hasRnn) {
        i32 id = 0;
        if (b >= 2 && b != nb - 1) {
          id = env.resolveRnnId(reprBldr, ni.entryPtr(), util::ArraySlice<i32>{ebuf.data(), esz}, xtra);
        }
        st->nodeRnnId.push_back(id);
      }
      if (expected) {
        auto row = starts->beamData().row(n);
        for (u32 k = 0; k < beam; ++k) {
          auto& el = row.at(k);
          bool fake = EntryBeam::isFake(el);
          st->beamLP.push_back(fake ? 0xFFFF : el.ptr.left);
          st->beamLP.push_back(fake ? 0xFFFF : el.ptr.beam);
          st->beamTotal.push_back(el.totalScore);
          for (u32 s = 0; s < S; ++s) {
            float v = 0;
            if (!fake && b >= 2) v = bnd->scores()->forPtr(el.ptr).at(s);
            st->beamScores.push_back(v);
          }
        }
        auto prow = starts->patternFeatureData().row(n);
        for (u32 p = 0; p < stored; ++p) st->patterns.push_back(prow.at(p));
      }
    }
    st->bndNodeOff.push_back((uint32_t)st->nodeEntryPtr.size());
    for (auto& e : bnd->ends()->nodePtrs()) {
      st->ends.push_back(e.boundary);
      st->ends.push_back(e.position);
    }
    st->bndEndOff.push_back((uint32_t)(st->ends.size() / 2));
  }
  st->sentBndOff.push_back((uint32_t)(st->bndNodeOff.size() - 1));
  for (auto& cp : ana->input().codepoints()) {
    st->codepoints.push_back((uint32_t)cp.codepoint);
    st->codepoints.push_back((uint32_t)cp.charClass);
  }
  st->sentCpOff.push_back((uint32_t)(st->codepoints.size() / 2));
}

std::vector<std::string> readLines(const std::string& path, long maxSent) {
  std::ifstream in(path);
  if (!in) die("cannot open corpus " + path);
  std::vector<std::string> lines;
  std::string l;
  while (std::getline(in, l)) {
    if (l.empty()) continue;
    lines.push_back(l);
    if (maxSent > 0 && (long)lines.size() >= maxSent) break;
  }
  return lines;
}

int cmdStage(const Args& a) {
  RefEnv env;
  env.init(a);
  SpecArrays sa;
  sa.fill(env);
  bool expected = a.num("expected", 1) != 0;
  auto lines = readLines(a.str("corpus"), a.num("max_sent", 0));
  long chunk = a.num("chunk", (long)lines.size());
  if (chunk <= 0) chunk = (long)lines.size();
  long nchunks = ((long)lines.size() + chunk - 1) / chunk;
  unsigned nt = (unsigned)a.num("threads", 1);
  std::atomic<long> next{0};
  std::string out = a.str("out");
  std::vector<std::thread> ts;
  std::atomic<long> failed{0};
  for (unsigned t = 0; t < std::max(1u, nt); ++t) {
    ts.emplace_back([&] {
      auto ana = env.makeAnalyzer();
      jumanpp::core::analysis::rnn::RnnReprBuilder reprBldr;
      while (true) {
        long c = next.fetch_add(1);
        if (c >= nchunks) break;
        Staged st;
        long lo = c * chunk, hi = std::min<long>(lo + chunk, (long)lines.size());
        for (long i = lo; i < hi; ++i) {
          Status s = ana->resetForInput(lines[i]);
          if (s) s = ana->prepareNodeSeeds();
          if (s) s = ana->buildLattice();
          if (s && expected) {
            s = ana->bootstrapAnalysis();
            if (s) s = ana->computeScores(&env.sdef);
          }
          if (!s) {
            failed.fetch_add(1);
            continue;
          }
          st.sentIndex.push_back((uint32_t)i);
          stageSentence(env, ana.get(), expected, &st, &reprBldr);
        }
        JppbWriter w;
        std::vector<int32_t> cfg{env.sconf.beamSize, env.acfg.globalBeamSize, env.acfg.rightGbeamCheck,
                                 env.acfg.rightGbeamSize, env.sconf.numScorers, (int32_t)a.num("log2w", 20),
                                 (int32_t)a.num("wseed", 1), expected ? 1 : 0, a.num("dynamic", 0) != 0 ? 1 : 0};
        std::vector<float> fcfg{(float)a.real("wscale", 0.5), env.sdef.scoreWeights[0],
                                env.sdef.scoreWeights.size() > 1 ? env.sdef.scoreWeights[1] : 0.0f};
        w.add("cfg", I32, cfg);
        w.add("fcfg", F32, fcfg);
        std::vector<float> wcopy;
        if (env.explicitWeights) {
          wcopy.assign(env.weights, env.weights + env.numWeights);
          w.add("weights", F32, wcopy);
        }
        sa.addTo(w);
        std::vector<float> rnnf;
        std::vector<int32_t> rnni;
        if (env.hasRnn) {
          auto& h = env.rnnReader.header();
          rnni = std::vector<int32_t>{(int32_t)h.layerSize, (int32_t)h.maxentOrder, (int32_t)h.maxentSize,
                                      (int32_t)h.vocabSize, env.resolver.unkId()};
          // MikolovRnn::nceConstant() after setConfig (rnn_scorer_gbeam.cc:352-358)
          float nceC = env.rnnFactory.config().nceBias.isDefault() ? h.nceLnz : env.rnnFactory.config().nceBias.value();
          float unkC = env.rnnFactory.config().unkConstantTerm.value();
          float unkL = env.rnnFactory.config().unkLengthPenalty.value();
          rnnf = std::vector<float>{nceC, unkC, unkL};
          w.add("rnn_icfg", I32, rnni);
          w.add("rnn_fcfg", F32, rnnf);
        }
        w.add("sent_index", U32, st.sentIndex);
        w.add("sent_boundary_off", U32, st.sentBndOff);
        w.add("sent_codepoint_off", U32, st.sentCpOff);
        w.add("boundary_node_off", U32, st.bndNodeOff);
        w.add("boundary_end_off", U32, st.bndEndOff);
        w.add("node_entry_ptr", I32, st.nodeEntryPtr);
        w.add("node_start_end", U16, st.nodeStartEnd, {st.nodeStartEnd.size() / 2, 2});
        uint64_t esz = (uint64_t)sa.counts[5], nph = (uint64_t)sa.counts[6];
        w.add("node_entry", I32, st.nodeEntry, {st.nodeEntry.size() / esz, esz});
        if (nph) w.add("node_placeholders", I32, st.nodePlaceholders, {st.nodePlaceholders.size() / nph, nph});
        if (env.hasRnn) w.add("node_rnn_id", I32, st.nodeRnnId);
        w.add("ends", U16, st.ends, {st.ends.size() / 2, 2});
        w.add("codepoints", U32, st.codepoints, {st.codepoints.size() / 2, 2});
        if (expected) {
          uint64_t B = (uint64_t)env.sconf.beamSize, S = (uint64_t)env.sconf.numScorers;
          uint64_t nn = st.nodeEntryPtr.size();
          uint64_t stored = nn ? st.patterns.size() / nn : 0;
          w.add("exp_beam_lp", U16, st.beamLP, {nn, B, 2});
          w.add("exp_beam_total", F32, st.beamTotal, {nn, B});
          w.add("exp_beam_scores", F32, st.beamScores, {nn, B, S});
          w.add("exp_patterns", U64, st.patterns, {nn, stored});
        }
        char name[512];
        if (nchunks == 1) {
          std::snprintf(name, sizeof name, "%s", out.c_str());
        } else {
          std::snprintf(name, sizeof name, "%s.%05ld", out.c_str(), c);
        }
        w.write(name);
      }
    });
  }
  for (auto& t : ts) t.join();
  std::printf("{\"staged_sentences\": %ld, \"failed\": %ld, \"chunks\": %ld}\n",
              (long)lines.size() - failed.load(), failed.load(), nchunks);
  return 0;
}

int cmdBench(const Args& a) {
  RefEnv env;
  env.init(a);
  auto lines = readLines(a.str("corpus"), a.num("max_sent", 0));
  unsigned nt = (unsigned)std::max<long>(1, a.num("threads", 1));
  int warmup = (int)a.num("warmup", 0);
  int steps = (int)std::max<long>(1, a.num("steps", a.num("repeat", 1)));
  using clk = std::chrono::steady_clock;
  std::vector<std::unique_ptr<AnalyzerImpl>> anas;
  for (unsigned t = 0; t < nt; ++t) anas.push_back(env.makeAnalyzer());
  // format=1: the default Juman output of every sentence is produced too (JumanFormat over an Analyzer
  // wrapped around the same AnalyzerImpl), inside analyze_s: the whole per-sentence work of the CLI
  const bool withFormat = a.num("format", 0) != 0;
  std::vector<std::unique_ptr<Analyzer>> wrappers;
  std::vector<std::unique_ptr<jumandic::output::JumanFormat>> formats;
  if (withFormat) {
    for (unsigned t = 0; t < nt; ++t) {
      wrappers.emplace_back(new Analyzer{});
      check(wrappers.back()->initialize(anas[t].get(), &env.sdef), "analyzer wrapper");
      formats.emplace_back(new jumandic::output::JumanFormat{});
      check(formats.back()->initialize(wrappers.back()->output()), "juman format");
    }
  }
  std::vector<size_t> outBytes(nt, 0);
  std::vector<double> stepScore, stepTotal, stepWall;
  long n = 0;
  double cs = 0;
  for (int pass = 0; pass < warmup + steps; ++pass) {
    std::vector<double> scoreSec(nt, 0), totalSec(nt, 0), checksum(nt, 0);
    std::vector<long> done(nt, 0);
    auto wall0 = clk::now();
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < nt; ++t) {
      ts.emplace_back([&, t] {
        auto* ana = anas[t].get();
        size_t lo = lines.size() * t / nt, hi = lines.size() * (t + 1) / nt;
        for (size_t i = lo; i < hi; ++i) {
          auto t0 = clk::now();
          Status s = ana->resetForInput(lines[i]);
          if (s) s = ana->prepareNodeSeeds();
          if (s) s = ana->buildLattice();
          if (s) s = ana->bootstrapAnalysis();
          auto t1 = clk::now();
          if (s) s = ana->computeScores(&env.sdef);
          if (s && withFormat) {
            s = formats[t]->format(*wrappers[t], StringPiece{});
            outBytes[t] += formats[t]->result().size();
          }
          auto t2 = clk::now();
          if (!s) continue;
          scoreSec[t] += std::chrono::duration<double>(t2 - t1).count();
          totalSec[t] += std::chrono::duration<double>(t2 - t0).count();
          done[t] += 1;
          auto lat = ana->lattice();
          auto& top = lat->boundary(lat->createdBoundaryCount() - 1)->starts()->beamData().row(0).at(0);
          checksum[t] += top.totalScore;
        }
      });
    }
    for (auto& t : ts) t.join();
    double wall = std::chrono::duration<double>(clk::now() - wall0).count();
    if (pass < warmup) continue;
    double maxScore = 0, maxTotal = 0;
    n = 0;
    cs = 0;
    for (unsigned t = 0; t < nt; ++t) {
      n += done[t];
      maxScore = std::max(maxScore, scoreSec[t]);
      maxTotal = std::max(maxTotal, totalSec[t]);
      cs += checksum[t];
    }
    stepScore.push_back(maxScore);
    stepTotal.push_back(maxTotal);
    stepWall.push_back(wall);
  }
  double sumScore = 0, sumTotal = 0, sumWall = 0;
  for (size_t i = 0; i < stepScore.size(); ++i) { sumScore += stepScore[i]; sumTotal += stepTotal[i]; sumWall += stepWall[i]; }
  double k = (double)stepScore.size();
  std::printf(
      "{\"sentences\": %ld, \"threads\": %u, \"steps\": %d, \"wall_s\": %.6f, \"score_path_s_max_thread\": %.6f, "
      "\"analyze_s_max_thread\": %.6f, \"score_path_sentences_per_s\": %.3f, "
      "\"analyze_sentences_per_s\": %.3f, \"checksum\": %.6f}\n",
      n, nt, (int)k, sumWall / k, sumScore / k, sumTotal / k, sumScore > 0 ? n * k / sumScore : 0.0,
      sumTotal > 0 ? n * k / sumTotal : 0.0, cs);
  return 0;
}

// model : write a complete model file (dictionary + perceptron + optional RNN parts, the container
//         of src/core/impl/model_io.cc) that the STOCK CLI loads with --model: the same synthetic
//         Jumandic model the other commands build in memory.  Perceptron part as
//         SoftConfidenceWeighted::exportModel writes it (scw.cc:129-151), RNN part through
//         RnnScorerGbeamFactory::makeInfo (rnn_scorer_gbeam.cc:375-399).
int cmdModel(const Args& a) {
  RefEnv env;
  env.init(a);
  model::ModelInfo nfo = env.dicOnlyInfo;
  util::CodedBuffer cbuf;
  util::serialization::Saver svr{&cbuf};
  PerceptronInfo pi;
  pi.modelSizeExponent = (u32)a.num("log2w", 20);
  svr.save(pi);
  model::ModelPart perc;
  perc.kind = model::ModelPartKind::Perceprton;
  perc.comment = "synthetic";
  perc.data.push_back(cbuf.contents());
  perc.data.push_back(StringPiece{reinterpret_cast<const char*>(env.weights),
                                  reinterpret_cast<const char*>(env.weights) + env.numWeights * sizeof(float)});
  nfo.parts.push_back(std::move(perc));
  if (env.hasRnn) check(env.rnnFactory.makeInfo(&nfo, "synthetic"), "rnn makeInfo");
  model::ModelSaver saver;
  check(saver.open(a.str("out")), "saver open");
  check(saver.save(nfo), "saver save");
  std::printf("{\"model\": \"%s\", \"parts\": %zu}\n", a.str("out").c_str(), nfo.parts.size());
  return 0;
}

int cmdSpec(const Args& a) {
  RefEnv env;
  env.init(a);
  SpecArrays sa;
  sa.fill(env);
  JppbWriter w;
  sa.addTo(w);
  w.write(a.str("out"));
  return 0;
}

void usage() {
  std::cerr << "usage: jpp_ref_harness <stage|bench|spec> key=value...\n"
               "  common: dict=CSV log2w=20 wseed=1 wscale=0.5 beam=5 gbeam=6 rcheck=1 rbeam=5\n"
               "          [rnn=PREFIX rnn_weight=0.0176 perc_weight=1 nce_bias= unk_const= unk_len=] [dynamic=1]\n"
               "  stage : corpus=TXT out=FILE [expected=1] [chunk=N] [threads=T] [max_sent=N]\n"
               "  bench : corpus=TXT [threads=T] [repeat=R] [max_sent=N]\n"
               "  spec  : out=FILE\n"
               "  model : out=FILE   (a model file the stock jumanpp_v2 --model loads)\n";
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    usage();
    return 2;
  }
  Args a;
  for (int i = 2; i < argc; ++i) {
    std::string s = argv[i];
    auto p = s.find('=');
    if (p == std::string::npos) die("bad argument " + s);
    a.kv[s.substr(0, p)] = s.substr(p + 1);
  }
  std::string cmd = argv[1];
  if (cmd == "stage") return cmdStage(a);
  if (cmd == "bench") return cmdBench(a);
  if (cmd == "spec") return cmdSpec(a);
  if (cmd == "model") return cmdModel(a);
  usage();
  return 2;
}
