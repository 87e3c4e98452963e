// jpp_gpu_batch.h -- the reference-side binding of the B200 lattice scorer: what a maintainer of
// ku-nlp/jumanpp adds next to src/core/analysis/analyzer.{h,cc} to run scoring on the GPU.
//
// It is host C++ written against the reference's own headers (compiled here against the sources
// under /root/reference by binding/build_binding.py, nothing is copied) and against ONE foreign
// header, include/jpp_b200.h -- the C ABI of libjpp_b200.so.  No CUDA header, no torch.
//
// BatchAnalyzer is the batched sibling of core::analysis::Analyzer (analyzer.h:38-62):
//   Analyzer::analyze(input)            BatchAnalyzer::analyzeBatch(inputs[0..n))
//     resetForInput                       for every sentence, on host threads:
//     prepareNodeSeeds                      resetForInput / prepareNodeSeeds / buildLattice  (unmodified)
//     buildLattice                          + BOS beams (bootstrapAnalysis, analyzer_impl.cc:185-192)
//     bootstrapAnalysis                   stage the lattices into pinned CSR arrays (jpp_b200_batch,
//     computeScores   <- the hot path        resident form: 8 bytes per dictionary node)
//                                         jpp_b200_score_batch               <- the only scoring call
//                                         un-stage: beams / scores / `previous` links back into
//                                         each sentence's Lattice
// after which analyzer(i) is in the state OutputManager / AnalysisPath::fillIn / the output
// formats expect (analysis_result.cc:25-76), so juman_format.cc, lattice_format.cc etc. run
// untouched.  AnalyzerImpl::computeScores is never entered (binding/jumanpp_gpu_main.cc counts it).
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "core/analysis/analyzer.h"
#include "core/analysis/analyzer_impl.h"
#include "core/analysis/rnn_id_resolver.h"
#include "core/analysis/rnn_scorer_gbeam.h"
#include "core/env.h"
#include "jpp_b200.h"

namespace jumanpp {
namespace b200 {

// persistent worker threads: run(count, fn) calls fn(index, worker) for every index in [0, count)
class ParallelRunner {
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cvStart_, cvDone_;
  const std::function<void(size_t, u32)>* fn_ = nullptr;
  std::atomic<size_t> next_{0};
  size_t count_ = 0, chunk_ = 1;
  u64 generation_ = 0;
  u32 running_ = 0;
  bool stop_ = false;

  void worker(u32 id);
  void drain(u32 id);

 public:
  explicit ParallelRunner(u32 numThreads);
  ~ParallelRunner();
  u32 size() const { return static_cast<u32>(threads_.size()) + 1; }
  void run(size_t count, const std::function<void(size_t, u32)>& fn, size_t chunk = 0);
};

// pinned host array that grows on demand (jpp_b200_alloc_pinned)
template <typename T>
class PinnedArray {
  T* p_ = nullptr;
  size_t cap_ = 0;

 public:
  PinnedArray() = default;
  PinnedArray(const PinnedArray&) = delete;
  ~PinnedArray() { jpp_b200_free_pinned(p_); }
  Status ensure(size_t n) {
    if (n <= cap_) return Status::Ok();
    size_t want = n + n / 4 + 64;
    void* np = nullptr;
    if (jpp_b200_alloc_pinned(want * sizeof(T), &np) != JPP_B200_OK) {
      return JPPS_INVALID_STATE << "failed to allocate " << want * sizeof(T) << " bytes of pinned host memory";
    }
    jpp_b200_free_pinned(p_);
    p_ = static_cast<T*>(np);
    cap_ = want;
    return Status::Ok();
  }
  T* data() { return p_; }
  T& operator[](size_t i) { return p_[i]; }
};

struct BatchOptions {
  int device = 0;
  u32 maxBatch = 8192;    // sentences per jpp_b200_score_batch call (one resident AnalyzerImpl each)
  u32 threads = 0;        // host threads for lattice building / staging / un-staging (0 = hardware)
  bool rnnCliConfigApplied = false;  // the command line carried RNN options (JumanppExec::init then calls setRnnConfig)
  bool fullBeams = false; // bring back every beam row and connection score (N-best / lattice output);
                          // otherwise only the 1-best chain is written back
};

struct BatchTimings {
  double buildSec = 0, stageSec = 0, tableSec = 0, scoreSec = 0, unstageSec = 0;
  u64 sentences = 0, nodes = 0, failed = 0, tableRowsAdded = 0, h2dBytes = 0, d2hBytes = 0;
};

class BatchAnalyzer {
 public:
  BatchAnalyzer();
  ~BatchAnalyzer();
  BatchAnalyzer(const BatchAnalyzer&) = delete;

  // `env` is a loaded reference environment (model, beams, RNN configuration already applied,
  // features initialised); `rnnFile` is the external --rnn-model prefix when one was given.
  Status initialize(const core::JumanppEnv& env, const core::analysis::AnalyzerConfig& acfg,
                    const core::ScoringConfig& sconf, StringPiece rnnFile, const BatchOptions& opts);

  // inputs.size() <= maxBatch().  On return analyzer(i) holds sentence i (status(i) says whether
  // it could be analysed; a failed sentence is the caller's `# ERROR`).
  Status analyzeBatch(const std::vector<StringPiece>& inputs) {
    JPP_RETURN_IF_ERROR(prepare(0, inputs));
    JPP_RETURN_IF_ERROR(score(0));
    return finish(0);
  }

  // The same in three steps over two sets of resident analyzers ("slots"), so that a caller can build
  // and stage batch k+1 and format batch k-1 on the host threads while the GPU scores batch k:
  //   prepare(slot)  lattices built and staged, new dictionary rows made resident   (host threads)
  //   score(slot)    jpp_b200_score_batch; may run on another thread, concurrently with prepare /
  //                  finish of the OTHER slot
  //   finish(slot)   beams / paths written back into the lattices                   (host threads)
  Status prepare(int slot, const std::vector<StringPiece>& inputs);
  Status score(int slot);
  Status finish(int slot);

  u32 maxBatch() const { return opts_.maxBatch; }
  core::analysis::Analyzer* analyzer(size_t i, int slot = 0) { return slots_[slot].analyzers[i].get(); }
  const Status& status(size_t i, int slot = 0) const { return slots_[slot].statuses[i]; }
  ParallelRunner& runner() { return *runner_; }
  const BatchTimings& timings() const { return timings_; }
  const core::analysis::ScorerDef* scorerDef() const { return &scorerDef_; }
  static constexpr int kSlots = 2;

 private:
  struct SentenceShape {
    u32 boundaries = 0, nodes = 0, ends = 0, codepoints = 0, extras = 0;
  };

  // one set of resident analyzers with its staged batch (pinned) and its result (pinned)
  struct Slot {
    std::vector<std::unique_ptr<core::analysis::AnalyzerImpl>> impls;
    std::vector<std::unique_ptr<core::analysis::Analyzer>> analyzers;
    std::vector<Status> statuses;
    std::vector<SentenceShape> shapes;
    std::vector<u32> sentBndOff, sentCpOff, sentExtraOff, sentNodeOff, sentEndOff;
    PinnedArray<u32> hSentBndOff, hSentCpOff, hSentExtraOff, hBndNodeOff, hBndEndOff;
    PinnedArray<jpp_b200_node_info> hNodeInfo;
    PinnedArray<jpp_b200_node_ptr> hEnds;
    PinnedArray<jpp_b200_codepoint> hCps;
    PinnedArray<i32> hExtraRows, hExtraPh, hExtraRnn;
    PinnedArray<jpp_b200_beam_element> hBeam;
    PinnedArray<float> hBeamScores;
    PinnedArray<u32> hPathLen, hPathNodes;
    PinnedArray<u8> hStatus;
    size_t batchSize = 0, NB = 0, NN = 0, NE = 0, NC = 0, NX = 0;
  };

  Status buildModel(const core::JumanppEnv& env, StringPiece rnnFile);
  Status makeAnalyzers();
  void buildOne(Slot& s, size_t i, StringPiece input);
  void stageOne(Slot& s, size_t i, u32 worker);
  Status addMissingRows();
  void unstageOne(Slot& s, size_t i);
  i32 resolveRnnId(core::analysis::rnn::RnnReprBuilder* bldr, core::EntryPtr eptr, util::ArraySlice<i32> entry,
                   const core::analysis::ExtraNodesContext* xtra) const;

  BatchOptions opts_;
  const core::JumanppEnv* env_ = nullptr;
  core::analysis::AnalyzerConfig acfg_;
  core::ScoringConfig sconf_;
  core::analysis::ScorerDef scorerDef_;  // weights as the environment's, no ScoreComputer instances
  std::unique_ptr<ParallelRunner> runner_;
  Slot slots_[kSlots];
  BatchTimings timings_;
  std::mutex timingsMutex_;

  // device
  jpp_b200_handle* handle_ = nullptr;
  u32 entrySize_ = 0, numPlaceholders_ = 0, beam_ = 0, numScorers_ = 1, storedPatterns_ = 0;
  bool hasRnn_ = false;
  // RNN id resolution (RnnIdContainer::resolveId, rnn_id_resolver.cc:278-309)
  core::analysis::rnn::RnnIdResolver resolver_;
  core::dic::DoubleArray knownIdx_, unkIdx_;

  // resident entry table bookkeeping: which dictionary entries are already on the device
  std::vector<u8> resident_;                       // [entry_ptr] 1 = row is on the device
  std::vector<std::vector<i32>> missingPerWorker_;  // entry_ptrs met for the first time in this batch
};

}  // namespace b200
}  // namespace jumanpp
