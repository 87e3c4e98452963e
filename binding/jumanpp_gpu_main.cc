// jumanpp_gpu_main.cc -- `jumanpp_v2_gpu`: the reference CLI (src/jumandic/main/jumanpp.cc) with the
// scoring of plain-text input moved to the B200.  Every option of the stock binary keeps its meaning
// (they are parsed by the reference's own jumandic::parseArgs); the model is loaded, the beams and
// the RNN configured and the output format chosen by the reference's own JumanppExec.  What changes:
// sentences are read in batches, their lattices are built on host threads by the unmodified
// AnalyzerImpl pipeline, scored by ONE jpp_b200_score_batch call per batch (binding/jpp_gpu_batch.cc),
// and formatted -- on host threads, by the reference's formatters -- in input order.
//
// Extra options (removed from argv before the reference parser sees it):
//   --gpu-device=N      CUDA device (default 0)
//   --gpu-batch=N       sentences per scoring call (default 8192; one resident analyzer each)
//   --gpu-threads=N     host threads (default: all)
//   --gpu-stats         timing / counters as one JSON line on stderr at exit
//   --gpu=off           run the stock per-sentence path (also taken for partially annotated input,
//                       --partial, and for --auto-beam: ScorePlugin callbacks and per-sentence beam
//                       retries do not cross the C ABI)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iostream>
#include <string>
#include <vector>

#include "core/analysis/analyzer_impl.h"
#include "core/impl/segmented_format.h"
#include "core/input/pex_stream_reader.h"
#include "jpp_gpu_batch.h"
#include "jumandic/shared/juman_format.h"
#include "jumandic/shared/jumandic_env.h"
#include "jumandic/shared/jumanpp_args.h"
#include "jumandic/shared/lattice_format.h"
#include "jumandic/shared/morph_format.h"
#include "jumandic/shared/subset_format.h"
#include "util/logging.hpp"

using namespace jumanpp;

// ---- proof that the CPU scorer is not on the path: every call of AnalyzerImpl::computeScores in
// this binary goes through this counter (ld --wrap on the member function's symbol; the ABI of a
// member function returning Status equals that of a free function taking `this` first)
static std::atomic<unsigned long> g_computeScoresCalls{0};
extern "C" Status __real__ZN7jumanpp4core8analysis12AnalyzerImpl13computeScoresEPKNS1_9ScorerDefE(
    core::analysis::AnalyzerImpl* self, const core::analysis::ScorerDef* def);
extern "C" Status __wrap__ZN7jumanpp4core8analysis12AnalyzerImpl13computeScoresEPKNS1_9ScorerDefE(
    core::analysis::AnalyzerImpl* self, const core::analysis::ScorerDef* def) {
  g_computeScoresCalls.fetch_add(1);
  return __real__ZN7jumanpp4core8analysis12AnalyzerImpl13computeScoresEPKNS1_9ScorerDefE(self, def);
}

namespace {

class GpuExec : public jumandic::JumanppExec {
 public:
  explicit GpuExec(const jumandic::JumanppConf& c) : jumandic::JumanppExec(c) {}
  const core::JumanppEnv& environment() const { return env; }
  const jumandic::JumanppConf& config() const { return conf; }
};

// JumanppExec::initOutput (jumandic_env.cc:60-146) for one more analyzer: a formatter keeps
// accessors into ITS analyzer's OutputManager, so every resident analyzer gets its own
Status makeFormat(const jumandic::JumanppConf& conf, const core::JumanppEnv& env, const core::analysis::Analyzer& an,
                  std::unique_ptr<core::OutputFormat>* out) {
  using jumandic::OutputType;
  switch (conf.outputType.value()) {
    case OutputType::Juman: {
      auto f = new jumandic::output::JumanFormat;
      out->reset(f);
      return f->initialize(an.output());
    }
    case OutputType::Morph:
    case OutputType::FullMorph: {
      auto f = new jumandic::output::MorphFormat(conf.outputType.value() == OutputType::FullMorph);
      out->reset(f);
      return f->initialize(an.output());
    }
    case OutputType::DicSubset: {
      auto f = new jumandic::output::SubsetFormat{};
      out->reset(f);
      return f->initialize(an.output());
    }
    case OutputType::Lattice: {
      i32 numOutput = conf.beamOutput;
      if (numOutput == -1) numOutput = conf.beamSize;
      auto f = new jumandic::output::LatticeFormat{numOutput};
      out->reset(f);
      return f->initialize(an.output());
    }
    case OutputType::Segmentation: {
      auto f = new core::output::SegmentedFormat{};
      out->reset(f);
      return f->initialize(an.output(), *env.coreHolder(), conf.segmentSeparator.value());
    }
    default:
      return JPPS_NOT_IMPLEMENTED << "this output format is not available on the GPU path";
  }
}

struct Example {
  std::string input, comment;
  Status readStatus = Status::Ok();
};

// PlainStreamReader::readExample (stream_reader.cc:12-40) + the CLI's limits (jumanpp.cc:78)
struct LineSource {
  const std::vector<std::string>* files;
  size_t nextFile = 0;
  std::unique_ptr<std::ifstream> file;
  std::istream* in = nullptr;

  bool open() {
    if (files->empty()) {
      in = &std::cin;
      return true;
    }
    return advance();
  }
  bool advance() {
    if (nextFile >= files->size()) return false;
    file.reset(new std::ifstream{(*files)[nextFile++]});
    in = file.get();
    return true;
  }
  bool hasNext() {
    for (;;) {
      if (in->good() && in->peek() != std::char_traits<char>::eof()) return true;
      if (files->empty() || !advance()) return false;
    }
  }
  void read(Example* ex) {
    ex->comment.clear();
    while (true) {
      ex->input.clear();
      std::getline(*in, ex->input);
      if (ex->input.size() > 2 && ex->input[0] == '#' && ex->input[1] == ' ') std::swap(ex->comment, ex->input);
      else break;
    }
    ex->readStatus = Status::Ok();
    if (ex->comment.size() > 1024)
      ex->readStatus = Status::InvalidParameter() << "Comment size was: " << ex->comment.size() << " which is more than max: " << 1024;
    else if (ex->input.size() > 65535)
      ex->readStatus = Status::InvalidParameter() << "Input size was: " << ex->input.size() << " which is more than max: " << 65535;
  }
};

int stockLoop(GpuExec& exec, const jumandic::JumanppConf& conf, std::ostream& out) {
  // the reference's own loop (jumanpp.cc:154-175) on the reference's own analyzer
  std::unique_ptr<core::input::StreamReader> reader;
  if (conf.inputType.value() == jumandic::InputType::Raw) {
    auto r = new core::input::PlainStreamReader{};
    reader.reset(r);
    r->setMaxSizes(65535, 1024);
  } else {
    auto r = new core::input::PexStreamReader{};
    reader.reset(r);
    Status s = r->initialize(exec.core(), '&');
    if (!s) {
      std::cerr << "Failed to initialize I/O: " << s;
      return 1;
    }
  }
  LineSource src{&conf.inputFiles.value()};
  if (!src.open()) return 1;
  int result = 0;
  while (src.hasNext()) {
    Status s = reader->readExample(src.in);
    if (!s) {
      std::cerr << "failed to read an example: " << s;
      result = 1;
      continue;
    }
    result = 0;
    s = reader->analyzeWith(exec.analyzerPtr());
    if (!s) {
      std::cerr << s;
      out << exec.emptyResult();
      continue;
    }
    s = exec.format()->format(*exec.analyzerPtr(), reader->comment());
    if (!s) std::cerr << s;
    else out << exec.format()->result();
  }
  return result;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, const char** argv) {
  b200::BatchOptions opts;
  bool stats = false, gpuOff = false;
  std::vector<const char*> args;
  for (int i = 0; i < argc; ++i) {
    const char* a = argv[i];
    if (!std::strncmp(a, "--gpu-device=", 13)) opts.device = std::atoi(a + 13);
    else if (!std::strncmp(a, "--gpu-batch=", 12)) opts.maxBatch = (u32)std::atoi(a + 12);
    else if (!std::strncmp(a, "--gpu-threads=", 14)) opts.threads = (u32)std::atoi(a + 14);
    else if (!std::strcmp(a, "--gpu-stats")) stats = true;
    else if (!std::strcmp(a, "--gpu=off")) gpuOff = true;
    else args.push_back(a);
  }
  jumandic::JumanppConf conf;
  Status s = jumandic::parseArgs((int)args.size(), args.data(), &conf);
  if (!s) {
    std::cerr << s << "\n";
    return 1;
  }
  GpuExec exec{conf};
  s = exec.init();
  if (!s.isOk()) {
    if (conf.outputType == jumandic::OutputType::Version) { exec.printFullVersion(); return 0; }
    if (conf.modelFile.isDefault()) { std::cerr << "Model file was not specified\n"; return 1; }
    if (conf.outputType == jumandic::OutputType::ModelInfo) { exec.printModelInfo(); return 0; }
    std::cerr << "failed to load model from disk: " << s;
    return 1;
  }
  if (conf.outputType == jumandic::OutputType::Version) { exec.printFullVersion(); return 0; }
  if (conf.outputType == jumandic::OutputType::ModelInfo) { exec.printModelInfo(); return 0; }

  std::unique_ptr<std::ofstream> fileOut;
  std::ostream* out = &std::cout;
  if (!(exec.config().outputFile == "-")) {
    fileOut.reset(new std::ofstream{exec.config().outputFile.value()});
    out = fileOut.get();
  }
  const jumandic::JumanppConf& cf = exec.config();  // merged with the configuration files
  const bool plain = cf.inputType.value() == jumandic::InputType::Raw;
  if (gpuOff || !plain || cf.autoStep.defined() || !cf.graphvizDir.value().empty()) {
    return stockLoop(exec, cf, *out);
  }

  const core::JumanppEnv& env = exec.environment();
  core::analysis::AnalyzerConfig acfg{};  // JumanppEnv::analyzerConfig_ as JumanppExec::init leaves it
  acfg.globalBeamSize = cf.globalBeam;
  acfg.rightGbeamCheck = cf.rightCheck;
  acfg.rightGbeamSize = cf.rightBeam;
  core::ScoringConfig sconf{(u32)cf.beamSize.value(), (u32)env.scorers()->scoreWeights.size()};
  // N-best / lattice output walks more than the 1-best chain
  opts.fullBeams = cf.outputType.value() == jumandic::OutputType::Lattice;
  opts.rnnCliConfigApplied = !cf.rnnConfig.isDefault();
  double tInit0 = now();
  b200::BatchAnalyzer batch;
  s = batch.initialize(env, acfg, sconf, cf.rnnModelFile.value(), opts);
  if (!s) {
    std::cerr << "failed to initialise the GPU scorer: " << s << "\n";
    return 1;
  }
  // every resident analyzer (two sets: one being scored, one being built / printed) gets its formatter
  std::vector<std::unique_ptr<core::OutputFormat>> formats[b200::BatchAnalyzer::kSlots];
  for (int sl = 0; sl < b200::BatchAnalyzer::kSlots; ++sl) {
    formats[sl].resize(batch.maxBatch());
    std::vector<Status> errs;
    for (u32 i = 0; i < batch.maxBatch(); ++i) errs.emplace_back(Status::Ok());
    batch.runner().run(batch.maxBatch(), [&](size_t i, u32) { errs[i] = makeFormat(cf, env, *batch.analyzer(i, sl), &formats[sl][i]); });
    for (auto& e : errs)
      if (!e) {
        std::cerr << "failed to initialise the output format: " << e << "\n";
        return 1;
      }
  }
  double tInit1 = now();

  LineSource src{&cf.inputFiles.value()};
  if (!src.open()) return 1;
  struct InFlight {
    size_t n = 0;
    std::vector<Example> examples;
    std::vector<StringPiece> inputs;
    std::vector<size_t> slot;  // example -> index in the batch (examples that could not be read take none)
    std::future<Status> scoring;
  } fl[b200::BatchAnalyzer::kSlots];
  for (auto& f : fl) f.examples.resize(batch.maxBatch());
  std::vector<Status> fmtStatus;
  for (u32 i = 0; i < batch.maxBatch(); ++i) fmtStatus.emplace_back(Status::Ok());
  double formatSec = 0, writeSec = 0, readSec = 0, total0 = now();
  int result = 0;
  const std::string empty = exec.emptyResult().str();

  // un-stage, format (in parallel: every analyzer has its own formatter) and write (in input order) one batch
  auto emit = [&](int sl) -> bool {
    InFlight& f = fl[sl];
    Status s2 = batch.finish(sl);
    if (!s2) {
      std::cerr << s2 << "\n";
      return false;
    }
    double f0 = now();
    std::vector<size_t> exampleOfSlot(f.inputs.size());
    for (size_t e = 0; e < f.n; ++e)
      if (f.slot[e] != (size_t)-1) exampleOfSlot[f.slot[e]] = e;
    batch.runner().run(f.inputs.size(), [&](size_t k, u32) {
      if (!batch.status(k, sl)) return;
      const std::string& c = f.examples[exampleOfSlot[k]].comment;
      StringPiece comment = c.size() < 2 ? EMPTY_SP : StringPiece{c}.from(2);
      fmtStatus[k] = formats[sl][k]->format(*batch.analyzer(k, sl), comment);
    }, 8);
    double f1 = now();
    for (size_t e = 0; e < f.n; ++e) {
      if (!f.examples[e].readStatus) {
        std::cerr << "failed to read an example: " << f.examples[e].readStatus;
        result = 1;
        continue;
      }
      result = 0;
      const size_t k = f.slot[e];
      if (!batch.status(k, sl)) {
        std::cerr << batch.status(k, sl);
        *out << empty;
        continue;
      }
      if (!fmtStatus[k]) std::cerr << fmtStatus[k];
      else *out << formats[sl][k]->result();
    }
    formatSec += f1 - f0;
    writeSec += now() - f1;
    return true;
  };

  // pipeline: while the GPU scores batch k (a thread of its own inside jpp_b200_score_batch), the host
  // threads print batch k-1 and build + stage batch k+1
  int cur = 0;
  bool havePrev = false;
  while (src.hasNext()) {
    InFlight& f = fl[cur];
    double r0 = now();
    f.n = 0;
    f.inputs.clear();
    f.slot.assign(batch.maxBatch(), (size_t)-1);
    while (f.n < batch.maxBatch() && src.hasNext()) {
      src.read(&f.examples[f.n]);
      if (f.examples[f.n].readStatus) {
        f.slot[f.n] = f.inputs.size();
        f.inputs.emplace_back(f.examples[f.n].input);
      }
      ++f.n;
    }
    readSec += now() - r0;
    s = batch.prepare(cur, f.inputs);
    if (!s) {
      std::cerr << s << "\n";
      return 1;
    }
    const int prev = cur ^ 1;
    if (havePrev) {
      s = fl[prev].scoring.get();
      if (!s) {
        std::cerr << s << "\n";
        return 1;
      }
    }
    f.scoring = std::async(std::launch::async, [&batch, cur] { return batch.score(cur); });
    if (havePrev && !emit(prev)) return 1;
    havePrev = true;
    cur ^= 1;
  }
  if (havePrev) {
    const int last = cur ^ 1;
    s = fl[last].scoring.get();
    if (!s) {
      std::cerr << s << "\n";
      return 1;
    }
    if (!emit(last)) return 1;
  }
  out->flush();
  double total1 = now();
  if (stats) {
    const b200::BatchTimings& t = batch.timings();
    std::fprintf(stderr,
                 "{\"gpu_stats\": {\"sentences\": %llu, \"nodes\": %llu, \"failed\": %llu, \"batch\": %u, \"threads\": %u, "
                 "\"init_s\": %.4f, \"total_s\": %.4f, \"read_s\": %.4f, \"build_lattice_s\": %.4f, \"stage_s\": %.4f, "
                 "\"entry_table_s\": %.4f, \"score_call_s\": %.4f, \"unstage_s\": %.4f, \"format_s\": %.4f, \"write_s\": %.4f, "
                 "\"entry_rows_added\": %llu, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"full_beams\": %s, "
                 "\"compute_scores_calls\": %lu}}\n",
                 (unsigned long long)t.sentences, (unsigned long long)t.nodes, (unsigned long long)t.failed, batch.maxBatch(),
                 batch.runner().size(), tInit1 - tInit0, total1 - total0, readSec, t.buildSec, t.stageSec, t.tableSec, t.scoreSec,
                 t.unstageSec, formatSec, writeSec, (unsigned long long)t.tableRowsAdded, (unsigned long long)t.h2dBytes,
                 (unsigned long long)t.d2hBytes, opts.fullBeams ? "true" : "false", g_computeScoresCalls.load());
  }
  return result;
}
