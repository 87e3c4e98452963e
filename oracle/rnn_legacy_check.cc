// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the shipped product.
//
// Independent check of the RNN arithmetic: the reference vendors
// a second, Eigen-free implementation of the same language model,
// src/rnn/legacy/rnnlmlib_static.cpp (maxent hash :111-137, forward pass :636-704,
// step :727-775).  This driver scores the same random word sequences
//   (a) through that legacy code, compiled unmodified from /root/reference, and
//   (b) through the reference's MikolovRnn -- apply() (mikolov_rnn_impl.h:134-181) and
//       the batched computeNewParCtx() + applyParallel() pair that the global-beam scorer
//       uses (:202-257) -- compiled unmodified against oracle/eigen_shim,
// with the semantics of the reference's own comparison (mikolov_rnn_test.cc:240-277,
// 300-327): real shifting maxent history, log10 scores.  It prints the largest absolute
// differences as one JSON line; tests/test_rnn_shim_vs_legacy.py holds them to the
// reference test's 1e-3 and reports what they really are.
//
// The known-answer fixture of those tests (test/rnn/testlm) is a configure-time download
// and absent here, so the model is synthetic (jumanpp_b200/synth.py:write_rnn_model).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "rnn/legacy/rnnlmlib_static.h"
#include "rnn/mikolov_rnn.h"
#include "util/logging.hpp"
#include "util/memory.hpp"
#include "util/sliceable_array.h"

using namespace jumanpp;
using namespace jumanpp::rnn::mikolov;

namespace {

template <typename T>
util::Sliceable<T> asSlice(util::MutableArraySlice<T> mas, u32 nRows = 1) {
  return util::Sliceable<T>{mas, mas.size() / nRows, nRows};
}

struct Ctx {
  std::vector<i32> prevWords;  // newest first, at most maxentOrder - 1
  std::vector<float> context;
};

std::vector<i32> changeState(const std::vector<i32>& old, i32 one, u32 maxentOrder) {
  std::vector<i32> buf(std::min<size_t>(old.size() + 1, maxentOrder - 1));
  for (size_t i = 0; i + 1 < buf.size(); ++i) buf[i + 1] = old[i];
  if (!buf.empty()) buf[0] = one;
  return buf;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: rnn_legacy_check MODEL_PREFIX [sequences=200] [length=12] [seed=7]\n");
    return 2;
  }
  util::logging::CurrentLogLevel = util::logging::Level::Info;
  const std::string prefix = argv[1];
  const int nSeq = argc > 2 ? std::atoi(argv[2]) : 200;
  const int len = argc > 3 ? std::atoi(argv[3]) : 12;
  const unsigned seed = argc > 4 ? (unsigned)std::atoi(argv[4]) : 7u;

  // (b) the reference's reader + MikolovRnn
  MikolovModelReader rdr;
  if (!rdr.open(prefix) || !rdr.parse()) {
    std::fprintf(stderr, "cannot read model %s\n", prefix.c_str());
    return 2;
  }
  MikolovRnn rnn;
  if (!rnn.init(rdr.header(), rdr.rnnMatrix(), rdr.maxentWeights())) return 2;
  const auto& h = rdr.header();
  const u32 H = h.layerSize, V = (u32)h.vocabSize;
  std::vector<std::string> words;
  for (auto& w : rdr.words()) words.push_back(w.str());

  // (a) the legacy implementation on the same files
  RNNLM_legacy::CRnnLM_stat lm;
  lm.setDebugMode(0);
  lm.setRnnLMFile(prefix.c_str());
  RNNLM_legacy::context init;
  lm.get_initial_context_FR(&init);  // loads the model, BOS context

  util::memory::Manager mgr{4 * 1024 * 1024};
  auto alloc = mgr.core();
  util::MutableArraySlice<float> embedHack{const_cast<float*>(rdr.embeddings().data()), rdr.embeddings().size()};
  auto embeds = asSlice(embedHack, V);
  util::MutableArraySlice<float> nceHack{const_cast<float*>(rdr.nceEmbeddings().data()), rdr.nceEmbeddings().size()};
  auto nceSlice = asSlice(nceHack, V);

  std::mt19937 rng(seed);
  double maxApply = 0, maxParallel = 0, maxApplyVsParallel = 0, maxCtx = 0;
  long steps = 0;
  for (int s = 0; s < nSeq; ++s) {
    RNNLM_legacy::context lctx = init;
    Ctx c;
    c.prevWords = {0};
    c.context.assign(H, 0.0f);
    for (int t = 0; t < len; ++t) {
      // ids >= 2: id 0 (sentence end) stops the legacy maxent loop early (:133), a difference
      // in maxent history handling that the reference's own comparison avoids as well
      const i32 wid = 2 + (i32)(rng() % (V - 2));
      RNNLM_legacy::context lnext;
      const float legacy = lm.test_word_selfnm(&lctx, &lnext, words[wid], words[wid].size());

      alloc->reset();
      auto prevBuf = alloc->allocateBuf<i32>(c.prevWords.size());
      for (size_t i = 0; i < c.prevWords.size(); ++i) prevBuf[i] = c.prevWords[i];
      auto ctxBuf = alloc->allocateBuf<float>(H, 64);
      for (u32 i = 0; i < H; ++i) ctxBuf[i] = c.context[i];
      auto newCtx = alloc->allocateBuf<float>(H, 64);
      auto score = alloc->allocateBuf<float>(1, 64);
      score[0] = 0;
      i32 widMut = wid;
      util::MutableArraySlice<i32> widSlice{&widMut, 1};
      util::Sliceable<float> nceItem{nceSlice.row(wid), H, 1};
      StepData d{asSlice(prevBuf), widSlice, asSlice(ctxBuf), embeds.row(c.prevWords[0]), nceItem, asSlice(newCtx),
                 asSlice(score)};
      rnn.apply(&d);
      const float viaApply = std::log10(std::exp(score[0]));

      // the batched pair the global-beam scorer calls (one column)
      auto leftEmb = alloc->allocateBuf<float>(H, 64);
      for (u32 i = 0; i < H; ++i) leftEmb[i] = embeds.row(c.prevWords[0]).at(i);
      auto newCtx2 = alloc->allocateBuf<float>(H, 64);
      auto nceBuf = alloc->allocateBuf<float>(H, 64);
      for (u32 i = 0; i < H; ++i) nceBuf[i] = nceSlice.row(wid).at(i);
      auto score2 = alloc->allocateBuf<float>(1, 64);
      score2[0] = 0;
      ParallelContextData pcd{asSlice(ctxBuf), asSlice(leftEmb), asSlice(newCtx2)};
      rnn.computeNewParCtx(&pcd);
      ParallelStepData psd{asSlice(prevBuf), widSlice, asSlice(newCtx2), asSlice(nceBuf), score2};
      rnn.applyParallel(&psd);
      const float viaParallel = std::log10(std::exp(score2[0]));

      maxApply = std::max(maxApply, (double)std::fabs(viaApply - legacy));
      maxParallel = std::max(maxParallel, (double)std::fabs(viaParallel - legacy));
      maxApplyVsParallel = std::max(maxApplyVsParallel, (double)std::fabs(viaParallel - viaApply));
      for (u32 i = 0; i < H; ++i)
        maxCtx = std::max(maxCtx, (double)std::fabs(newCtx2[i] - lnext.l1_neuron[i]));
      ++steps;

      c.prevWords = changeState(c.prevWords, wid, h.maxentOrder);
      c.context.assign(newCtx2.begin(), newCtx2.end());
      lctx = lnext;
    }
  }
  std::printf(
      "{\"steps\": %ld, \"layer_size\": %u, \"vocab\": %u, \"maxent_order\": %u, \"max_abs_log10_apply_vs_legacy\": %.3e, "
      "\"max_abs_log10_parallel_vs_legacy\": %.3e, \"max_abs_log10_parallel_vs_apply\": %.3e, "
      "\"max_abs_context_vs_legacy\": %.3e}\n",
      steps, H, V, h.maxentOrder, maxApply, maxParallel, maxApplyVsParallel, maxCtx);
  return 0;
}
