/* jpp_oracle.h -- TEST INFRASTRUCTURE ONLY (see jpp_oracle.c). */
#ifndef JPP_ORACLE_H
#define JPP_ORACLE_H
#include <stdint.h>
#include "jpp_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct jpp_oracle_stats {
  uint64_t sentences, boundaries, nodes;
  uint64_t gathers;
  uint64_t rnn_nodes, rnn_scores;
} jpp_oracle_stats;

/* Score a staged batch on the CPU.  `weights` is the materialised perceptron
 * table (model->weights may be NULL for synthetic tables: fill one with
 * jpp_oracle_fill_weights).  result->node_beam and result->node_patterns are
 * required; node_beam_scores is required when the model has an RNN.
 * Returns 0, or a negative code (-1 bad argument, -2 internal capacity,
 * -3 beam too wide for the oracle, -4 unsupported primitive). */
int jpp_oracle_score_batch(const jpp_b200_model* model, const jpp_b200_config* cfg, const jpp_b200_batch* batch,
                           jpp_b200_result* result, const float* weights, int num_threads, jpp_oracle_stats* stats);

/* exact emulations of libstdc++ std::sort / std::nth_element and of jumanpp's
 * util::partition over u32 index arrays, comp(a,b) := score[a] > score[b] */
void jpp_oracle_std_sort(uint32_t* first, long n, const float* score);
void jpp_oracle_std_nth_element(uint32_t* first, long nth, long n, const float* score);
long jpp_oracle_ref_partition(uint32_t* first, long n, const float* score, long min_size, long max_size);

uint64_t jpp_oracle_pack_candidate(float score, uint16_t left, uint16_t beam);
float jpp_oracle_candidate_score(uint64_t packed);
float jpp_oracle_unrolled4(const float* weights, const uint32_t* idx, int n);
uint64_t jpp_oracle_mix(uint64_t state, uint64_t data);

float jpp_oracle_synth_weight(uint64_t seed, uint64_t i, float scale);
void jpp_oracle_fill_weights(float* out, uint32_t log2_size, uint64_t seed, float scale);

#ifdef __cplusplus
}
#endif
#endif
