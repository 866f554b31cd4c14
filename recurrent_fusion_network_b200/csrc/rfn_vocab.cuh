#pragma once
#include "rfn_internal.cuh"

namespace rfn {

// Device-resident bookkeeping of the batched beam search (Appendix C of SURVEY.md).
struct BeamState {
  int images, beam, L;
  int32_t* beam_seq;   // [2][images][beam][L]  ping-pong
  float* beam_lp;      // [2][images][beam][L]
  float* beam_sum;     // [images][beam]
  uint8_t* finished;   // [images]
  int32_t* done_seq;   // [images][beam*L][L]  insertion order
  float* done_lp;      // [images][beam*L][L]
  float* done_p;       // [images][beam*L]
  int32_t* n_done;     // [images]
};

int beam_merge(const BeamState& bs, int t, const float* top_val, const int32_t* top_idx, int32_t* src_row,
               int32_t* next_tok, cudaStream_t st);
int beam_finalize(const BeamState& bs, int64_t* seq, float* seq_lp, int32_t* done_seq, float* done_lp, float* done_p,
                  int32_t* n_done, cudaStream_t st);
int sample_select(const float* logits, int ld, int V, const float* rowmax, const float* logsum, const float* top_val,
                  const int32_t* top_idx, const float* uniforms, int ld_u, float temperature, int t, int L,
                  int32_t* tok_next, uint8_t* unfinished, int32_t* any_unfinished, int64_t* seq, float* seq_lp,
                  int rows, cudaStream_t st);
int sample_finalize(const int32_t* any_unfinished, int L, int32_t* d_T, cudaStream_t st);
int log_softmax_rows(const float* logits, int ld_in, float* lp, int ld_out, int rows, int V, cudaStream_t st);
int xe_loss(const float* logprobs, const int64_t* target, const float* mask, int ld_t, int rows, int T, int V,
            float eps, float* out, cudaStream_t st);
int rl_loss(const float* slp, const int64_t* seq, const float* reward, const float* lp_all, int ld_lp_rows, int rows,
            int T, int V, float entropy_reg, float* out, cudaStream_t st);

int multilabel_margin(const float* pred, const int64_t* target, int rows, int K, float weight, int accumulate,
                      float* out, cudaStream_t st);

}  // namespace rfn
