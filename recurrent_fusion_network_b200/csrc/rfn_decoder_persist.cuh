// Persistent cooperative decoder (rfn_decoder_persist.cu): argument block and host entry points.
#pragma once
#include "rfn_internal.cuh"
#include "rfn_vocab.cuh"

namespace rfn {

constexpr int PD_MAX_ROWS = 64;    // decoder rows (batch, or images x beam) the persistent kernel takes

struct PDArgs {
  const float *i2h_w, *i2h_b, *h2h_w, *h2h_b, *z2h_w, *z2h_b;   // (4R,E), (4R,R), (4R,R)
  const float *hatt_w, *hatt_b;                                 // h_2_att_h (A,R)
  const float *out_w, *out_b;                                   // att_h_2_out (A), (1)
  const float *logit_w, *logit_b;                               // (V,R)
  const float* embed;                                           // (V,E)
  const float* TVc;                                             // (rowsA,S1,R)
  const float* Pdec;                                            // (rowsA,S1,A) = att_2_att_h(TVc), hoisted
  int rows, div, R, A, E, V, S1, L;
  int UPS, OPA, VPS, nslice;
  float* hbuf[2];                                               // (rows,R) ping-pong; [0] = initial state
  float* cbuf[2];
  float *g, *z;                                                 // (rows,A), (rows,R)
  float *part_max, *part_sum, *part_val;                        // [nslice][rows] (, ktop)
  int32_t* part_idx;
  float* logits;                                                // (rows,V), only when lp_all is wanted
  float *rowmax, *logsum;                                       // (rows)
  int32_t *tok, *src;                                           // next input token / state source row (identity for greedy)
  int ktop, steps;
  // greedy bookkeeping (sample_select semantics, misc/RecurrentFusionModel.py:616-649)
  int64_t* seq;
  float* seq_lp;
  uint8_t* unfinished;
  int32_t* any_unfinished;
  float* lp_all;                                                // (rows, L+1, V) or nullptr
  // beam bookkeeping
  int beam;                                                     // 0 = greedy
  BeamState bs;
  long long* dbg;                                               // optional per-step phase stamps (rfn_debug_set_pd_timeline)
};


bool pd_supported(const rfn_dims& d, int rows);
// floats of per-slice statistics scratch (part_max, part_sum, part_val) + ints (part_idx): 2 + 2 * ktop values per (slice, row)
size_t pd_part_floats(const rfn_dims& d, int rows, int ktop);
void pd_bind_parts(const rfn_dims& d, PDArgs& a, float* base);
int pd_launch(const rfn_dims& d, PDArgs& a, cudaStream_t st);

}  // namespace rfn
