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

// ---- beam merge of ONE image (misc/RecurrentFusionModel.py:465-514), run by one thread ------------------------------
struct Cand { int c; int q; float p; float r; };

__device__ inline void beam_merge_image(const BeamState& bs, int k, int t, const float* __restrict__ top_val,
                                        const int32_t* __restrict__ top_idx, int32_t* __restrict__ src_row,
                                        int32_t* __restrict__ next_tok) {
  const int beam = bs.beam, L = bs.L, cap = beam * L;
  const int base = k * beam;
  const int cur = (t - 1) & 1, nxt = t & 1;  // ping-pong copies of the per-beam sequences
  const size_t plane = (size_t)bs.images * beam * L;
  const int32_t* seq_c = bs.beam_seq + cur * plane + (size_t)base * L;
  int32_t* seq_n = bs.beam_seq + nxt * plane + (size_t)base * L;
  const float* lp_c = bs.beam_lp + cur * plane + (size_t)base * L;
  float* lp_n = bs.beam_lp + nxt * plane + (size_t)base * L;
  float* sum = bs.beam_sum + base;

  auto idle = [&]() {
    for (int v = 0; v < beam; ++v) { src_row[base + v] = base + v; next_tok[base + v] = 0; }
    // keep both sequence copies coherent so later parity flips stay harmless
    for (int v = 0; v < beam; ++v)
      for (int i = 0; i < L; ++i) { seq_n[v * L + i] = seq_c[v * L + i]; lp_n[v * L + i] = lp_c[v * L + i]; }
  };
  if (bs.finished[k]) { idle(); return; }

  Cand cand[RFN_MAX_BEAM * RFN_MAX_BEAM];
  int n = 0;
  const int nq = (t == 1) ? 1 : beam;                       // :468-469
  for (int c = 0; c < beam; ++c)                            // c OUTER :470
    for (int q = 0; q < nq; ++q) {                          // q INNER :471
      if (t > 1 && seq_c[q * L + (t - 2)] == 0) continue;   // :475
      const float local = top_val[(size_t)(base + q) * beam + c];
      Cand cd;
      cd.c = top_idx[(size_t)(base + q) * beam + c];
      cd.q = q;
      cd.p = __fadd_rn(sum[q], local);                      // fp32 add :474
      cd.r = local;
      cand[n++] = cd;
    }
  if (n == 0) { bs.finished[k] = 1; idle(); return; }       // :480-481
  // stable insertion sort by -p (:482)
  for (int i = 1; i < n; ++i) {
    const Cand x = cand[i];
    int j = i - 1;
    while (j >= 0 && cand[j].p < x.p) { cand[j + 1] = cand[j]; --j; }
    cand[j + 1] = x;
  }
  float new_sum[RFN_MAX_BEAM];
  for (int v = 0; v < beam; ++v) new_sum[v] = sum[v];
  const int nv = n < beam ? n : beam;
  for (int v = 0; v < beam; ++v) {
    if (v < nv) {
      const Cand cd = cand[v];
      for (int i = 0; i < t - 1; ++i) { seq_n[v * L + i] = seq_c[cd.q * L + i]; lp_n[v * L + i] = lp_c[cd.q * L + i]; }
      seq_n[v * L + (t - 1)] = cd.c;                        // :504
      lp_n[v * L + (t - 1)] = cd.r;                         // :505
      for (int i = t; i < L; ++i) { seq_n[v * L + i] = 0; lp_n[v * L + i] = 0.f; }
      new_sum[v] = cd.p;                                    // :506
      src_row[base + v] = base + cd.q;                      // :499-501
      next_tok[base + v] = cd.c;
      if (cd.c == 0 || t == L) {                            // :508
        const int d = bs.n_done[k];
        if (d < cap) {
          int32_t* ds = bs.done_seq + ((size_t)k * cap + d) * L;
          float* dl = bs.done_lp + ((size_t)k * cap + d) * L;
          for (int i = 0; i < L; ++i) { ds[i] = seq_n[v * L + i]; dl[i] = lp_n[v * L + i]; }
          bs.done_p[(size_t)k * cap + d] = cd.p;
          bs.n_done[k] = d + 1;
        }
      }
    } else {  // unreachable in the reference (len(cand) >= beam); keep the old beam in place
      for (int i = 0; i < L; ++i) { seq_n[v * L + i] = seq_c[v * L + i]; lp_n[v * L + i] = lp_c[v * L + i]; }
      src_row[base + v] = base + v;
      next_tok[base + v] = (t >= 2) ? seq_c[v * L + (t - 2)] : 0;
    }
  }
  for (int v = 0; v < beam; ++v) sum[v] = new_sum[v];
}

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
// lp[b, t, :] at logprobs + b * ld_b + t * ld_s (contiguous (rows, T, V): ld_b = T * V, ld_s = V; time-major: ld_b = V, ld_s = rows * V)
int xe_loss(const float* logprobs, size_t ld_b, size_t ld_s, const int64_t* target, const float* mask, int ld_t, int rows, int T,
            int V, float eps, float* out, cudaStream_t st);
int rl_loss(const float* slp, const int64_t* seq, const float* reward, const float* lp_all, size_t ld_lp_rows, size_t ld_lp_s,
            int rows, int T, int V, float entropy_reg, float* out, cudaStream_t st);

int multilabel_margin(const float* pred, const int64_t* target, int rows, int K, float weight, int accumulate,
                      float* out, cudaStream_t st);

}  // namespace rfn
