// Vocabulary-side kernels: row statistics of the logits (max, log-sum-exp, top-k with
// deterministic tie order), log-softmax materialisation, greedy / inverse-CDF token selection,
// the batched beam merge and its finalisation, and the fused criteria reductions.
#include "rfn_internal.cuh"
#include "rfn_vocab.cuh"

namespace rfn {

constexpr int VT = 256;  // threads per vocab row

__device__ __forceinline__ bool better(float v, int i, float v2, int i2) {
  return v > v2 || (v == v2 && i < i2);
}

// One CTA per row.  lp = (x - max) - log(sum exp(x - max)), the same association torch's CPU
// log_softmax uses.  top_val holds log-probs; ties resolve to the lower index (torch.max
// semantics, misc/RecurrentFusionModel.py:620; torch.sort's tie order is unspecified, :463).
template <int KMAX>
__global__ void __launch_bounds__(VT)
vocab_stats_topk_kernel(const float* __restrict__ logits, int ld, int V, int k,
                        float* __restrict__ rowmax, float* __restrict__ logsum,
                        float* __restrict__ top_val, int32_t* __restrict__ top_idx) {
  __shared__ float s_v[VT * KMAX];
  __shared__ int s_i[VT * KMAX];
  __shared__ float s_red[VT / 32];
  __shared__ int s_redi[VT / 32];
  __shared__ float s_b;
  __shared__ int s_bi;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)r * ld;

  float tv[KMAX];
  int ti[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  float m = -INFINITY;
  for (int v = tid; v < V; v += VT) {
    const float xv = x[v];
    m = fmaxf(m, xv);
    if (better(xv, v, tv[KMAX - 1], ti[KMAX - 1])) {
      tv[KMAX - 1] = xv; ti[KMAX - 1] = v;
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        if (better(tv[j], ti[j], tv[j - 1], ti[j - 1])) {
          const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
          const int fi = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = fi;
        }
      }
    }
  }
  // block max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < VT / 32; ++i) v = fmaxf(v, s_red[i]);
    s_b = v;
  }
  __syncthreads();
  m = s_b;
  float sum = 0.f;
  for (int v = tid; v < V; v += VT) sum += expf(x[v] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < VT / 32; ++i) v += s_red[i];
    s_b = logf(v);
  }
  __syncthreads();
  const float ls = s_b;
  if (tid == 0) { rowmax[r] = m; logsum[r] = ls; }
  if (k <= 0) return;

  // merge the per-thread candidate lists: k rounds of block arg-best
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { s_v[tid * KMAX + j] = tv[j]; s_i[tid * KMAX + j] = ti[j]; }
  __syncthreads();
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int c = tid; c < VT * KMAX; c += VT)
      if (better(s_v[c], s_i[c], bv, bi)) { bv = s_v[c]; bi = s_i[c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_red[warp] = bv; s_redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = s_red[0]; int i0 = s_redi[0];
      for (int i = 1; i < VT / 32; ++i)
        if (better(s_red[i], s_redi[i], v, i0)) { v = s_red[i]; i0 = s_redi[i]; }
      s_b = v; s_bi = i0;
      top_val[(size_t)r * k + round] = (v - m) - ls;
      top_idx[(size_t)r * k + round] = i0;
    }
    __syncthreads();
    const int win = s_bi;
    for (int c = tid; c < VT * KMAX; c += VT)
      if (s_i[c] == win) { s_v[c] = -INFINITY; s_i[c] = 0x7fffffff; }
    __syncthreads();
  }
}

int vocab_stats_topk(const float* logits, int ld, int rows, int V, int k, float* rowmax, float* logsum,
                     float* top_val, int32_t* top_idx, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  RFN_CHECK_ARG(logits && rowmax && logsum, "vocab_stats_topk: null pointer");
  RFN_CHECK_ARG(k >= 0 && k <= RFN_MAX_BEAM && k <= V, "vocab_stats_topk: k=%d not in 0..%d", k, RFN_MAX_BEAM);
  if (rows == 0) return RFN_OK;
  if (k <= 1)
    vocab_stats_topk_kernel<1><<<rows, VT, 0, st>>>(logits, ld, V, k, rowmax, logsum, top_val, top_idx);
  else if (k <= 4)
    vocab_stats_topk_kernel<4><<<rows, VT, 0, st>>>(logits, ld, V, k, rowmax, logsum, top_val, top_idx);
  else if (k <= 8)
    vocab_stats_topk_kernel<8><<<rows, VT, 0, st>>>(logits, ld, V, k, rowmax, logsum, top_val, top_idx);
  else
    vocab_stats_topk_kernel<16><<<rows, VT, 0, st>>>(logits, ld, V, k, rowmax, logsum, top_val, top_idx);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

__global__ void vocab_write_lp_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ rowmax,
                                      const float* __restrict__ logsum, float* __restrict__ lp, size_t ld_out, int V) {
  const int r = blockIdx.x;
  const float m = rowmax[r], ls = logsum[r];
  const float* x = logits + (size_t)r * ld;
  float* o = lp + (size_t)r * ld_out;
  for (int v = threadIdx.x; v < V; v += blockDim.x) o[v] = (x[v] - m) - ls;
}
int vocab_write_lp(const float* logits, int ld, const float* rowmax, const float* logsum, float* lp, size_t ld_out,
                   int rows, int V, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  vocab_write_lp_kernel<<<rows, 256, 0, st>>>(logits, ld, rowmax, logsum, lp, ld_out, V);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// Stand-alone log-softmax of a row (F.log_softmax, misc/RecurrentFusionModel.py:278): statistics and
// the write in one kernel, the row staying in L1/L2 between the three passes.
__global__ void __launch_bounds__(VT)
log_softmax_rows_kernel(const float* __restrict__ logits, int ld_in, float* __restrict__ lp, int ld_out, int V) {
  __shared__ float s_red[VT / 32];
  __shared__ float s_b;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)r * ld_in;
  float m = -INFINITY;
  for (int v = tid; v < V; v += VT) m = fmaxf(m, x[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < VT / 32; ++i) v = fmaxf(v, s_red[i]);
    s_b = v;
  }
  __syncthreads();
  m = s_b;
  float sum = 0.f;
  for (int v = tid; v < V; v += VT) sum += expf(x[v] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < VT / 32; ++i) v += s_red[i];
    s_b = logf(v);
  }
  __syncthreads();
  const float ls = s_b;
  float* o = lp + (size_t)r * ld_out;
  for (int v = tid; v < V; v += VT) o[v] = (x[v] - m) - ls;
}
int log_softmax_rows(const float* logits, int ld_in, float* lp, int ld_out, int rows, int V, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  RFN_CHECK_ARG(logits && lp && V > 0, "log_softmax: bad arguments");
  if (rows == 0) return RFN_OK;
  log_softmax_rows_kernel<<<rows, VT, 0, st>>>(logits, ld_in, lp, ld_out, V);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// Merge of the fused logits epilogue's per-slice statistics (one warp per row): same outputs as
// vocab_stats_topk_kernel, same association lp = (x - max) - log(sum).
__global__ void __launch_bounds__(128)
vocab_merge_kernel(const float* __restrict__ st_max, const float* __restrict__ st_sum, const float* __restrict__ st_val,
                   const int32_t* __restrict__ st_idx, int slices, int rows, int k, float* __restrict__ rowmax,
                   float* __restrict__ logsum, float* __restrict__ top_val, int32_t* __restrict__ top_idx) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (r >= rows) return;
  float m = -INFINITY;
  for (int s = lane; s < slices; s += 32) m = fmaxf(m, st_max[(size_t)s * rows + r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int s = lane; s < slices; s += 32) {
    const float ms = st_max[(size_t)s * rows + r];
    if (ms > -INFINITY) sum += st_sum[(size_t)s * rows + r] * expf(ms - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float ls = logf(sum);
  if (lane == 0) { rowmax[r] = m; logsum[r] = ls; }
  float pv = INFINITY;
  int pi = -1;
  const int ncand = slices * k;
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < ncand; c += 32) {
      const int s = c / k, j = c % k;
      const float v = st_val[((size_t)s * rows + r) * k + j];
      const int i = st_idx[((size_t)s * rows + r) * k + j];
      const bool after = (v < pv) || (v == pv && i > pi);
      if (after && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      top_val[(size_t)r * k + round] = (bv - m) - ls;
      top_idx[(size_t)r * k + round] = bi;
    }
    pv = bv; pi = bi;
  }
}
int vocab_merge(const float* st_max, const float* st_sum, const float* st_val, const int32_t* st_idx, int slices, int rows,
                int k, float* rowmax, float* logsum, float* top_val, int32_t* top_idx, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  RFN_CUDA(launch_pdl(vocab_merge_kernel, dim3((rows + 3) / 4), dim3(128), 0, st, st_max, st_sum, st_val, st_idx, slices, rows, k, rowmax,
                      logsum, top_val, top_idx));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- sample(): token selection + bookkeeping (misc/RecurrentFusionModel.py:616-649) ------------
// Greedy: it = argmax (top-1).  Multinomial: inverse CDF of p = exp(lp / temperature) in index
// order against an externally supplied uniform, accumulated in fp64 like the oracle.
__global__ void __launch_bounds__(VT)
sample_select_kernel(const float* __restrict__ logits, int ld, int V, const float* __restrict__ rowmax,
                     const float* __restrict__ logsum, const float* __restrict__ top_val,
                     const int32_t* __restrict__ top_idx, const float* __restrict__ uniforms, int ld_u,
                     float temperature, int t, int L, int32_t* __restrict__ tok_next,
                     uint8_t* __restrict__ unfinished, int32_t* __restrict__ any_unfinished,
                     int64_t* __restrict__ seq, float* __restrict__ seq_lp) {
  __shared__ double s_pref[VT + 1];
  __shared__ int s_cnt[VT / 32];
  __shared__ int s_tok;
  const int r = blockIdx.x, tid = threadIdx.x;
  int it;
  float slp;
  if (uniforms == nullptr) {
    it = top_idx[r];
    slp = top_val[r];
  } else {
    const float* x = logits + (size_t)r * ld;
    const float m = rowmax[r], ls = logsum[r];
    const int seg = (V + VT - 1) / VT;
    const int v0 = min(V, tid * seg), v1 = min(V, v0 + seg);
    double loc = 0.0;
    for (int v = v0; v < v1; ++v) {
      const float lp = (x[v] - m) - ls;
      const float p = (temperature == 1.0f) ? expf(lp) : expf(lp / temperature);
      loc += (double)p;
    }
    s_pref[tid + 1] = loc;
    __syncthreads();
    if (tid == 0) {
      s_pref[0] = 0.0;
      for (int i = 1; i <= VT; ++i) s_pref[i] += s_pref[i - 1];
    }
    __syncthreads();
    const double thr = (double)uniforms[(size_t)r * ld_u + (t - 1)] * s_pref[VT];
    double run = s_pref[tid];
    int cnt = 0;
    for (int v = v0; v < v1; ++v) {
      const float lp = (x[v] - m) - ls;
      const float p = (temperature == 1.0f) ? expf(lp) : expf(lp / temperature);
      run += (double)p;
      cnt += (run <= thr) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
      int c = 0;
      for (int i = 0; i < VT / 32; ++i) c += s_cnt[i];
      s_tok = min(c, V - 1);
    }
    __syncthreads();
    it = s_tok;
    slp = (x[it] - m) - ls;
  }
  if (tid == 0) {
    const bool un = (t == 1 ? true : unfinished[r] != 0) && (it > 0);
    unfinished[r] = un ? 1 : 0;
    if (un) atomicOr(&any_unfinished[t], 1);
    tok_next[r] = it;                                   // embed() sees the UNMASKED token (:637)
    seq[(size_t)r * L + (t - 1)] = un ? (int64_t)it : 0;  // it * unfinished (:647)
    seq_lp[(size_t)r * L + (t - 1)] = slp;              // not masked (:649)
  }
}

__global__ void sample_finalize_kernel(const int32_t* __restrict__ any_unfinished, int L, int32_t* __restrict__ d_T) {
  int T = L;
  for (int t = 1; t <= L; ++t)
    if (any_unfinished[t] == 0) { T = t - 1; break; }
  *d_T = T;
}

int sample_select(const float* logits, int ld, int V, const float* rowmax, const float* logsum, const float* top_val,
                  const int32_t* top_idx, const float* uniforms, int ld_u, float temperature, int t, int L,
                  int32_t* tok_next, uint8_t* unfinished, int32_t* any_unfinished, int64_t* seq, float* seq_lp,
                  int rows, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  sample_select_kernel<<<rows, VT, 0, st>>>(logits, ld, V, rowmax, logsum, top_val, top_idx, uniforms, ld_u,
                                            temperature, t, L, tok_next, unfinished, any_unfinished, seq, seq_lp);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}
int sample_finalize(const int32_t* any_unfinished, int L, int32_t* d_T, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  sample_finalize_kernel<<<1, 1, 0, st>>>(any_unfinished, L, d_T);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- beam merge (misc/RecurrentFusionModel.py:465-514), one thread per image: beam_merge_image in rfn_vocab.cuh ----
__global__ void beam_merge_kernel(BeamState bs, int t, const float* __restrict__ top_val,
                                  const int32_t* __restrict__ top_idx, int32_t* __restrict__ src_row,
                                  int32_t* __restrict__ next_tok) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (k >= bs.images) return;
  beam_merge_image(bs, k, t, top_val, top_idx, src_row, next_tok);
}

// stable sort of the finished beams by -p, best -> seq/seq_lp (:529-541)
__global__ void beam_finalize_kernel(BeamState bs, int64_t* __restrict__ seq, float* __restrict__ seq_lp,
                                     int32_t* __restrict__ out_done_seq, float* __restrict__ out_done_lp,
                                     float* __restrict__ out_done_p, int32_t* __restrict__ out_n_done) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= bs.images) return;
  const int L = bs.L, cap = bs.beam * L;
  const int n = bs.n_done[k];
  int order[RFN_MAX_BEAM * 64];
  const float* p = bs.done_p + (size_t)k * cap;
  for (int i = 0; i < n; ++i) {
    int j = i - 1;
    while (j >= 0 && p[order[j]] < p[i]) { order[j + 1] = order[j]; --j; }
    order[j + 1] = i;
  }
  if (out_n_done) out_n_done[k] = n;
  for (int i = 0; i < n; ++i) {
    const int s = order[i];
    const int32_t* ds = bs.done_seq + ((size_t)k * cap + s) * L;
    const float* dl = bs.done_lp + ((size_t)k * cap + s) * L;
    if (i == 0)
      for (int j = 0; j < L; ++j) { seq[(size_t)k * L + j] = ds[j]; seq_lp[(size_t)k * L + j] = dl[j]; }
    if (out_done_seq)
      for (int j = 0; j < L; ++j) out_done_seq[((size_t)k * cap + i) * L + j] = ds[j];
    if (out_done_lp)
      for (int j = 0; j < L; ++j) out_done_lp[((size_t)k * cap + i) * L + j] = dl[j];
    if (out_done_p) out_done_p[(size_t)k * cap + i] = p[s];
  }
  if (n == 0)
    for (int j = 0; j < L; ++j) { seq[(size_t)k * L + j] = 0; seq_lp[(size_t)k * L + j] = 0.f; }
}

int beam_merge(const BeamState& bs, int t, const float* top_val, const int32_t* top_idx, int32_t* src_row,
               int32_t* next_tok, cudaStream_t st) {
  ProfScope prof__(TAG_BEAM, st);
  if (bs.images == 0) return RFN_OK;
  RFN_CUDA(launch_pdl(beam_merge_kernel, dim3((bs.images + 63) / 64), dim3(64), 0, st, bs, t, top_val, top_idx, src_row, next_tok));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}
int beam_finalize(const BeamState& bs, int64_t* seq, float* seq_lp, int32_t* done_seq, float* done_lp, float* done_p,
                  int32_t* n_done, cudaStream_t st) {
  ProfScope prof__(TAG_BEAM, st);
  if (bs.images == 0) return RFN_OK;
  beam_finalize_kernel<<<(bs.images + 63) / 64, 64, 0, st>>>(bs, seq, seq_lp, done_seq, done_lp, done_p, n_done);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- fused criteria reductions (misc/utils.py:161-184, :50-72) ---------------------------------
// one CTA per (row, t): needs only lp[y] and sum_v lp_v (XE) or sum_v p log p (RL entropy)
__global__ void __launch_bounds__(VT)
xe_loss_kernel(const float* __restrict__ lp, size_t ld_b, size_t ld_s, const int64_t* __restrict__ target,
               const float* __restrict__ mask, int ld_t, int T, int V, float eps, float inv_rows, float* __restrict__ out) {
  __shared__ float s_red[VT / 32];
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const float mk = mask[(size_t)b * ld_t + t];
  if (mk == 0.f) return;
  const float* x = lp + (size_t)b * ld_b + (size_t)t * ld_s;
  float sum = 0.f;
  if (eps > 0.f) {
    for (int v = threadIdx.x; v < V; v += VT) sum += x[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float tot = 0.f;
    if (eps > 0.f)
      for (int i = 0; i < VT / 32; ++i) tot += s_red[i];
    long long y = target[(size_t)b * ld_t + t];
    y = y < 0 ? 0 : (y >= V ? V - 1 : y);
    const float term = (1.f - eps) * x[y] + (eps / (float)V) * tot;
    atomicAdd(out, -term * mk * inv_rows);
  }
}

int xe_loss(const float* logprobs, size_t ld_b, size_t ld_s, const int64_t* target, const float* mask, int ld_t, int rows, int T,
            int V, float eps, float* out, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  RFN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (rows * T == 0) return RFN_OK;
  xe_loss_kernel<<<rows * T, VT, 0, st>>>(logprobs, ld_b, ld_s, target, mask, ld_t, T, V, eps, 1.f / (float)rows, out);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

__global__ void __launch_bounds__(VT)
rl_loss_kernel(const float* __restrict__ slp, const int64_t* __restrict__ seq, const float* __restrict__ reward,
               const float* __restrict__ lp_all, size_t ld_lp_rows, size_t ld_lp_s, int T, int V, float entropy_reg,
               float inv_rows, float* __restrict__ out) {
  __shared__ float s_red[VT / 32];
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const bool m0 = seq[(size_t)b * T + t] > 0;
  const bool m = (t == 0) ? true : (seq[(size_t)b * T + t - 1] > 0);
  float ent = 0.f;
  if (m0 && entropy_reg != 0.f) {
    const float* x = lp_all + (size_t)b * ld_lp_rows + (size_t)t * ld_lp_s;
    for (int v = threadIdx.x; v < V; v += VT) { const float l = x[v]; ent += l * expf(l); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ent;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float tot = 0.f;
    if (m0 && entropy_reg != 0.f)
      for (int i = 0; i < VT / 32; ++i) tot += s_red[i];
    float v = entropy_reg * tot;
    if (m) v += -slp[(size_t)b * T + t] * reward[(size_t)b * T + t];
    if (v != 0.f) atomicAdd(out, v * inv_rows);
  }
}

int rl_loss(const float* slp, const int64_t* seq, const float* reward, const float* lp_all, size_t ld_lp_rows, size_t ld_lp_s,
            int rows, int T, int V, float entropy_reg, float* out, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  RFN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (rows * T == 0) return RFN_OK;
  rl_loss_kernel<<<rows * T, VT, 0, st>>>(slp, seq, reward, lp_all, ld_lp_rows, ld_lp_s, T, V, entropy_reg, 1.f / (float)rows, out);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}


// nn.MultiLabelMarginLoss, mean reduction (misc/utils.py:188): per row
//   sum_{j in targets} sum_{i not in targets} max(0, 1 - (x[y_j] - x[i])) / K, averaged over rows.
// target rows hold class ids first and are -1 terminated.
__global__ void __launch_bounds__(VT)
multilabel_margin_kernel(const float* __restrict__ x, const int64_t* __restrict__ target, int K, float scale,
                         float* __restrict__ out) {
  extern __shared__ unsigned char s_is[];  // K flags
  __shared__ float s_red[VT / 32];
  __shared__ int s_nt;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* xr = x + (size_t)r * K;
  const int64_t* tr = target + (size_t)r * K;
  for (int i = tid; i < K; i += VT) s_is[i] = 0;
  if (tid == 0) {
    int n = 0;
    while (n < K && tr[n] >= 0) ++n;
    s_nt = n;
  }
  __syncthreads();
  const int nt = s_nt;
  for (int j = tid; j < nt; j += VT) s_is[(int)tr[j]] = 1;
  __syncthreads();
  float acc = 0.f;
  for (int i = tid; i < K; i += VT) {
    if (s_is[i]) continue;
    const float xi = xr[i];
    for (int j = 0; j < nt; ++j) {
      const float v = 1.f - (xr[(int)tr[j]] - xi);
      acc += v > 0.f ? v : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < VT / 32; ++i) t += s_red[i];
    atomicAdd(out, t * scale);
  }
}

int multilabel_margin(const float* pred, const int64_t* target, int rows, int K, float weight, int accumulate,
                      float* out, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (!accumulate) RFN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (rows == 0) return RFN_OK;
  multilabel_margin_kernel<<<rows, VT, (size_t)K, st>>>(pred, target, K, weight / ((float)K * (float)rows), out);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // namespace rfn

extern "C" int rfn_log_softmax_f32(const float* logits, int ld_in, float* lp, int ld_out, int rows, int V,
                                   rfn_stream_t stream) {
  return rfn::log_softmax_rows(logits, ld_in, lp, ld_out, rows, V, (cudaStream_t)stream);
}

extern "C" int rfn_xe_loss_f32(const float* logprobs, const int64_t* target, const float* mask, int ld_t, int rows,
                               int T, int V, float eps, float* out, rfn_stream_t stream) {
  RFN_CHECK_ARG(logprobs && target && mask && out, "rfn_xe_loss_f32: null pointer");
  return rfn::xe_loss(logprobs, (size_t)T * V, (size_t)V, target, mask, ld_t, rows, T, V, eps, out, (cudaStream_t)stream);
}
extern "C" int rfn_xe_loss_strided_f32(const float* logprobs, size_t ld_b, size_t ld_s, const int64_t* target, const float* mask,
                                       int ld_t, int rows, int T, int V, float eps, float* out, rfn_stream_t stream) {
  RFN_CHECK_ARG(logprobs && target && mask && out, "rfn_xe_loss_strided_f32: null pointer");
  return rfn::xe_loss(logprobs, ld_b, ld_s, target, mask, ld_t, rows, T, V, eps, out, (cudaStream_t)stream);
}
extern "C" int rfn_rl_loss_f32(const float* sample_logprobs, const int64_t* seq, const float* reward,
                               const float* logprobs_all, int ld_lp_rows, int rows, int T, int V, float entropy_reg,
                               float* out, rfn_stream_t stream) {
  RFN_CHECK_ARG(sample_logprobs && seq && reward && logprobs_all && out, "rfn_rl_loss_f32: null pointer");
  return rfn::rl_loss(sample_logprobs, seq, reward, logprobs_all, (size_t)ld_lp_rows, (size_t)V, rows, T, V, entropy_reg, out,
                      (cudaStream_t)stream);
}
extern "C" int rfn_rl_loss_strided_f32(const float* sample_logprobs, const int64_t* seq, const float* reward,
                                       const float* logprobs_all, size_t ld_b, size_t ld_s, int rows, int T, int V,
                                       float entropy_reg, float* out, rfn_stream_t stream) {
  RFN_CHECK_ARG(sample_logprobs && seq && reward && logprobs_all && out, "rfn_rl_loss_strided_f32: null pointer");
  return rfn::rl_loss(sample_logprobs, seq, reward, logprobs_all, ld_b, ld_s, rows, T, V, entropy_reg, out, (cudaStream_t)stream);
}

extern "C" int rfn_multilabel_margin_f32(const float* pred, const int64_t* target, int rows, int K, float weight,
                                         int accumulate, float* out, rfn_stream_t stream) {
  RFN_CHECK_ARG(pred && target && out && K >= 1 && K <= 48 * 1024, "rfn_multilabel_margin_f32: bad arguments");
  return rfn::multilabel_margin(pred, target, rows, K, weight, accumulate, out, (cudaStream_t)stream);
}

extern "C" int rfn_mean_log_softmax_f32(int n, const float* const* logits, int rows, int V, float* mean_scratch,
                                        float* lp, rfn_stream_t stream) {
  RFN_CHECK_ARG(n >= 1 && n <= 8 && logits && mean_scratch && lp, "rfn_mean_log_softmax_f32: bad arguments");
  rfn::PtrList8 pl{};
  for (int m = 0; m < n; ++m) pl.p[m] = logits[m];
  RFN_TRY(rfn::mean_logits8(pl, n, mean_scratch, (size_t)rows * V, (cudaStream_t)stream));
  return rfn::log_softmax_rows(mean_scratch, V, lp, V, rows, V, (cudaStream_t)stream);
}
