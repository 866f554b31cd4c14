// Small fused pointwise kernels of the path: LSTM cell update, embedding gather, beam state
// gather, state averaging, max over review steps.
#include "rfn_internal.cuh"

namespace rfn {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gate layout [i | f | o | g] along 4R  (misc/RecurrentFusionModel.py:55-73)
// c_prev may alias c_out (in-place cell state), so neither is __restrict__.
__global__ void lstm_cell_kernel(const float* __restrict__ G, const float* c_prev,
                                 float* __restrict__ h_out, float* c_out,
                                 float* __restrict__ h_out2, int ldh2, float* __restrict__ h_out3,
                                 int ldh3, int rows, int R, const float* __restrict__ mask, float scale, int maxout) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), k = (int)(i % R);
    const float* Gr = G + (size_t)r * (4 + maxout) * R;
    const float ig = sigmoidf_(Gr[k]);
    const float fg = sigmoidf_(Gr[R + k]);
    const float og = sigmoidf_(Gr[2 * R + k]);
    // maxout: in_transform = max of the last two R-blocks of a 5R-wide gate row, no tanh (misc/LSTMSoftAttentionCore.py:89-91)
    const float gg = maxout ? fmaxf(Gr[3 * R + k], Gr[4 * R + k]) : tanhf(Gr[3 * R + k]);
    const float c2 = fg * c_prev[i] + ig * gg;
    float h2 = og * tanhf(c2);
    if (mask) h2 = scale * h2 * mask[i];   // nn.Dropout with an explicit keep-mask (training): the state carries the dropped h
    c_out[i] = c2;
    if (h_out) h_out[i] = h2;
    if (h_out2) h_out2[(size_t)r * ldh2 + k] = h2;
    if (h_out3) h_out3[(size_t)r * ldh3 + k] = h2;
  }
}

int lstm_cell(const float* G, const float* c_prev, float* h_out, float* c_out, float* h_out2, int ldh2,
              float* h_out3, int ldh3, int rows, int R, cudaStream_t st, const float* mask, float scale, int maxout) {
  ProfScope prof__(TAG_CELL, st);
  RFN_CHECK_ARG(G && c_prev && c_out, "lstm_cell: null pointer");
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * R;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  RFN_CUDA(launch_pdl(lstm_cell_kernel, dim3(blocks), dim3(256), 0, st, G, c_prev, h_out, c_out, h_out2, ldh2, h_out3, ldh3, rows, R, mask, scale, maxout ? 1 : 0));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

template <typename TokT>
__global__ void embed_gather_kernel(const TokT* __restrict__ tok, int ld_tok, const float* __restrict__ embed,
                                    float* __restrict__ x, int rows, int E, int V1) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)rows * E;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / E), k = (int)(i % E);
    long long t = (long long)tok[(size_t)r * ld_tok];
    t = t < 0 ? 0 : (t >= V1 ? V1 - 1 : t);
    x[i] = __ldg(embed + (size_t)t * E + k);
  }
}

int embed_gather_i64(const int64_t* tok, int ld_tok, const float* embed, float* x, int rows, int E, int V1,
                     cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * E;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  RFN_CUDA(launch_pdl(embed_gather_kernel<int64_t>, dim3(blocks), dim3(256), 0, st, tok, ld_tok, embed, x, rows, E, V1));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}
int embed_gather_i32(const int32_t* tok, const float* embed, float* x, int rows, int E, int V1, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * E;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  RFN_CUDA(launch_pdl(embed_gather_kernel<int32_t>, dim3(blocks), dim3(256), 0, st, tok, 1, embed, x, rows, E, V1));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int div,
                                   float* __restrict__ dst, int rows, int R) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), k = (int)(i % R);
    const int s = idx ? idx[r] : r / div;
    dst[i] = src[(size_t)s * R + k];
  }
}
int gather_rows(const float* src, const int32_t* idx, int div, float* dst, int rows, int R, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * R;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  RFN_CUDA(launch_pdl(gather_rows_kernel, dim3(blocks), dim3(256), 0, st, src, idx, div, dst, rows, R));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// out[r,k] = (((0 + in_0[r,k]) + in_1[r,k]) + ...) / n   -- Python sum() then "/ J"
// (misc/RecurrentFusionModel.py:307-309).  in_j[r,k] = in[j*stride + r*ld_in + k]
__global__ void mean_tensors_kernel(const float* __restrict__ in, size_t stride, int n, float* __restrict__ out,
                                    int ld_out, size_t count, int R, int ld_in) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / R, k = i % R;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += in[(size_t)j * stride + r * ld_in + k];
    out[r * ld_out + k] = s / (float)n;
  }
}
int mean_tensors(const float* in, size_t stride, int n, float* out, int ld_out, size_t count, int R, int ld_in,
                 cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (count == 0) return RFN_OK;
  const int blocks = (int)min((size_t)148 * 8, (count + 255) / 256);
  mean_tensors_kernel<<<blocks, 256, 0, st>>>(in, stride, n, out, ld_out, count, R, ld_in);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// out[r,k] = max_s in[r,s,k]   (torch.max(reason_mat, 1), misc/RecurrentFusionModel.py:303)
__global__ void max_over_steps_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int S, int K) {
  const size_t total = (size_t)rows * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, k = i % K;
    float m = in[(r * S) * K + k];
    for (int s = 1; s < S; ++s) m = fmaxf(m, in[(r * S + s) * K + k]);
    out[i] = m;
  }
}
int max_over_steps(const float* in, float* out, int rows, int S, int K, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * K;
  const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
  max_over_steps_kernel<<<blocks, 256, 0, st>>>(in, out, rows, S, K);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

__global__ void mean_logits_kernel(PtrList8 ptrs, int n, float* __restrict__ out, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int m = 0; m < n; ++m) s += ptrs.p[m][i];
    out[i] = s / (float)n;
  }
}
int mean_logits8(const PtrList8& ptrs, int n, float* out, size_t count, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (count == 0) return RFN_OK;
  const int blocks = (int)min((size_t)148 * 16, (count + 255) / 256);
  mean_logits_kernel<<<blocks, 256, 0, st>>>(ptrs, n, out, count);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // namespace rfn

extern "C" int rfn_lstm_cell_f32(const float* G, const float* c_prev, float* h_out, float* c_out, float* h_out2,
                                 int ldh2, int rows, int R, rfn_stream_t stream) {
  return rfn::lstm_cell(G, c_prev, h_out, c_out, h_out2, ldh2, nullptr, 0, rows, R, (cudaStream_t)stream);
}

extern "C" int rfn_lstm_cell_ex_f32(const float* G, const float* c_prev, const float* mask, float scale, int maxout, float* h_out,
                                    float* c_out, float* h_out2, int ldh2, float* h_out3, int ldh3, int rows, int R,
                                    rfn_stream_t stream) {
  return rfn::lstm_cell(G, c_prev, h_out, c_out, h_out2, ldh2, h_out3, ldh3, rows, R, (cudaStream_t)stream, mask, scale, maxout);
}

extern "C" int rfn_lstm_cell_drop_f32(const float* G, const float* c_prev, const float* mask, float scale, float* h_out,
                                      float* c_out, float* h_out2, int ldh2, float* h_out3, int ldh3, int rows, int R,
                                      rfn_stream_t stream) {
  return rfn::lstm_cell(G, c_prev, h_out, c_out, h_out2, ldh2, h_out3, ldh3, rows, R, (cudaStream_t)stream, mask, scale);
}

namespace rfn {
struct SumSrcs {
  const float* p[8];
  int ld[8];
  int n;
};
__global__ void sum_strided_kernel(SumSrcs src, float alpha, float* __restrict__ out, int ldo, int rows, int R) {
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), k = (int)(i % R);
    float s = 0.f;
    for (int j = 0; j < src.n; ++j) s += src.p[j][(size_t)r * src.ld[j] + k];
    out[(size_t)r * ldo + k] = alpha * s;
  }
}
}  // namespace rfn

extern "C" int rfn_sum_strided_f32(int n, const float* const* x, const int* ldx, float alpha, float* out, int ldo, int rows, int R,
                                   rfn_stream_t stream) {
  RFN_CHECK_ARG(n >= 1 && n <= 8 && x && ldx && out, "rfn_sum_strided_f32: 1..8 sources");
  if (rows == 0 || R == 0) return RFN_OK;
  rfn::SumSrcs src{};
  src.n = n;
  for (int j = 0; j < n; ++j) {
    RFN_CHECK_ARG(x[j] != nullptr, "rfn_sum_strided_f32: source %d is null", j);
    src.p[j] = x[j];
    src.ld[j] = ldx[j];
  }
  const size_t total = (size_t)rows * R;
  rfn::sum_strided_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, alpha, out, ldo, rows, R);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

extern "C" int rfn_mean_tensors_f32(const float* in, size_t stride, int n, float* out, int ld_out, int rows, int R, int ld_in,
                                    rfn_stream_t stream) {
  RFN_CHECK_ARG(in && out && n >= 1, "rfn_mean_tensors_f32: bad arguments");
  return rfn::mean_tensors(in, stride, n, out, ld_out, (size_t)rows * R, R, ld_in, (cudaStream_t)stream);
}
