// Fused additive soft attention step (misc/AttentionModelCore.py:36-47):
//   e[n] = w . tanh(P[n,:] + g) + wb ;  a = softmax_n(e) ;  z = sum_n a[n] A[n,:]
// One CTA per (query row, D-slice).  P = att_2_att_h(A) comes from the GEMM engine; the feature map
// A is streamed exactly once per CTA slice with 128-bit coalesced loads; reductions are
// warp-shuffle based; no intermediate (rows,N,Ah) tensor is ever written (the reference
// materialises three).
#include <cuda_bf16.h>

#include <atomic>

#include "rfn_internal.cuh"

namespace rfn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

constexpr int ATT_THREADS = 256;
static std::atomic<int> g_att_bf16_variant{2};

// A16 != nullptr (engine mode 5): the feature map is read as the bf16 copy the GEMM engine already made (row pitch lda16
// elements), half the bytes of the fp32 map; everything else (scores, softmax, accumulation) stays fp32
template <bool BF16, int WIDE = 0>
__global__ void __launch_bounds__(ATT_THREADS)
attention_step_kernel(const float* __restrict__ A, const float* __restrict__ P,
                      const float* __restrict__ g, const float* __restrict__ w,
                      const float* __restrict__ d_wb, float* __restrict__ z, int ldz,
                      float* __restrict__ alpha, int N, int D, int Ah, int div,
                      const float* __restrict__ scores, int nslices, size_t slice_stride,
                      const __nv_bfloat16* __restrict__ A16, int lda16) {
  extern __shared__ __align__(16) float smem[];
  float* s_g = smem;            // Ah
  float* s_w = smem + Ah;       // Ah
  float* s_e = smem + 2 * Ah;   // N
  __shared__ float s_red[ATT_THREADS / 32];
  __shared__ float s_bcast;

  const int r = blockIdx.x;
  const int ra = r / div;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = ATT_THREADS / 32;
  pdl_trigger();
  pdl_wait();

  const float wb = __ldg(d_wb);
  if (scores) {
    // scores were reduced by the tcgen05 GEMM epilogue; only the bias is missing
    for (int n = tid; n < N; n += ATT_THREADS) {
      float e = 0.f;
      for (int sl = 0; sl < nslices; ++sl) e += scores[(size_t)sl * slice_stride + (size_t)r * N + n];  // fixed order
      s_e[n] = e + wb;
    }
  } else {
    for (int k = tid; k < Ah; k += ATT_THREADS) {
      s_g[k] = g[(size_t)r * Ah + k];
      s_w[k] = __ldg(w + k);
    }
    __syncthreads();
  }

  // ---- scores: one warp per attention location ---------------------------------------------
  const float* Pr = P + (size_t)ra * N * Ah;
  for (int n = warp; n < N && !scores; n += NW) {
    const float* pn = Pr + (size_t)n * Ah;
    float acc = 0.f;
    for (int k = lane * 4; k < Ah; k += 128) {
      const float4 p = *reinterpret_cast<const float4*>(pn + k);
      const float4 gg = *reinterpret_cast<const float4*>(s_g + k);
      const float4 ww = *reinterpret_cast<const float4*>(s_w + k);
      acc = fmaf(ww.x, tanhf(p.x + gg.x), acc);
      acc = fmaf(ww.y, tanhf(p.y + gg.y), acc);
      acc = fmaf(ww.z, tanhf(p.z + gg.z), acc);
      acc = fmaf(ww.w, tanhf(p.w + gg.w), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_e[n] = acc + wb;
  }
  __syncthreads();

  // ---- softmax over the N locations (no mask: SURVEY D3) ------------------------------------
  float m = -INFINITY;
  for (int n = tid; n < N; n += ATT_THREADS) m = fmaxf(m, s_e[n]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < NW; ++i) v = fmaxf(v, s_red[i]);
    s_bcast = v;
  }
  __syncthreads();
  m = s_bcast;
  float sum = 0.f;
  for (int n = tid; n < N; n += ATT_THREADS) {
    const float ex = expf(s_e[n] - m);
    s_e[n] = ex;
    sum += ex;
  }
  sum = warp_sum(sum);
  __syncthreads();  // everyone has read s_bcast
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < NW; ++i) v += s_red[i];
    s_bcast = v;
  }
  __syncthreads();
  const float total = s_bcast;
  for (int n = tid; n < N; n += ATT_THREADS) {
    const float an = s_e[n] / total;
    s_e[n] = an;
    if (alpha && blockIdx.y == 0) alpha[(size_t)r * N + n] = an;
  }
  __syncthreads();

  // ---- context: z[d] = sum_n a[n] A[n,d], this CTA's slice of D ------------------------------
  const int nvec = D >> 2;
  const int per = (nvec + gridDim.y - 1) / gridDim.y;
  const int v0 = blockIdx.y * per;
  const int v1 = min(nvec, v0 + per);
  if (BF16 && WIDE && ((D | lda16) & 7) == 0) {
    // 16-byte loads (8 bf16 per thread and location), four in flight per thread (64 bytes, what the fp32 path keeps in flight), the
    // row pointer advanced by adds, four softmax weights per LDS.128: 1.1 instructions per feature byte against 1.6 for the
    // 8-byte form below -- at the power-capped clocks of a long decode (1.3 GHz) the 8-byte form was bound by instruction issue,
    // not by HBM (ncu: 32 % issue-active at 56 % of the DRAM peak)
    const int nvec8 = D >> 3;
    const int per8 = (nvec8 + gridDim.y - 1) / gridDim.y;
    const int w0 = blockIdx.y * per8;
    const int w1 = min(nvec8, w0 + per8);
    const size_t rowb = (size_t)lda16 * 2;
    const char* base = reinterpret_cast<const char*>(A16 + (size_t)ra * N * lda16);
    for (int v = w0 + tid; v < w1; v += ATT_THREADS) {
      const char* p = base + (size_t)v * 16;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      auto fma8 = [&](float a0, const uint4& q) {
        acc[0] = fmaf(a0, __uint_as_float(q.x << 16), acc[0]);
        acc[1] = fmaf(a0, __uint_as_float(q.x & 0xffff0000u), acc[1]);
        acc[2] = fmaf(a0, __uint_as_float(q.y << 16), acc[2]);
        acc[3] = fmaf(a0, __uint_as_float(q.y & 0xffff0000u), acc[3]);
        acc[4] = fmaf(a0, __uint_as_float(q.z << 16), acc[4]);
        acc[5] = fmaf(a0, __uint_as_float(q.z & 0xffff0000u), acc[5]);
        acc[6] = fmaf(a0, __uint_as_float(q.w << 16), acc[6]);
        acc[7] = fmaf(a0, __uint_as_float(q.w & 0xffff0000u), acc[7]);
      };
      int n = 0;
      if (WIDE == 2) {   // eight 16-byte loads in flight: short feature maps (N = 49 / 64) are bound by the latency of their few batches
        for (; n + 8 <= N; n += 8) {
          uint4 q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) q[i] = __ldg(reinterpret_cast<const uint4*>(p + i * rowb));
          p += 8 * rowb;
          const float4 a4 = *reinterpret_cast<const float4*>(s_e + n);
          const float4 b4 = *reinterpret_cast<const float4*>(s_e + n + 4);
          fma8(a4.x, q[0]); fma8(a4.y, q[1]); fma8(a4.z, q[2]); fma8(a4.w, q[3]);
          fma8(b4.x, q[4]); fma8(b4.y, q[5]); fma8(b4.z, q[6]); fma8(b4.w, q[7]);
        }
      }
      for (; n + 4 <= N; n += 4) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(p + rowb));
        const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(p + 2 * rowb));
        const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(p + 3 * rowb));
        p += 4 * rowb;
        const float4 a4 = *reinterpret_cast<const float4*>(s_e + n);
        fma8(a4.x, q0); fma8(a4.y, q1); fma8(a4.z, q2); fma8(a4.w, q3);
      }
      for (; n < N; ++n, p += rowb) fma8(s_e[n], __ldg(reinterpret_cast<const uint4*>(p)));
      float4* zo = reinterpret_cast<float4*>(z + (size_t)r * ldz + v * 8);
      zo[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      zo[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    return;
  }
  if (BF16) {
    const uint2* Ar16 = reinterpret_cast<const uint2*>(A16 + (size_t)ra * N * lda16);
    const int pitch = lda16 >> 2;   // uint2 (4 bf16) per feature row
    for (int v = v0 + tid; v < v1; v += ATT_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // 8 loads of 8 bytes issued before their first use = the 64 bytes per thread the fp32 path keeps in flight (4 x 16);
      // a plain `#pragma unroll 8` is scheduled by ptxas as two groups of 4 loads (checked in the SASS)
      int n = 0;
      for (; n + 8 <= N; n += 8) {
        uint2 q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = __ldg(Ar16 + (size_t)(n + i) * pitch + v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a0 = s_e[n + i];
          acc.x = fmaf(a0, __uint_as_float(q[i].x << 16), acc.x);
          acc.y = fmaf(a0, __uint_as_float(q[i].x & 0xffff0000u), acc.y);
          acc.z = fmaf(a0, __uint_as_float(q[i].y << 16), acc.z);
          acc.w = fmaf(a0, __uint_as_float(q[i].y & 0xffff0000u), acc.w);
        }
      }
      for (; n < N; ++n) {
        const uint2 q = __ldg(Ar16 + (size_t)n * pitch + v);
        const float a0 = s_e[n];
        acc.x = fmaf(a0, __uint_as_float(q.x << 16), acc.x);
        acc.y = fmaf(a0, __uint_as_float(q.x & 0xffff0000u), acc.y);
        acc.z = fmaf(a0, __uint_as_float(q.y << 16), acc.z);
        acc.w = fmaf(a0, __uint_as_float(q.y & 0xffff0000u), acc.w);
      }
      *reinterpret_cast<float4*>(z + (size_t)r * ldz + v * 4) = acc;
    }
    return;
  }
  const float4* Ar = reinterpret_cast<const float4*>(A + (size_t)ra * N * D);
  for (int v = v0 + tid; v < v1; v += ATT_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int n = 0;
    for (; n + 4 <= N; n += 4) {
      const float4 x0 = __ldg(Ar + (size_t)(n + 0) * nvec + v);
      const float4 x1 = __ldg(Ar + (size_t)(n + 1) * nvec + v);
      const float4 x2 = __ldg(Ar + (size_t)(n + 2) * nvec + v);
      const float4 x3 = __ldg(Ar + (size_t)(n + 3) * nvec + v);
      const float a0 = s_e[n], a1 = s_e[n + 1], a2 = s_e[n + 2], a3 = s_e[n + 3];
      acc.x = fmaf(a0, x0.x, acc.x); acc.y = fmaf(a0, x0.y, acc.y); acc.z = fmaf(a0, x0.z, acc.z); acc.w = fmaf(a0, x0.w, acc.w);
      acc.x = fmaf(a1, x1.x, acc.x); acc.y = fmaf(a1, x1.y, acc.y); acc.z = fmaf(a1, x1.z, acc.z); acc.w = fmaf(a1, x1.w, acc.w);
      acc.x = fmaf(a2, x2.x, acc.x); acc.y = fmaf(a2, x2.y, acc.y); acc.z = fmaf(a2, x2.z, acc.z); acc.w = fmaf(a2, x2.w, acc.w);
      acc.x = fmaf(a3, x3.x, acc.x); acc.y = fmaf(a3, x3.y, acc.y); acc.z = fmaf(a3, x3.z, acc.z); acc.w = fmaf(a3, x3.w, acc.w);
    }
    for (; n < N; ++n) {
      const float4 x0 = __ldg(Ar + (size_t)n * nvec + v);
      const float a0 = s_e[n];
      acc.x = fmaf(a0, x0.x, acc.x); acc.y = fmaf(a0, x0.y, acc.y); acc.z = fmaf(a0, x0.z, acc.z); acc.w = fmaf(a0, x0.w, acc.w);
    }
    *reinterpret_cast<float4*>(z + (size_t)r * ldz + v * 4) = acc;
  }
}

int attention_step(const float* A, const float* P, const float* g, const float* w, const float* d_wb,
                   float* z, int ldz, float* alpha, int rows, int N, int D, int Ah, int div,
                   cudaStream_t st) {
  ProfScope prof__(TAG_ATTN_SMALL, st);
  RFN_CHECK_ARG(A && P && g && w && d_wb && z, "attention_step: null pointer");
  RFN_CHECK_ARG(rows >= 0 && N > 0 && div >= 1, "attention_step: bad rows/N/div");
  RFN_CHECK_ARG(D % 4 == 0 && Ah % 4 == 0 && ldz % 4 == 0, "attention_step: D=%d Ah=%d ldz=%d must be multiples of 4", D, Ah, ldz);
  if (rows == 0) return RFN_OK;
  // few rows: split D over several CTAs so the feature map streams from more SMs
  int dsplit = 1;
  while (rows * dsplit < 296 && dsplit < 8 && (D / 4) / (dsplit * 2) >= 64) dsplit *= 2;
  const size_t smem = (size_t)(2 * Ah + N) * sizeof(float);
  RFN_CHECK_ARG(smem <= 200 * 1024, "attention_step: N=%d Ah=%d exceed shared memory", N, Ah);
  if (smem > 48 * 1024)
    RFN_CUDA(cudaFuncSetAttribute(attention_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RFN_CUDA(launch_pdl(attention_step_kernel<false>, dim3(rows, dsplit), dim3(ATT_THREADS), smem, st, A, P, g, w, d_wb, z, ldz, alpha, N, D, Ah,
                      div, (const float*)nullptr, 0, (size_t)0, (const __nv_bfloat16*)nullptr, 0));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

int attention_from_scores(const float* A, const float* scores, int nslices, const float* d_wb, float* z, int ldz,
                          float* alpha, int rows, int N, int D, int div, cudaStream_t st, const void* A_bf16, int lda_bf16) {
  ProfScope prof__(TAG_ATTN_SMALL, st);
  RFN_CHECK_ARG((A || A_bf16) && scores && d_wb && z, "attention_from_scores: null pointer");
  RFN_CHECK_ARG(!A_bf16 || (lda_bf16 % 4 == 0 && lda_bf16 >= D), "attention_from_scores: bf16 feature pitch %d", lda_bf16);
  RFN_CHECK_ARG(rows >= 0 && N > 0 && div >= 1 && D % 4 == 0 && ldz % 4 == 0, "attention_from_scores: bad shape");
  if (rows == 0) return RFN_OK;
  int dsplit = 1;
  while (rows * dsplit < 296 && dsplit < 8 && (D / 4) / (dsplit * 2) >= 64) dsplit *= 2;
  const size_t smem = (size_t)N * sizeof(float);
  RFN_CHECK_ARG(smem <= 200 * 1024, "attention_from_scores: N=%d exceeds shared memory", N);
  if (A_bf16) {
    // rfn_set_att_bf16_variant: 0 = 8-byte loads, 1 = 16-byte loads, 2 (default) = 16-byte loads in batches of eight locations and
    // no second pass over a handful of leftover columns (D = 2208 is 276 16-byte vectors: two CTAs of 138 instead of 256 + 20).
    // Measured inside the bench on one box (profiles/r2_attention_bf16_variants.txt): 4.2 / 5.2 / 6.2 TB/s
    const int wide = g_att_bf16_variant.load();
    auto kern = wide == 2 ? attention_step_kernel<true, 2> : wide == 1 ? attention_step_kernel<true, 1> : attention_step_kernel<true, 0>;
    if (wide == 2 && D % 8 == 0 && (D / 8) / dsplit > ATT_THREADS && (D / 8) / dsplit < 2 * ATT_THREADS) dsplit *= 2;
    if (smem > 48 * 1024) RFN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RFN_CUDA(launch_pdl(kern, dim3(rows, dsplit), dim3(ATT_THREADS), smem, st, A, (const float*)nullptr,
                        (const float*)nullptr, (const float*)nullptr, d_wb, z, ldz, alpha, N, D, 0, div, scores, nslices,
                        (size_t)rows * N, (const __nv_bfloat16*)A_bf16, lda_bf16));
  } else {
    if (smem > 48 * 1024)
      RFN_CUDA(cudaFuncSetAttribute(attention_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RFN_CUDA(launch_pdl(attention_step_kernel<false>, dim3(rows, dsplit), dim3(ATT_THREADS), smem, st, A, (const float*)nullptr,
                        (const float*)nullptr, (const float*)nullptr, d_wb, z, ldz, alpha, N, D, 0, div, scores, nslices,
                        (size_t)rows * N, (const __nv_bfloat16*)nullptr, 0));
  }
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // namespace rfn

extern "C" int rfn_set_att_bf16_variant(int variant) {
  RFN_CHECK_ARG(variant >= 0 && variant <= 2, "rfn_set_att_bf16_variant: 0, 1 or 2");
  rfn::g_att_bf16_variant.store(variant);
  return RFN_OK;
}
extern "C" int rfn_get_att_bf16_variant(void) { return rfn::g_att_bf16_variant.load(); }

extern "C" int rfn_attention_step_f32(const float* A, const float* P, const float* g, const float* w,
                                      const float* d_wb, float* z, int ldz, float* alpha, int rows,
                                      int N, int D, int Ah, int div, rfn_stream_t stream) {
  return rfn::attention_step(A, P, g, w, d_wb, z, ldz, alpha, rows, N, D, Ah, div, (cudaStream_t)stream);
}

extern "C" int rfn_attention_core_f32(const float* h, const float* A, const float* U_w, const float* U_b,
                                      const float* Wh_w, const float* Wh_b, const float* v_w,
                                      const float* d_v_b, float* z, float* alpha, int rows, int N,
                                      int D, int R, int Ah, void* workspace, size_t workspace_bytes,
                                      rfn_stream_t stream) {
  const size_t need = ((size_t)rows * N * Ah + (size_t)rows * Ah) * sizeof(float);
  if (workspace_bytes < need || !workspace) {
    rfn::set_error("rfn_attention_core_f32: workspace %zu < %zu bytes", workspace_bytes, need);
    return RFN_ERR_WORKSPACE;
  }
  float* P = (float*)workspace;
  float* g = P + (size_t)rows * N * Ah;
  cudaStream_t st = (cudaStream_t)stream;
  RFN_TRY(rfn::gemm(rfn::gemm1(A, D, U_w, U_b, D, P, Ah, rows * N, Ah), st));
  RFN_TRY(rfn::gemm(rfn::gemm1(h, R, Wh_w, Wh_b, R, g, Ah, rows, Ah), st));
  return rfn::attention_step(A, P, g, v_w, d_v_b, z, D, alpha, rows, N, D, Ah, 1, st);
}
