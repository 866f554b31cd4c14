// fp32 SIMT GEMM: y[M,N] = sum_i x_i[M,K_i] . W_i[N,K_i]^T + bias.  Both operands are K-major
// (activations row-major, nn.Linear weights (out,in)), so tiles are read with 128-bit loads along K
// and transposed into shared memory.  This is the "warp-level FMA" engine the path uses whenever
// the row count is too small to be a dense tensor-core contraction, and the exact-fp32 reference
// the tcgen05 engine is validated against.
#include "rfn_internal.cuh"

namespace rfn {

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_tn_simt_kernel(GemmArgs a) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_F4 = BM * BK / 4;
  constexpr int B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT;
  constexpr int B_PER = (B_F4 + NT - 1) / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int M = a.M, N = a.N;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_PER], rb[B_PER];
  for (int s = 0; s < a.nsrc; ++s) {
    const float* __restrict__ x = a.src[s].x;
    const float* __restrict__ w = a.src[s].w;
    const int ldx = a.src[s].ldx, ldw = a.src[s].ldw, K = a.src[s].K;

    auto gload = [&](int k0) {
#pragma unroll
      for (int p = 0; p < A_PER; ++p) {
        const int i = tid + p * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < A_F4) {
          const int row = i >> 2, k = k0 + (i & 3) * 4;
          if (m0 + row < M && k < K) v = *reinterpret_cast<const float4*>(x + (size_t)(m0 + row) * ldx + k);
        }
        ra[p] = v;
      }
#pragma unroll
      for (int p = 0; p < B_PER; ++p) {
        const int i = tid + p * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < B_F4) {
          const int row = i >> 2, k = k0 + (i & 3) * 4;
          if (n0 + row < N && k < K) v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(n0 + row) * ldw + k));
        }
        rb[p] = v;
      }
    };
    auto sstore = [&]() {
#pragma unroll
      for (int p = 0; p < A_PER; ++p) {
        const int i = tid + p * NT;
        if (i < A_F4) {
          const int row = i >> 2, kq = (i & 3) * 4;
          As[kq + 0][row] = ra[p].x; As[kq + 1][row] = ra[p].y;
          As[kq + 2][row] = ra[p].z; As[kq + 3][row] = ra[p].w;
        }
      }
#pragma unroll
      for (int p = 0; p < B_PER; ++p) {
        const int i = tid + p * NT;
        if (i < B_F4) {
          const int row = i >> 2, kq = (i & 3) * 4;
          Bs[kq + 0][row] = rb[p].x; Bs[kq + 1][row] = rb[p].y;
          Bs[kq + 2][row] = rb[p].z; Bs[kq + 3][row] = rb[p].w;
        }
      }
    };

    gload(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
      sstore();
      __syncthreads();
      if (k0 + BK < K) gload(k0 + BK);  // prefetch the next tile into registers
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: bias (summed over sources), optional accumulate
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + tx * TN + j;
    if (n >= N) continue;
    float b = 0.f;
    for (int s = 0; s < a.nsrc; ++s)
      if (a.src[s].bias) b += __ldg(a.src[s].bias + n);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= M) continue;
      float* yp = a.y + (size_t)m * a.ldy + n;
      float v = acc[i][j] + b;
      if (a.accumulate) v += *yp;
      *yp = v;
    }
  }
}


// ---- skinny GEMM: M <= 32 rows ("warp-level FMA when batch x beam is small") --------------------------------
// The problem is a weight stream: every W element is used M times.  One warp per output column, lanes stride
// K with 128-bit coalesced loads of the W row, x staged in shared memory in K chunks, M accumulators per lane,
// deterministic shuffle reduction.  grid = N / 8 CTAs so that even N = 512 launches 64 CTAs, and K is consumed
// at several hundred GB/s per SM instead of one 16-wide tile per __syncthreads.
constexpr int SK_WARPS = 8;
constexpr int SK_KCHUNK = 512;   // floats of K staged per round: M * 512 * 4 bytes <= 64 KB

template <int MMAX>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(GemmArgs a) {
  extern __shared__ __align__(16) float s_x[];   // [MMAX][SK_KCHUNK]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * SK_WARPS + warp;
  const int M = a.M;
  float acc[MMAX];
#pragma unroll
  for (int m = 0; m < MMAX; ++m) acc[m] = 0.f;
  for (int s = 0; s < a.nsrc; ++s) {
    const float* __restrict__ x = a.src[s].x;
    const float* __restrict__ w = a.src[s].w + (size_t)min(n, a.N - 1) * a.src[s].ldw;
    const int K = a.src[s].K, ldx = a.src[s].ldx;
    for (int k0 = 0; k0 < K; k0 += SK_KCHUNK) {
      const int kc = min(SK_KCHUNK, K - k0);
      __syncthreads();
      for (int i = threadIdx.x * 4; i < MMAX * kc; i += SK_WARPS * 32 * 4) {
        const int m = i / kc, k = i % kc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M) v = *reinterpret_cast<const float4*>(x + (size_t)m * ldx + k0 + k);
        *reinterpret_cast<float4*>(s_x + m * SK_KCHUNK + k) = v;
      }
      __syncthreads();
      for (int k = lane * 4; k < kc; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k0 + k));
#pragma unroll
        for (int m = 0; m < MMAX; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(s_x + m * SK_KCHUNK + k);
          acc[m] = fmaf(wv.x, xv.x, acc[m]);
          acc[m] = fmaf(wv.y, xv.y, acc[m]);
          acc[m] = fmaf(wv.z, xv.z, acc[m]);
          acc[m] = fmaf(wv.w, xv.w, acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MMAX; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (n < a.N && lane == 0) {
    float b = 0.f;
    for (int s = 0; s < a.nsrc; ++s)
      if (a.src[s].bias) b += __ldg(a.src[s].bias + n);
#pragma unroll
    for (int m = 0; m < MMAX; ++m) {
      if (m < M) {
        float* yp = a.y + (size_t)m * a.ldy + n;
        float v = acc[m] + b;
        if (a.accumulate) v += *yp;
        *yp = v;
      }
    }
  }
}

template <int MMAX>
static int launch_skinny(const GemmArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)MMAX * SK_KCHUNK * sizeof(float);
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    RFN_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<MMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  gemm_skinny_kernel<MMAX><<<(a.N + SK_WARPS - 1) / SK_WARPS, SK_WARPS * 32, smem, st>>>(a);
  RFN_LAUNCH_CHECK();
  count_engine(ENG_SIMT_SKINNY);
  return RFN_OK;
}

int gemm_simt(const GemmArgs& a, cudaStream_t st) {
  ProfScope prof__(TAG_GEMM_OTHER, st);
  RFN_CHECK_ARG(a.nsrc >= 1 && a.nsrc <= 3, "gemm: n_src %d not in 1..3", a.nsrc);
  RFN_CHECK_ARG(a.M >= 0 && a.N > 0 && a.y != nullptr, "gemm: bad M/N/y");
  if (a.M == 0) return RFN_OK;
  for (int s = 0; s < a.nsrc; ++s) {
    const GemmSrc& g = a.src[s];
    RFN_CHECK_ARG(g.x && g.w, "gemm: null operand %d", s);
    RFN_CHECK_ARG(g.K > 0 && g.K % 4 == 0 && g.ldx % 4 == 0 && g.ldw % 4 == 0,
                  "gemm: K=%d ldx=%d ldw=%d must be multiples of 4", g.K, g.ldx, g.ldw);
    RFN_CHECK_ARG(((uintptr_t)g.x % 16 == 0) && ((uintptr_t)g.w % 16 == 0), "gemm: operands must be 16-byte aligned");
  }
  if (a.M <= 8) return launch_skinny<8>(a, st);
  if (a.M <= 16) return launch_skinny<16>(a, st);
  if (a.M <= 32) return launch_skinny<32>(a, st);
  if (a.M < 128) {
    // still a weight stream: 32 rows at a time (W comes from L2 after the first group)
    for (int m0 = 0; m0 < a.M; m0 += 32) {
      GemmArgs g = a;
      g.M = (a.M - m0 < 32) ? a.M - m0 : 32;
      g.y = a.y + (size_t)m0 * a.ldy;
      for (int s = 0; s < a.nsrc; ++s) g.src[s].x = a.src[s].x + (size_t)m0 * a.src[s].ldx;
      RFN_TRY(g.M <= 8 ? launch_skinny<8>(g, st) : (g.M <= 16 ? launch_skinny<16>(g, st) : launch_skinny<32>(g, st)));
    }
    return RFN_OK;
  }
  count_engine(ENG_SIMT_TILED);
  if (a.M <= 64) {
    dim3 grid((a.N + 63) / 64, (a.M + 63) / 64);
    gemm_tn_simt_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((a.N + 63) / 64, (a.M + 127) / 128);
    gemm_tn_simt_kernel<128, 64, 8, 4><<<grid, 256, 0, st>>>(a);
  }
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // namespace rfn

extern "C" int rfn_linear_f32(int n_src, const float* const* x, const int* ldx, const float* const* W,
                              const int* K, const float* const* bias, float* y, int ldy, int M, int N,
                              int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(n_src >= 1 && n_src <= 3 && x && ldx && W && K, "rfn_linear_f32: bad source arrays");
  rfn::GemmArgs a{};
  a.nsrc = n_src;
  for (int s = 0; s < n_src; ++s) a.src[s] = rfn::GemmSrc{x[s], W[s], bias ? bias[s] : nullptr, ldx[s], K[s], K[s]};
  a.y = y;
  a.ldy = ldy;
  a.M = M;
  a.N = N;
  a.accumulate = accumulate & 1; a.splitk_ok = (accumulate & RFN_GEMM_SPLITK) ? 1 : 0;
  return rfn::gemm(a, (cudaStream_t)stream);
}
