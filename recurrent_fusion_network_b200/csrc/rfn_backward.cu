// Backward kernels of the path (XE / RL training): general-layout SIMT
// GEMMs for dX = dY.W and dW += dY^T.X, the additive-attention step backward, the LSTM cell backward,
// log-softmax / embedding / max-over-steps backward and the criteria gradients.
#include "rfn_internal.cuh"
#include "rfn_vocab.cuh"

namespace rfn {

// ---- general-layout fp32 GEMM: C[M,N] (+)= op(A)[M,K] . op(B)[K,N] -------------------------------------
//   A_KMAJOR: A stored (M, K) row-major (element (m,k) at m*lda + k); else stored (K, M) (at k*lda + m)
//   B_KMAJOR: B stored (N, K) row-major (element (k,n) at n*ldb + k); else stored (K, N) (at k*ldb + n)
struct GemmGenArgs {
  const float* A;
  const float* B;
  float* C;
  int lda, ldb, ldc;
  int M, N, K;
  int accumulate;
  int ksplit;   // > 1: blockIdx.z owns a K slice and adds its partial sums with atomics (C pre-zeroed / accumulated)
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_gen_kernel(GemmGenArgs a) {
  constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int kper = ((a.K + a.ksplit - 1) / a.ksplit + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kper, kend = min(a.K, kbeg + kper);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // each thread stages 4 elements of A and 4 of B (64x16 tiles, 256 threads)
    {
      if (A_KMAJOR) {
        const int row = tid >> 2, kq = (tid & 3) * 4;
        const int m = m0 + row;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + kq + e;
          As[kq + e][row] = (m < a.M && k < kend) ? a.A[(size_t)m * a.lda + k] : 0.f;
        }
      } else {
        const int kk = tid >> 4, mq = (tid & 15) * 4;
        const int k = k0 + kk;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int m = m0 + mq + e;
          As[kk][mq + e] = (m < a.M && k < kend) ? a.A[(size_t)k * a.lda + m] : 0.f;
        }
      }
      if (B_KMAJOR) {
        const int row = tid >> 2, kq = (tid & 3) * 4;
        const int n = n0 + row;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = k0 + kq + e;
          Bs[kq + e][row] = (n < a.N && k < kend) ? __ldg(a.B + (size_t)n * a.ldb + k) : 0.f;
        }
      } else {
        const int kk = tid >> 4, nq = (tid & 15) * 4;
        const int k = k0 + kk;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = n0 + nq + e;
          Bs[kk][nq + e] = (n < a.N && k < kend) ? __ldg(a.B + (size_t)k * a.ldb + n) : 0.f;
        }
      }
    }
    __syncthreads();
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
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= a.N) continue;
      float* cp = a.C + (size_t)m * a.ldc + n;
      if (a.ksplit > 1) atomicAdd(cp, acc[i][j]);
      else *cp = a.accumulate ? (*cp + acc[i][j]) : acc[i][j];
    }
  }
}

// dst[c][r] = src[r][c], zero padding up to ld_dst
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int ld_src, int rows, int cols,
                                                        float* __restrict__ dst, int ld_dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < ld_dst) dst[(size_t)c * ld_dst + r] = tile[tx][i];
  }
}

static bool tc_general_ok(const float* A, int lda, const float* B, int ldb, const float* C, int ldc, int N, int K) {
  return N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && ((uintptr_t)A % 16) == 0 &&
         ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0;
}
static int gemm_general_tc(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                           int accumulate, int passes, cudaStream_t st) {
  GemmArgs g{};
  g.src[0] = GemmSrc{A, B, nullptr, lda, ldb, K};
  g.nsrc = 1;
  g.y = C; g.ldy = ldc; g.M = M; g.N = N; g.accumulate = accumulate ? 1 : 0; g.splitk_ok = 1;
  return gemm_tc_splitk(g, true, passes, st);
}

static int gemm_general_simt(bool a_kmajor, bool b_kmajor, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                             int M, int N, int K, int accumulate, cudaStream_t st);

int gemm_general(bool a_kmajor, bool b_kmajor, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M,
                 int N, int K, int accumulate, cudaStream_t st) {
  ProfScope prof__(TAG_GEMM_BWD, st);
  RFN_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0, "gemm_general: bad arguments");
  // dX = dY . W with W = nn.Linear weights (out, in) = (K, N) row-major: tensor engine, B operand MN-major
  if (a_kmajor && !b_kmajor && gemm_mode() >= 1 && (long)M * N * K >= (1L << 25) && N >= 128 &&
      tc_general_ok(A, lda, B, ldb, C, ldc, N, K))
    return gemm_general_tc(A, lda, B, ldb, C, ldc, M, N, K, accumulate, tc_passes(gemm_mode()), st);
  return gemm_general_simt(a_kmajor, b_kmajor, A, lda, B, ldb, C, ldc, M, N, K, accumulate, st);
}

static int gemm_general_simt(bool a_kmajor, bool b_kmajor, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                             int M, int N, int K, int accumulate, cudaStream_t st) {
  if (M == 0 || N == 0) return RFN_OK;
  // few output tiles and a long contraction (dX = dY . W at small batch): split K over blockIdx.z
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  int ksplit = 1;
  while (tiles * ksplit < 296 && K / (ksplit * 2) >= 64 && ksplit < 32) ksplit *= 2;
  if (ksplit > 1 && !accumulate) RFN_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st));
  GemmGenArgs a{A, B, C, lda, ldb, ldc, M, N, K, accumulate, ksplit};
  dim3 grid((N + 63) / 64, (M + 63) / 64, ksplit);
  if (a_kmajor && b_kmajor) gemm_gen_kernel<true, true><<<grid, 256, 0, st>>>(a);
  else if (a_kmajor && !b_kmajor) gemm_gen_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else if (!a_kmajor && b_kmajor) gemm_gen_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else gemm_gen_kernel<false, false><<<grid, 256, 0, st>>>(a);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// db[n] (+)= sum_m dY[m,n]: 32 columns x 8 row-lanes per block, each block reduces a 256-row slab and adds
// its partial with one atomic per column
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dY, int ld, int M, int N, float* __restrict__ db) {
  __shared__ float s_p[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * 256;
  float s = 0.f;
  if (n < N)
    for (int m = m0 + ty; m < min(M, m0 + 256); m += 8) s += dY[(size_t)m * ld + n];
  s_p[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_p[i][tx];
    atomicAdd(db + n, t);
  }
}
// the same for M <= 256 (one row slab per column block): plain store / read-add-store, no memset and no atomics, so
// the result does not depend on arrival order (every bias gradient of an 80-row training batch takes this route)
__global__ void __launch_bounds__(256) colsum_small_kernel(const float* __restrict__ dY, int ld, int M, int N, float* db,
                                                           int accumulate) {
  __shared__ float s_p[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 8) s += dY[(size_t)m * ld + n];
  s_p[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_p[i][tx];
    db[n] = accumulate ? db[n] + t : t;
  }
}
int colsum(const float* dY, int ld, int M, int N, float* db, int accumulate, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (N == 0) return RFN_OK;
  if (M > 0 && M <= 256) {
    colsum_small_kernel<<<(N + 31) / 32, 256, 0, st>>>(dY, ld, M, N, db, accumulate);
    RFN_LAUNCH_CHECK();
    return RFN_OK;
  }
  if (!accumulate) RFN_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
  if (M == 0) return RFN_OK;
  colsum_kernel<<<dim3((N + 31) / 32, (M + 255) / 256), 256, 0, st>>>(dY, ld, M, N, db);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- additive attention step backward (misc/AttentionModelCore.py:36-47) ------------------------------
// forward:  u = P[n,:] + g ; t = tanh(u) ; e[n] = w.t + wb ; a = softmax(e) ; z = sum_n a[n] A[n,:]
// given dz: da[n] = dz.A[n,:] ; de = a * (da - sum_m a[m] da[m]) ; dwb += sum de ; dw[k] += sum_n de[n] t[n,k]
//           du[n,k] = de[n] w[k] (1 - t^2) = dP[n,k] ; dg[k] = sum_n du[n,k] ; dA[n,:] += a[n] dz   (if requested)
// one CTA per query row; dw / dwb are accumulated with atomics (one per CTA per element).
constexpr int AB_THREADS = 512;
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_kernel(const float* __restrict__ A, const float* __restrict__ P, const float* __restrict__ g,
                     const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ dz, int lddz,
                     float* __restrict__ dP, float* __restrict__ dg, float* __restrict__ dw, float* __restrict__ dwb,
                     float* __restrict__ dA, int N, int D, int Ah, int div) {
  extern __shared__ __align__(16) float sm[];
  float* s_de = sm;          // N
  float* s_dg = sm + N;      // Ah
  float* s_dw = s_dg + Ah;   // Ah
  __shared__ float s_red[AB_THREADS / 32];
  __shared__ float s_b;
  const int r = blockIdx.x, ra = r / div;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = AB_THREADS / 32;
  const float* Ar = A + (size_t)ra * N * D;
  const float* dzr = dz + (size_t)r * lddz;
  // da[n] = dz . A[n,:]   (one warp per location)
  for (int n = warp; n < N; n += NW) {
    float s = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 x = *reinterpret_cast<const float4*>(Ar + (size_t)n * D + d);
      const float4 y = *reinterpret_cast<const float4*>(dzr + d);
      s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_de[n] = s;
  }
  for (int k = tid; k < Ah; k += AB_THREADS) { s_dg[k] = 0.f; s_dw[k] = 0.f; }
  __syncthreads();
  // dot = sum_m a[m] da[m]
  float part = 0.f;
  for (int n = tid; n < N; n += AB_THREADS) part += alpha[(size_t)r * N + n] * s_de[n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < NW; ++i) v += s_red[i];
    s_b = v;
  }
  __syncthreads();
  const float dot = s_b;
  float dwb_part = 0.f;
  for (int n = tid; n < N; n += AB_THREADS) {
    const float de = alpha[(size_t)r * N + n] * (s_de[n] - dot);
    s_de[n] = de;
    dwb_part += de;
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dwb_part += __shfl_xor_sync(0xffffffffu, dwb_part, o);
  if (lane == 0) s_red[warp] = dwb_part;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < NW; ++i) v += s_red[i];
    atomicAdd(dwb, v);
  }
  // du / dP, dg, dw: thread k-strided over Ah, loop over n
  const float* Pr = P + (size_t)ra * N * Ah;
  float* dPr = dP + (size_t)r * N * Ah;
  for (int k = tid; k < Ah; k += AB_THREADS) {
    const float gk = g[(size_t)r * Ah + k], wk = __ldg(w + k);
    float dgk = 0.f, dwk = 0.f;
    for (int n = 0; n < N; ++n) {
      const float t = tanhf(Pr[(size_t)n * Ah + k] + gk);
      const float de = s_de[n];
      const float du = de * wk * (1.f - t * t);
      dPr[(size_t)n * Ah + k] = du;
      dgk += du;
      dwk += de * t;
    }
    dg[(size_t)r * Ah + k] = dgk;
    atomicAdd(dw + k, dwk);
  }
  // dA[n,:] += a[n] dz   (thought vectors need gradients; CNN features do not)
  if (dA) {
    float* dAr = dA + (size_t)ra * N * D;
    for (int i = tid; i < N * D; i += AB_THREADS) {
      const int n = i / D, d = i % D;
      const float v = alpha[(size_t)r * N + n] * dzr[d];
      if (div == 1) dAr[i] += v; else atomicAdd(dAr + i, v);
    }
  }
}

int attention_bwd(const float* A, const float* P, const float* g, const float* w, const float* alpha, const float* dz,
                  int lddz, float* dP, float* dg, float* dw, float* dwb, float* dA, int rows, int N, int D, int Ah, int div,
                  cudaStream_t st) {
  ProfScope prof__(TAG_ATTN_BWD, st);
  RFN_CHECK_ARG(A && P && g && w && alpha && dz && dP && dg && dw && dwb, "attention_bwd: null pointer");
  RFN_CHECK_ARG(D % 4 == 0 && lddz % 4 == 0, "attention_bwd: D and lddz must be multiples of 4");
  if (rows == 0) return RFN_OK;
  const size_t smem = (size_t)(N + 2 * Ah) * sizeof(float);
  RFN_CHECK_ARG(smem <= 200 * 1024, "attention_bwd: N/Ah too large");
  if (smem > 48 * 1024)
    RFN_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_bwd_kernel<<<rows, AB_THREADS, smem, st>>>(A, P, g, w, alpha, dz, lddz, dP, dg, dw, dwb, dA, N, D, Ah, div);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- LSTM cell backward (gate order [i|f|o|g]) ---------------------------------------------------------
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ G, const float* __restrict__ c_prev,
                                     const float* __restrict__ dh, const float* __restrict__ dc_next,
                                     float* __restrict__ dG, float* __restrict__ dc_prev, int rows, int R) {
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), k = (int)(i % R);
    const float* Gr = G + (size_t)r * 4 * R;
    const float ig = sigm(Gr[k]), fg = sigm(Gr[R + k]), og = sigm(Gr[2 * R + k]), gg = tanhf(Gr[3 * R + k]);
    const float cp = c_prev[i];
    const float c2 = fg * cp + ig * gg;
    const float tc = tanhf(c2);
    const float dhv = dh ? dh[i] : 0.f;
    const float dc = (dc_next ? dc_next[i] : 0.f) + dhv * og * (1.f - tc * tc);
    float* dGr = dG + (size_t)r * 4 * R;
    dGr[k] = dc * gg * ig * (1.f - ig);
    dGr[R + k] = dc * cp * fg * (1.f - fg);
    dGr[2 * R + k] = dhv * tc * og * (1.f - og);
    dGr[3 * R + k] = dc * ig * (1.f - gg * gg);
    dc_prev[i] = dc * fg;
  }
}
int lstm_cell_bwd(const float* G, const float* c_prev, const float* dh, const float* dc_next, float* dG, float* dc_prev,
                  int rows, int R, cudaStream_t st) {
  ProfScope prof__(TAG_CELL, st);
  RFN_CHECK_ARG(G && c_prev && dG && dc_prev, "lstm_cell_bwd: null pointer");
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * R;
  lstm_cell_bwd_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, st>>>(G, c_prev, dh, dc_next, dG, dc_prev, rows, R);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// The same with dh given as the SUM of up to 8 strided sources (the thought-vector slot's gradient, the slice of dH the
// next fusion step returned, the query gradient of the next step's attention, ...), an optional dropout keep-mask on h
// (h = mask * scale * o tanh(c), so dh_raw = mask * scale * dh) -- replaces the add / mul kernels between the steps of a
// hand-scheduled backward pass (recurrent_fusion_network_b200/tape.py).
struct CellBwdSrcs {
  const float* p[8];
  int ld[8];
  int n;
};
__global__ void lstm_cell_bwd_multi_kernel(const float* __restrict__ G, const float* __restrict__ c_prev, CellBwdSrcs src,
                                           const float* __restrict__ mask, float scale, const float* __restrict__ dc_next,
                                           float* __restrict__ dG, float* __restrict__ dc_prev, int rows, int R, int maxout) {
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / R), k = (int)(i % R);
    const float* Gr = G + (size_t)r * (4 + maxout) * R;
    const float ig = sigm(Gr[k]), fg = sigm(Gr[R + k]), og = sigm(Gr[2 * R + k]);
    const float g1 = Gr[3 * R + k], g2 = maxout ? Gr[4 * R + k] : 0.f;
    const float gg = maxout ? fmaxf(g1, g2) : tanhf(g1);
    const float cp = c_prev[i];
    const float c2 = fg * cp + ig * gg;
    const float tc = tanhf(c2);
    float dhv = 0.f;
    for (int s = 0; s < src.n; ++s) dhv += src.p[s][(size_t)r * src.ld[s] + k];
    if (mask) dhv *= mask[i] * scale;
    const float dc = (dc_next ? dc_next[i] : 0.f) + dhv * og * (1.f - tc * tc);
    float* dGr = dG + (size_t)r * (4 + maxout) * R;
    dGr[k] = dc * gg * ig * (1.f - ig);
    dGr[R + k] = dc * cp * fg * (1.f - fg);
    dGr[2 * R + k] = dhv * tc * og * (1.f - og);
    if (maxout) {   // d max(g1, g2): to the larger one; split evenly on a tie (torch.max(a, b)'s backward)
      const float dt = dc * ig;
      dGr[3 * R + k] = g1 > g2 ? dt : (g1 == g2 ? 0.5f * dt : 0.f);
      dGr[4 * R + k] = g2 > g1 ? dt : (g1 == g2 ? 0.5f * dt : 0.f);
    } else {
      dGr[3 * R + k] = dc * ig * (1.f - gg * gg);
    }
    dc_prev[i] = dc * fg;
  }
}
int lstm_cell_bwd_multi(const float* G, const float* c_prev, const CellBwdSrcs& src, const float* mask, float scale,
                        const float* dc_next, float* dG, float* dc_prev, int rows, int R, cudaStream_t st, int maxout = 0) {
  ProfScope prof__(TAG_CELL, st);
  RFN_CHECK_ARG(G && c_prev && dG && dc_prev, "lstm_cell_bwd_multi: null pointer");
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * R;
  lstm_cell_bwd_multi_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, st>>>(G, c_prev, src, mask, scale, dc_next,
                                                                                          dG, dc_prev, rows, R, maxout ? 1 : 0);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- log-softmax backward: dx = dlp - exp(lp) * sum_v dlp ---------------------------------------------
__global__ void __launch_bounds__(256)
log_softmax_bwd_kernel(const float* __restrict__ lp, size_t ld_lp, const float* __restrict__ dlp, size_t ld_d,
                       float* __restrict__ dx, size_t ld_x, int V) {
  __shared__ float s_red[8];
  __shared__ float s_b;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* l = lp + (size_t)r * ld_lp;
  const float* d = dlp + (size_t)r * ld_d;
  float s = 0.f;
  for (int v = tid; v < V; v += 256) s += d[v];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += s_red[i];
    s_b = v;
  }
  __syncthreads();
  const float tot = s_b;
  float* o = dx + (size_t)r * ld_x;
  for (int v = tid; v < V; v += 256) o[v] = d[v] - expf(l[v]) * tot;
}
int log_softmax_bwd(const float* lp, size_t ld_lp, const float* dlp, size_t ld_d, float* dx, size_t ld_x, int rows, int V,
                    cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  log_softmax_bwd_kernel<<<rows, 256, 0, st>>>(lp, ld_lp, dlp, ld_d, dx, ld_x, V);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- embedding backward: dE[tok[r], :] += dx[r, :] -----------------------------------------------------
__global__ void embed_bwd_kernel(const int64_t* __restrict__ tok, int ld_tok, const float* __restrict__ dx,
                                 float* __restrict__ dE, int rows, int E, int V1) {
  const size_t total = (size_t)rows * E;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / E), k = (int)(i % E);
    long long t = tok[(size_t)r * ld_tok];
    t = t < 0 ? 0 : (t >= V1 ? V1 - 1 : t);
    atomicAdd(dE + (size_t)t * E + k, dx[i]);
  }
}
int embed_bwd(const int64_t* tok, int ld_tok, const float* dx, float* dE, int rows, int E, int V1, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * E;
  embed_bwd_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, st>>>(tok, ld_tok, dx, dE, rows, E, V1);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- max over steps backward: the gradient goes to the first arg-max step (torch.max semantics) ----------
__global__ void max_over_steps_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                          float* __restrict__ din, int rows, int S, int K) {
  const size_t total = (size_t)rows * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, k = i % K;
    int best = 0;
    float m = in[(r * S) * K + k];
    for (int s = 1; s < S; ++s) {
      const float v = in[(r * S + s) * K + k];
      if (v > m) { m = v; best = s; }
    }
    for (int s = 0; s < S; ++s) din[(r * S + s) * K + k] = (s == best) ? dout[i] : 0.f;
  }
}
int max_over_steps_bwd(const float* in, const float* dout, float* din, int rows, int S, int K, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (rows == 0) return RFN_OK;
  const size_t total = (size_t)rows * K;
  max_over_steps_bwd_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, st>>>(in, dout, din, rows, S, K);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- out = alpha * x + beta * y (y may be null) ------------------------------------------------------------
__global__ void axpby_kernel(float alpha, const float* __restrict__ x, float beta, const float* y, float* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = alpha * x[i] + (y ? beta * y[i] : 0.f);
}
int axpby(float alpha, const float* x, float beta, const float* y, float* out, size_t n, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (n == 0) return RFN_OK;
  axpby_kernel<<<(int)min((size_t)148 * 8, (n + 255) / 256), 256, 0, st>>>(alpha, x, beta, y, out, n);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}


// ---- out = s * x * m  (dropout with an explicit keep-mask) ---------------------------------------------
__global__ void mul_scale_kernel(float sc, const float* __restrict__ x, const float* __restrict__ m, float* out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = sc * x[i] * m[i];
}
int mul_scale(float sc, const float* x, const float* m, float* out, size_t n, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  if (n == 0) return RFN_OK;
  mul_scale_kernel<<<(int)min((size_t)148 * 8, (n + 255) / 256), 256, 0, st>>>(sc, x, m, out, n);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- token selection from a log-prob row (misc/RecurrentFusionModel.py:619-635) --------------------------
// uniforms == NULL: arg-max, ties -> lower index (torch.max); else inverse CDF of exp(lp / temperature) in
// index order, accumulated in fp64, against one uniform per row.
__global__ void __launch_bounds__(256)
lp_select_kernel(const float* __restrict__ lp, size_t ld, int V, const float* __restrict__ uniforms, float temperature,
                 int64_t* __restrict__ tok, float* __restrict__ lp_out) {
  __shared__ double s_pref[257];
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_cnt[8];
  __shared__ int s_tok;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = lp + (size_t)r * ld;
  if (uniforms == nullptr) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int v = tid; v < V; v += 256) {
      const float xv = x[v];
      if (xv > bv || (xv == bv && v < bi)) { bv = xv; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < 8; ++i)
        if (s_v[i] > bv || (s_v[i] == bv && s_i[i] < bi)) { bv = s_v[i]; bi = s_i[i]; }
      tok[r] = bi;
      lp_out[r] = bv;
    }
    return;
  }
  const int seg = (V + 255) / 256;
  const int v0 = min(V, tid * seg), v1 = min(V, v0 + seg);
  double loc = 0.0;
  for (int v = v0; v < v1; ++v) loc += (double)((temperature == 1.0f) ? expf(x[v]) : expf(x[v] / temperature));
  s_pref[tid + 1] = loc;
  __syncthreads();
  if (tid == 0) {
    s_pref[0] = 0.0;
    for (int i = 1; i <= 256; ++i) s_pref[i] += s_pref[i - 1];
  }
  __syncthreads();
  const double thr = (double)uniforms[r] * s_pref[256];
  double run = s_pref[tid];
  int cnt = 0;
  for (int v = v0; v < v1; ++v) {
    run += (double)((temperature == 1.0f) ? expf(x[v]) : expf(x[v] / temperature));
    cnt += (run <= thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int i = 0; i < 8; ++i) c += s_cnt[i];
    c = min(c, V - 1);
    tok[r] = c;
    lp_out[r] = x[c];
  }
}
int lp_select(const float* lp, size_t ld, int rows, int V, const float* uniforms, float temperature, int64_t* tok,
              float* lp_out, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  lp_select_kernel<<<rows, 256, 0, st>>>(lp, ld, V, uniforms, temperature, tok, lp_out);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// out[r] = x[r, idx[r]]  and its backward  dx[r, idx[r]] = dout[r] (dx zero elsewhere)
__global__ void gather_cols_kernel(const float* __restrict__ x, size_t ld, const int64_t* __restrict__ idx, float* out, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) out[r] = x[(size_t)r * ld + idx[r]];
}
__global__ void scatter_cols_kernel(const float* __restrict__ dout, const int64_t* __restrict__ idx, float* dx, size_t ld, int V) {
  const int r = blockIdx.x;
  const int64_t t = idx[r];
  const float d = dout[r];
  for (int v = threadIdx.x; v < V; v += blockDim.x) dx[(size_t)r * ld + v] = (v == t) ? d : 0.f;
}


// ---- row replication (dataloader.py:251-252 repeats every image seq_per_img times) ---------------------
// out[r, :] = sum_{i < g} x[r*g + i, :]   -- backward of expanding each unique row to g identical rows
__global__ void group_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int g, int R) {
  const size_t total = (size_t)rows * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / R, k = i % R;
    float s = 0.f;
    for (int j = 0; j < g; ++j) s += x[(r * g + j) * R + k];
    out[i] = s;
  }
}

// ---- criteria gradients ------------------------------------------------------------------------------------
// XE (misc/utils.py:161-184): dL/dlp[b,t,v] = -gout * mask[b,t]/rows * ((1-eps) 1[v == y] + eps/V)
__global__ void __launch_bounds__(256)
xe_loss_bwd_kernel(const int64_t* __restrict__ target, const float* __restrict__ mask, int ld_t, int T, int V, float eps,
                   float inv_rows, const float* __restrict__ gout, float* __restrict__ dlp, size_t ld_b, size_t ld_s) {
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const float mk = mask[(size_t)b * ld_t + t] * inv_rows * gout[0];
  long long y = target[(size_t)b * ld_t + t];
  y = y < 0 ? 0 : (y >= V ? V - 1 : y);
  float* o = dlp + (size_t)b * ld_b + (size_t)t * ld_s;
  const float base = -mk * (eps / (float)V);
  for (int v = threadIdx.x; v < V; v += 256) o[v] = base + (v == (int)y ? -mk * (1.f - eps) : 0.f);
}
int xe_loss_bwd(const int64_t* target, const float* mask, int ld_t, int rows, int T, int V, float eps, const float* gout,
                float* dlp, size_t ld_b, size_t ld_s, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows * T == 0) return RFN_OK;
  xe_loss_bwd_kernel<<<rows * T, 256, 0, st>>>(target, mask, ld_t, T, V, eps, 1.f / (float)rows, gout, dlp, ld_b, ld_s);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// RL (misc/utils.py:50-72): d/d slp[b,t] = -gout * R * m / rows ;
//   d/d lp_all[b,t,v] = gout * entropy_reg/rows * m0 * p (1 + lp)   for t < T, 0 for the extra last step
__global__ void __launch_bounds__(256)
rl_loss_bwd_kernel(const int64_t* __restrict__ seq, const float* __restrict__ reward, const float* __restrict__ lp_all,
                   size_t ld_lp_rows, size_t ld_lp_s, int T, int T1, int V, float entropy_reg, float inv_rows,
                   const float* __restrict__ gout, float* __restrict__ dslp, float* __restrict__ dlp_all) {
  const int b = blockIdx.x / T1, t = blockIdx.x % T1;
  float* o = dlp_all + (size_t)b * ld_lp_rows + (size_t)t * ld_lp_s;
  if (t >= T) {
    for (int v = threadIdx.x; v < V; v += 256) o[v] = 0.f;
    return;
  }
  const bool m0 = seq[(size_t)b * T + t] > 0;
  const bool m = (t == 0) ? true : (seq[(size_t)b * T + t - 1] > 0);
  const float go = gout[0];
  if (threadIdx.x == 0) dslp[(size_t)b * T + t] = m ? -go * reward[(size_t)b * T + t] * inv_rows : 0.f;
  const float* x = lp_all + (size_t)b * ld_lp_rows + (size_t)t * ld_lp_s;
  const float c = m0 ? go * entropy_reg * inv_rows : 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float l = x[v];
    o[v] = c * expf(l) * (1.f + l);
  }
}
int rl_loss_bwd(const int64_t* seq, const float* reward, const float* lp_all, size_t ld_lp_rows, size_t ld_lp_s, int rows, int T,
                int T1, int V, float entropy_reg, const float* gout, float* dslp, float* dlp_all, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows * T1 == 0) return RFN_OK;
  rl_loss_bwd_kernel<<<rows * T1, 256, 0, st>>>(seq, reward, lp_all, ld_lp_rows, ld_lp_s, T, T1, V, entropy_reg, 1.f / (float)rows,
                                                gout, dslp, dlp_all);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// MultiLabelMarginLoss backward: for target j and non-target i with 1 - (x[y_j] - x[i]) > 0:
//   dx[y_j] -= s, dx[i] += s, s = gout * weight / (K * rows)
__global__ void __launch_bounds__(256)
multilabel_margin_bwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ target, int K, float scale,
                             const float* __restrict__ gout, float* __restrict__ dx) {
  extern __shared__ unsigned char s_is[];
  __shared__ int s_nt;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* xr = x + (size_t)r * K;
  const int64_t* tr = target + (size_t)r * K;
  float* dr = dx + (size_t)r * K;
  for (int i = tid; i < K; i += 256) { s_is[i] = 0; dr[i] = 0.f; }
  if (tid == 0) {
    int n = 0;
    while (n < K && tr[n] >= 0) ++n;
    s_nt = n;
  }
  __syncthreads();
  const int nt = s_nt;
  for (int j = tid; j < nt; j += 256) s_is[(int)tr[j]] = 1;
  __syncthreads();
  const float s = scale * gout[0];
  for (int i = tid; i < K; i += 256) {
    if (s_is[i]) continue;
    const float xi = xr[i];
    float mine = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int y = (int)tr[j];
      if (1.f - (xr[y] - xi) > 0.f) {
        mine += s;
        atomicAdd(dr + y, -s);
      }
    }
    dr[i] += mine;   // non-target slots are written by their owner thread only
  }
}
int multilabel_margin_bwd(const float* pred, const int64_t* target, int rows, int K, float weight, const float* gout,
                          float* dx, cudaStream_t st) {
  ProfScope prof__(TAG_VOCAB, st);
  if (rows == 0) return RFN_OK;
  multilabel_margin_bwd_kernel<<<rows, 256, (size_t)K, st>>>(pred, target, K, weight / ((float)K * (float)rows), gout, dx);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // namespace rfn

using namespace rfn;
extern "C" {

int rfn_transpose_f32(const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst, rfn_stream_t stream) {
  RFN_CHECK_ARG(src && dst && rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= rows, "rfn_transpose_f32: bad arguments");
  if (rows == 0 || cols == 0) return RFN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof__(TAG_GEMM_BWD, st);
  dim3 grid((ld_dst + 31) / 32, (cols + 31) / 32);
  transpose_kernel<<<grid, 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}
int rfn_gemm_general_f32_engine(int engine, int a_kmajor, int b_kmajor, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, int M, int N, int K, int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(engine >= 0 && engine <= 3, "rfn_gemm_general_f32_engine: engine %d not in {0,1,2,3}", engine);
  if (engine >= 1) {
    RFN_CHECK_ARG(a_kmajor && !b_kmajor && A && B && C && tc_general_ok(A, lda, B, ldb, C, ldc, N, K),
                  "rfn_gemm_general_f32_engine: the tensor engine takes the (1, 0) layout with 16-byte aligned rows");
    return gemm_general_tc(A, lda, B, ldb, C, ldc, M, N, K, accumulate, tc_passes(engine), (cudaStream_t)stream);
  }
  RFN_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0, "rfn_gemm_general_f32_engine: bad arguments");
  return gemm_general_simt(a_kmajor != 0, b_kmajor != 0, A, lda, B, ldb, C, ldc, M, N, K, accumulate, (cudaStream_t)stream);
}
int rfn_gemm_general_f32(int a_kmajor, int b_kmajor, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                         int M, int N, int K, int accumulate, rfn_stream_t stream) {
  return gemm_general(a_kmajor != 0, b_kmajor != 0, A, lda, B, ldb, C, ldc, M, N, K, accumulate, (cudaStream_t)stream);
}
int rfn_colsum_f32(const float* dY, int ld, int M, int N, float* db, int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(dY && db, "rfn_colsum_f32: null pointer");
  return colsum(dY, ld, M, N, db, accumulate, (cudaStream_t)stream);
}
int rfn_attention_step_bwd_f32(const float* A, const float* P, const float* g, const float* w, const float* alpha,
                               const float* dz, int lddz, float* dP, float* dg, float* dw, float* dwb, float* dA, int rows,
                               int N, int D, int Ah, int div, rfn_stream_t stream) {
  return attention_bwd(A, P, g, w, alpha, dz, lddz, dP, dg, dw, dwb, dA, rows, N, D, Ah, div, (cudaStream_t)stream);
}
int rfn_lstm_cell_bwd_f32(const float* G, const float* c_prev, const float* dh, const float* dc_next, float* dG,
                          float* dc_prev, int rows, int R, rfn_stream_t stream) {
  return lstm_cell_bwd(G, c_prev, dh, dc_next, dG, dc_prev, rows, R, (cudaStream_t)stream);
}
int rfn_log_softmax_bwd_f32(const float* lp, size_t ld_lp, const float* dlp, size_t ld_d, float* dx, size_t ld_x, int rows,
                            int V, rfn_stream_t stream) {
  RFN_CHECK_ARG(lp && dlp && dx, "rfn_log_softmax_bwd_f32: null pointer");
  return log_softmax_bwd(lp, ld_lp, dlp, ld_d, dx, ld_x, rows, V, (cudaStream_t)stream);
}
int rfn_embed_f32(const int64_t* tok, int ld_tok, const float* embed, float* x, int rows, int E, int V1, rfn_stream_t stream) {
  RFN_CHECK_ARG(tok && embed && x, "rfn_embed_f32: null pointer");
  return embed_gather_i64(tok, ld_tok, embed, x, rows, E, V1, (cudaStream_t)stream);
}
int rfn_embed_bwd_f32(const int64_t* tok, int ld_tok, const float* dx, float* dE, int rows, int E, int V1, rfn_stream_t stream) {
  RFN_CHECK_ARG(tok && dx && dE, "rfn_embed_bwd_f32: null pointer");
  return embed_bwd(tok, ld_tok, dx, dE, rows, E, V1, (cudaStream_t)stream);
}
int rfn_max_over_steps_f32(const float* in, float* out, int rows, int S, int K, rfn_stream_t stream) {
  RFN_CHECK_ARG(in && out, "rfn_max_over_steps_f32: null pointer");
  return max_over_steps(in, out, rows, S, K, (cudaStream_t)stream);
}
int rfn_max_over_steps_bwd_f32(const float* in, const float* dout, float* din, int rows, int S, int K, rfn_stream_t stream) {
  RFN_CHECK_ARG(in && dout && din, "rfn_max_over_steps_bwd_f32: null pointer");
  return max_over_steps_bwd(in, dout, din, rows, S, K, (cudaStream_t)stream);
}
int rfn_axpby_f32(float alpha, const float* x, float beta, const float* y, float* out, size_t n, rfn_stream_t stream) {
  RFN_CHECK_ARG(x && out, "rfn_axpby_f32: null pointer");
  return axpby(alpha, x, beta, y, out, n, (cudaStream_t)stream);
}
int rfn_xe_loss_bwd_f32(const int64_t* target, const float* mask, int ld_t, int rows, int T, int V, float eps,
                        const float* gout, float* dlp, rfn_stream_t stream) {
  RFN_CHECK_ARG(target && mask && gout && dlp, "rfn_xe_loss_bwd_f32: null pointer");
  return xe_loss_bwd(target, mask, ld_t, rows, T, V, eps, gout, dlp, (size_t)T * V, (size_t)V, (cudaStream_t)stream);
}
int rfn_xe_loss_bwd_strided_f32(const int64_t* target, const float* mask, int ld_t, int rows, int T, int V, float eps,
                                const float* gout, float* dlp, size_t ld_b, size_t ld_s, rfn_stream_t stream) {
  RFN_CHECK_ARG(target && mask && gout && dlp, "rfn_xe_loss_bwd_strided_f32: null pointer");
  return xe_loss_bwd(target, mask, ld_t, rows, T, V, eps, gout, dlp, ld_b, ld_s, (cudaStream_t)stream);
}
int rfn_rl_loss_bwd_f32(const int64_t* seq, const float* reward, const float* lp_all, int ld_lp_rows, int rows, int T, int T1,
                        int V, float entropy_reg, const float* gout, float* dslp, float* dlp_all, rfn_stream_t stream) {
  RFN_CHECK_ARG(seq && reward && lp_all && gout && dslp && dlp_all, "rfn_rl_loss_bwd_f32: null pointer");
  return rl_loss_bwd(seq, reward, lp_all, (size_t)ld_lp_rows, (size_t)V, rows, T, T1, V, entropy_reg, gout, dslp, dlp_all,
                     (cudaStream_t)stream);
}
int rfn_rl_loss_bwd_strided_f32(const int64_t* seq, const float* reward, const float* lp_all, size_t ld_b, size_t ld_s, int rows,
                                int T, int T1, int V, float entropy_reg, const float* gout, float* dslp, float* dlp_all,
                                rfn_stream_t stream) {
  RFN_CHECK_ARG(seq && reward && lp_all && gout && dslp && dlp_all, "rfn_rl_loss_bwd_strided_f32: null pointer");
  return rl_loss_bwd(seq, reward, lp_all, ld_b, ld_s, rows, T, T1, V, entropy_reg, gout, dslp, dlp_all, (cudaStream_t)stream);
}
int rfn_lstm_cell_bwd_multi_f32(const float* G, const float* c_prev, int n_dh, const float* const* dh, const int* ld_dh,
                                const float* mask, float scale, const float* dc_next, float* dG, float* dc_prev, int rows, int R,
                                rfn_stream_t stream) {
  return rfn_lstm_cell_bwd_ex_f32(G, c_prev, n_dh, dh, ld_dh, mask, scale, 0, dc_next, dG, dc_prev, rows, R, stream);
}
int rfn_lstm_cell_bwd_ex_f32(const float* G, const float* c_prev, int n_dh, const float* const* dh, const int* ld_dh,
                             const float* mask, float scale, int maxout, const float* dc_next, float* dG, float* dc_prev, int rows,
                             int R, rfn_stream_t stream) {
  RFN_CHECK_ARG(n_dh >= 0 && n_dh <= 8 && (n_dh == 0 || (dh && ld_dh)), "rfn_lstm_cell_bwd_multi_f32: 0..8 dh sources");
  CellBwdSrcs src{};
  src.n = n_dh;
  for (int s = 0; s < n_dh; ++s) {
    RFN_CHECK_ARG(dh[s] != nullptr, "rfn_lstm_cell_bwd_multi_f32: dh source %d is null", s);
    src.p[s] = dh[s];
    src.ld[s] = ld_dh[s];
  }
  return lstm_cell_bwd_multi(G, c_prev, src, mask, scale, dc_next, dG, dc_prev, rows, R, (cudaStream_t)stream, maxout);
}
/* dX[M,K] (+)= sum_i dY_i[M,N_i] . W_i[N_i,K]: the input gradient of y = sum_i x_i W_i^T, and with several dY_i the sum over
 * the consumers of one input (dH of a fusion step = sum over the J encoders' gate GEMMs) in ONE launch. */
int rfn_linear_bwd_x_f32(int n_src, const float* const* dY, const int* lddy, const float* const* W, const int* ldw,
                         const int* Nc, float* dX, int lddx, int M, int K, int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(n_src >= 1 && n_src <= 3 && dY && lddy && W && ldw && Nc && dX, "rfn_linear_bwd_x_f32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  bool tc = gemm_mode() >= 1 && K >= 128;
  long work = 0;
  for (int s = 0; s < n_src; ++s) {
    tc = tc && tc_general_ok(dY[s], lddy[s], W[s], ldw[s], dX, lddx, K, Nc[s]);
    work += (long)M * K * Nc[s];
  }
  if (tc && work >= (1L << 25)) {
    ProfScope prof__(TAG_GEMM_BWD, st);
    GemmArgs g{};
    for (int s = 0; s < n_src; ++s) g.src[s] = GemmSrc{dY[s], W[s], nullptr, lddy[s], ldw[s], Nc[s]};
    g.nsrc = n_src;
    g.y = dX; g.ldy = lddx; g.M = M; g.N = K; g.accumulate = accumulate ? 1 : 0; g.splitk_ok = 1;
    return gemm_tc_splitk(g, true, tc_passes(gemm_mode()), st);
  }
  for (int s = 0; s < n_src; ++s)
    RFN_TRY(gemm_general(true, false, dY[s], lddy[s], W[s], ldw[s], dX, lddx, M, K, Nc[s], (accumulate || s > 0) ? 1 : 0, st));
  return RFN_OK;
}
int rfn_multilabel_margin_bwd_f32(const float* pred, const int64_t* target, int rows, int K, float weight, const float* gout,
                                  float* dx, rfn_stream_t stream) {
  RFN_CHECK_ARG(pred && target && gout && dx && K <= 48 * 1024, "rfn_multilabel_margin_bwd_f32: bad arguments");
  return multilabel_margin_bwd(pred, target, rows, K, weight, gout, dx, (cudaStream_t)stream);
}

int rfn_mul_scale_f32(float scale, const float* x, const float* m, float* out, size_t n, rfn_stream_t stream) {
  RFN_CHECK_ARG(x && m && out, "rfn_mul_scale_f32: null pointer");
  return mul_scale(scale, x, m, out, n, (cudaStream_t)stream);
}
int rfn_select_token_f32(const float* lp, size_t ld, int rows, int V, const float* uniforms, float temperature,
                         int64_t* tok, float* lp_out, rfn_stream_t stream) {
  RFN_CHECK_ARG(lp && tok && lp_out && temperature > 0.f, "rfn_select_token_f32: bad arguments");
  return lp_select(lp, ld, rows, V, uniforms, temperature, tok, lp_out, (cudaStream_t)stream);
}
int rfn_gather_cols_f32(const float* x, size_t ld, const int64_t* idx, float* out, int rows, rfn_stream_t stream) {
  RFN_CHECK_ARG(x && idx && out, "rfn_gather_cols_f32: null pointer");
  if (rows == 0) return RFN_OK;
  gather_cols_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ld, idx, out, rows);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}
int rfn_scatter_cols_f32(const float* dout, const int64_t* idx, float* dx, size_t ld, int rows, int V, rfn_stream_t stream) {
  RFN_CHECK_ARG(dout && idx && dx, "rfn_scatter_cols_f32: null pointer");
  if (rows == 0) return RFN_OK;
  scatter_cols_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(dout, idx, dx, ld, V);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

int rfn_expand_rows_f32(const float* x, int g, float* out, int rows_out, int R, rfn_stream_t stream) {
  RFN_CHECK_ARG(x && out && g >= 1, "rfn_expand_rows_f32: bad arguments");
  return gather_rows(x, nullptr, g, out, rows_out, R, (cudaStream_t)stream);
}
int rfn_group_sum_f32(const float* x, int g, float* out, int rows_out, int R, rfn_stream_t stream) {
  RFN_CHECK_ARG(x && out && g >= 1, "rfn_group_sum_f32: bad arguments");
  if (rows_out == 0) return RFN_OK;
  const size_t total = (size_t)rows_out * R;
  group_sum_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, rows_out, g, R);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

}  // extern "C"
