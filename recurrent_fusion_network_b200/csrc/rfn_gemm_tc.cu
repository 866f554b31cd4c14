// tcgen05 tensor-core GEMM engine for sm_100a:  y[M,N] = sum_i x_i[M,K_i] . W_i[N,K_i]^T + bias.
//
// One CTA = one 128 x BN output tile (this file) or a 2-CTA cluster = one 256 x 256 tile (rfn_gemm_tc2.cu).
//   warp 0        TMA producer: cp.async.bulk.tensor loads of raw fp32 tiles (K-major, 128-byte rows,
//                 SWIZZLE_128B) of x and W into a multi-stage shared-memory ring, mbarrier complete_tx.
//   warp 1        MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 from shared-memory descriptors.
//                 The tensor core truncates its TF32 inputs, so the raw fp32 tile is the `hi` operand as is:
//                 hi.hi is issued as soon as TMA lands, lo.hi and hi.lo once the splitter has written `lo`.
//                 tcgen05.commit releases the stage / publishes an accumulator chunk.
//   warps 2..9    workers, 256 threads: (a) 3xTF32 splitter: lo = x - trunc_tf32(x) written beside the raw
//                 tile (element-wise, so the TMA swizzle is preserved), fence.proxy.async; (b) chunked drain:
//                 every CH k-blocks the MMA warp switches to the other of two TMEM accumulators and the workers
//                 tcgen05.ld the finished one into round-to-nearest fp32 registers (the tensor core itself
//                 accumulates with truncation, -0.5 ulp per MMA); (c) epilogue: bias + coalesced store staged
//                 through the dead operand ring, or the fused additive-attention score
//                 score[m] = sum_n w[n] tanh(acc + b[n] + g[m/Natt, n]) (U_a A never reaches HBM), or the fused
//                 vocabulary epilogue (row max, sum-exp and top-k of the logits; logits never reach HBM).
//   PASSES == 1   single-pass TF32 (no splitter, one accumulator): the reduced-precision mode.
// Both operands are K-major so activations (row-major) and nn.Linear weights (out,in) are consumed exactly as
// torch stores them; ragged M / N / K edges are zero-filled by TMA.
#include <cuda.h>

#include <algorithm>

#include <atomic>

#include "rfn_internal.cuh"
#include "rfn_tc_ptx.cuh"
#include "rfn_tc_epilogue.cuh"
#include "rfn_tc_args.cuh"

namespace rfn {


template <int BN, int PASSES>
struct TcSmem {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int TILE_BYTES = TC_A_BYTES + B_BYTES;          // one landed fp32 tile pair
  static constexpr int STAGE_BYTES = TILE_BYTES * (PASSES == 3 ? 2 : 1);
};

// The tensor core adds into its fp32 accumulator with truncation (measured on B200: -0.5 ulp per
// tcgen05.mma, so a K = 2048 chain drifts by 1.5e-5 relative).  The fp32-equivalent mode therefore
// keeps tensor-core chains short: every CH k-blocks the MMA warp switches to the other of two TMEM
// accumulators and the worker warps drain the finished one into round-to-nearest fp32 registers.
template <int BN, int STAGES, int PASSES, int CH, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcArgs a) {
  using S = TcSmem<BN, PASSES>;
  constexpr int NWORK = 8;                       // worker warps 2..9
  constexpr int COLS = BN / 2;                   // columns per worker thread (two warps share a lane quarter)
  constexpr bool DRAIN = (PASSES == 3);
  constexpr int NBUF = DRAIN ? 2 : 1;
  constexpr int TMEM_COLS = (BN * NBUF <= 128) ? 128 : (BN * NBUF <= 256 ? 256 : 512);
  static_assert(BN * NBUF <= 512 && COLS % 32 == 0, "accumulators must fit the 512 TMEM columns");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // round up to the 1024-byte alignment SWIZZLE_128B needs, keeping the pointer in the shared window
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full = bars;                 // TMA landed
  uint64_t* conv = bars + STAGES;        // hi/lo split done
  uint64_t* empty = bars + 2 * STAGES;   // MMAs that read the stage retired
  uint64_t* cfull = bars + 3 * STAGES;   // [2] accumulator chunk complete
  uint64_t* drained = cfull + 2;         // [2] accumulator buffer drained into registers
  uint32_t* tmem_slot = (uint32_t*)(drained + 2);
  float* s_bias = (float*)((uint8_t*)bars + 512);   // [BN] bias summed over the sources (16-byte aligned)
  float* s_wv = s_bias + 256;                // [BN] att_h_2_out weights (score epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;

  int total_kb = 0;
  for (int s = 0; s < a.nsrc; ++s) total_kb += (a.K[s] + TC_BK - 1) / TC_BK;
  // split-K (store epilogue only): blockIdx.z owns k-blocks [kb_first, kb_first + total_kb) of the concatenated sources
  int kb_first = 0;
  if (EPI == 0 && a.ksplit > 1) {
    const int kper = (total_kb + a.ksplit - 1) / a.ksplit;
    kb_first = (int)blockIdx.z * kper;
    total_kb = min(total_kb, kb_first + kper) - kb_first;
  }
  const bool b_mn = (EPI == 0) && a.b_mn;
  const int nchunk = DRAIN ? (total_kb + CH - 1) / CH : 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nsrc; ++s) { tma_prefetch_desc(&a.tm_x[s]); tma_prefetch_desc(&a.tm_w[s]); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&conv[s]), NWORK);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&cfull[b]), 1);
      mbar_init(smem_u32(&drained[b]), NWORK);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0, gkb = 0;
      for (int s = 0; s < a.nsrc; ++s) {
        const int nkb = (a.K[s] + TC_BK - 1) / TC_BK;
        for (int kb = 0; kb < nkb; ++kb, ++gkb) {
          if (gkb < kb_first || gkb >= kb_first + total_kb) continue;
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          ++it;
          mbar_wait(smem_u32(&empty[st]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full[st]);
          mbar_arrive_expect_tx(fb, (uint32_t)S::TILE_BYTES);
          uint8_t* stage = smem + st * S::STAGE_BYTES;
          tma_load_2d(&a.tm_x[s], fb, smem_u32(stage), kb * TC_BK, m0);
          if (b_mn) {   // (K, N) row-major W: BN / 32 boxes of 32 contraction rows x 32 columns, one 4 KB group each
#pragma unroll
            for (int q = 0; q < BN / 32; ++q)
              tma_load_2d(&a.tm_w[s], fb, smem_u32(stage + TC_A_BYTES + q * 4096), n0 + q * 32, kb * TC_BK);
          } else {
            tma_load_2d(&a.tm_w[s], fb, smem_u32(stage + TC_A_BYTES), kb * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_tf32(BN) | (b_mn ? IDESC_B_MN : 0u);
    const uint64_t bstep = b_mn ? 64ull : 2ull;   // descriptor advance per k-step of 8: 8 rows of 128 bytes / 32 bytes
    for (int it = 0; it < total_kb; ++it) {
      const int st = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      const int c = DRAIN ? it / CH : 0;
      const int b = c & 1;
      const bool chunk_start = DRAIN ? (it % CH == 0) : (it == 0);
      if (DRAIN && chunk_start && c >= 2) mbar_wait(smem_u32(&drained[b]), (uint32_t)((c >> 1) - 1) & 1u);
      // The tensor core truncates TF32 inputs (low 13 mantissa bits ignored; measured), so the raw fp32
      // tile IS the hi operand: hi.hi starts as soon as TMA lands, the cross terms wait for the splitter.
      mbar_wait(smem_u32(&full[st]), ph);
      tc_fence_after();
      const uint32_t td = tmem_base + (uint32_t)(b * BN);
      const uint32_t sa = smem_u32(smem + st * S::STAGE_BYTES);
      const uint64_t da_hi = make_desc_sw128(sa);
      const uint64_t db_hi = b_mn ? make_desc_sw128_mn(sa + TC_A_BYTES) : make_desc_sw128(sa + TC_A_BYTES);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);  // 8 tf32 = 32 bytes = 2 x 16-byte units inside the swizzle atom
          umma_tf32(td, da_hi + adv, db_hi + k * bstep, idesc, (chunk_start && k == 0) ? 0u : 1u);
        }
      }
      if (PASSES == 3) {
        mbar_wait(smem_u32(&conv[st]), ph);
        tc_fence_after();
      }
      if (lane == 0) {
        if (PASSES == 3) {
          const uint64_t da_lo = make_desc_sw128(sa + S::TILE_BYTES);
          const uint64_t db_lo = b_mn ? make_desc_sw128_mn(sa + S::TILE_BYTES + TC_A_BYTES)
                                      : make_desc_sw128(sa + S::TILE_BYTES + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            umma_tf32(td, da_lo + adv, db_hi + k * bstep, idesc, 1u);
            umma_tf32(td, da_hi + adv, db_lo + k * bstep, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&empty[st]));   // stage reusable once these MMAs retire
        const bool chunk_end = DRAIN ? (it % CH == CH - 1 || it == total_kb - 1) : (it == total_kb - 1);
        if (chunk_end) umma_commit(smem_u32(&cfull[b]));
      }
      __syncwarp();
    }
  } else {
    // ===================== workers: 3xTF32 splitter + accumulator drain + epilogue =====================
    const int wq = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which half of the BN columns this warp owns
    const int wt = threadIdx.x - 64;           // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * COLS);
    float acc[COLS];
#pragma unroll
    for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
    tc_stage_bias<BN>(a, n0, wt, s_bias, EPI == 1 ? s_wv : nullptr, kb_first == 0);

    auto drain = [&](int d) {
      const int b = d & 1;
      mbar_wait(smem_u32(&cfull[b]), (uint32_t)(d >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)(b * BN + c0), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c0 + i] += v[i];   // round-to-nearest fp32 add
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&drained[b]));
    };

    int next_drain = 0;
    if (DRAIN) {
      for (int it = 0; it < total_kb; ++it) {
        const int st = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(smem_u32(&full[st]), ph);
        const uint4* hi = reinterpret_cast<const uint4*>(smem + st * S::STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + st * S::STAGE_BYTES + S::TILE_BYTES);
#pragma unroll 4
        for (int i = wt; i < S::TILE_BYTES / 16; i += NWORK * 32) {
          const uint4 v = hi[i];   // raw fp32; the MMA reads it as hi = trunc_tf32(x)
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u));
          lo[i] = l;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&conv[st]));
        // chunks whose last k-block retired (implied by this stage having been refilled) drain for free
        while (next_drain < nchunk && min((next_drain + 1) * CH, total_kb) - 1 <= it - STAGES) drain(next_drain++);
      }
    }
    while (next_drain < nchunk) drain(next_drain++);

    // ----- epilogue -----
    const int m = m0 + wq * 32 + lane;
    const int nb = n0 + half * COLS;
    if (EPI == 0) {
      // all MMAs have retired (last chunk drained), so the operand ring is free: stage the tile through it
      float* stage = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * (COLS + 4);
      tc_epilogue_store<COLS>(acc, stage, a, m0 + wq * 32, nb, lane, s_bias + half * COLS, a.ksplit > 1);
    } else if (EPI == 1) {
      // fused additive-attention score (misc/AttentionModelCore.py:37-42):
      //   score[tile][m] = sum_{n in this thread's columns} w[n] * tanh(acc[m,n] + U_b[n] + g[m / natt, n])
      // partials are stored per column slice and summed in a fixed order by the attention kernel
      // g rows (one per image) are staged through the dead operand ring: a warp's 32 consecutive rows span
      // mg_first .. mg_last; reading g straight from global memory cost one L2 round trip per float4
      // (registers are full of accumulators, so the loads cannot be hoisted) -- ~10k cycles per tile.
      const int mg = (m < a.M ? m : a.M - 1) / a.natt;
      const int mg_first = __shfl_sync(0xffffffffu, mg, 0);
      const int nslots = __shfl_sync(0xffffffffu, mg, 31) - mg_first + 1;
      float* gst = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * COLS;
      for (int sl = 0; sl < nslots; ++sl) {
#pragma unroll
        for (int c = lane * 4; c < COLS; c += 128) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nb + c + 3 < a.N) v = *reinterpret_cast<const float4*>(a.g + (size_t)(mg_first + sl) * a.ldg + nb + c);
          *reinterpret_cast<float4*>(gst + sl * COLS + c) = v;
        }
      }
      __syncwarp();
      const float* gr = gst + (mg - mg_first) * COLS;
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < COLS / 4; ++q) {   // branch-free: columns beyond N carry zero weight (s_wv) and zero inputs
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + half * COLS + q * 4);
        const float4 gg = *reinterpret_cast<const float4*>(gr + q * 4);
        const float4 ww = *reinterpret_cast<const float4*>(s_wv + half * COLS + q * 4);
        part = fmaf(ww.x, tc_tanh(acc[q * 4 + 0] + b4.x + gg.x), part);
        part = fmaf(ww.y, tc_tanh(acc[q * 4 + 1] + b4.y + gg.y), part);
        part = fmaf(ww.z, tc_tanh(acc[q * 4 + 2] + b4.z + gg.z), part);
        part = fmaf(ww.w, tc_tanh(acc[q * 4 + 3] + b4.w + gg.w), part);
      }
      const int slice = blockIdx.x * 2 + half;   // N / COLS slices in total
      if (m < a.M) a.score[(size_t)slice * a.M + m] = part;
    } else {
      // fused vocabulary epilogue: logits never reach HBM.  Per thread (row m, COLS columns): running max,
      // sum exp(x - max) and the top-k logits with ties -> lower index; vocab_merge_kernel combines the slices.
      // branch-free on purpose: per-column `if`s compile to BSSY/BRA/BSYNC triplets (61k cycles per tile measured)
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < COLS; ++i) {
        const float v = acc[i] + s_bias[half * COLS + i];
        acc[i] = (nb + i < a.N) ? v : -INFINITY;      // columns beyond N never win and contribute exp(-inf) = 0
        mx = fmaxf(mx, acc[i]);
      }
      const float mref = (mx == -INFINITY) ? 0.f : mx;  // a slice entirely beyond N
      float se = 0.f;
#pragma unroll
      for (int i = 0; i < COLS; ++i) se += expf(acc[i] - mref);
      const int slice = blockIdx.x * 2 + half;
      if (m < a.M) {
        a.st_max[(size_t)slice * a.M + m] = mx;
        a.st_sum[(size_t)slice * a.M + m] = se;
      }
      float pv = INFINITY;
      int pi = -1;
      for (int r = 0; r < a.ktop; ++r) {   // k selection passes over the register tile (ties -> lower index)
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < COLS; ++i) {
          const float v = acc[i];
          const int n = nb + i;
          const bool after = (v < pv) | ((v == pv) & (n > pi));
          const bool take = after & ((v > bv) | ((v == bv) & (n < bi)));
          bv = take ? v : bv;
          bi = take ? n : bi;
        }
        if (m < a.M) {
          a.st_val[((size_t)slice * a.M + m) * a.ktop + r] = bv;
          a.st_idx[((size_t)slice * a.M + m) * a.ktop + r] = bi;
        }
        pv = bv; pi = bi;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map over a row-major (rows, K) matrix with leading dimension ld: box = 32 floats x box_rows
int tc_make_map(CUtensorMap* tm, const float* base, int rows, int K, int ld, int box_rows, bool atom32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RFN_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
    return RFN_ERR_CUDA;
  }
  return RFN_OK;
}

int tc_make_map_bf16_tiles(CUtensorMap* tm, const void* base, int Ng, int Kp) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RFN_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Kp * 8, (cuuint64_t)Ng};
  cuuint64_t gstr[1] = {(cuuint64_t)Kp * 8 * 2};
  cuuint32_t box[2] = {256, 16};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 tiles) failed (%d) Ng=%d Kp=%d", (int)r, Ng, Kp);
    return RFN_ERR_CUDA;
  }
  return RFN_OK;
}

template <int BN, int STAGES, int PASSES, int CH, int EPI>
static int launch_tc_epi(const TcArgs& t, cudaStream_t st) {
  using S = TcSmem<BN, PASSES>;
  const size_t smem = (size_t)STAGES * S::STAGE_BYTES + 1024 + 4096;
  static bool configured = false;
  if (!configured) {
    RFN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, PASSES, CH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((t.N + BN - 1) / BN, (t.M + TC_BM - 1) / TC_BM, (EPI == 0 && t.ksplit > 1) ? t.ksplit : 1);
  gemm_tc_kernel<BN, STAGES, PASSES, CH, EPI><<<grid, TC_THREADS, smem, st>>>(t);
  RFN_LAUNCH_CHECK();
  count_engine(t.ksplit > 1 ? ENG_TC1_SPLITK : ENG_TC1);
  return RFN_OK;
}

template <int BN, int STAGES, int PASSES, int CH>
static int launch_tc(const TcArgs& t, cudaStream_t st) {
  if (t.epi == 0) return launch_tc_epi<BN, STAGES, PASSES, CH, 0>(t, st);
  if (BN == 256 && t.epi == 1) return launch_tc_epi<256, STAGES, PASSES, CH, 1>(t, st);
  if (BN == 256 && t.epi == 2) return launch_tc_epi<256, STAGES, PASSES, CH, 2>(t, st);
  set_error("gemm_tc: fused epilogues need the 256-wide tile");
  return RFN_ERR_INVALID;
}

static std::atomic<long long*> g_tc_dbg{nullptr};
static std::atomic<int> g_tc_dbg_epi{-1};   // -1: every epilogue kind
static std::atomic<int> g_tc_cluster{2};   // default: persistent 2-CTA clusters for large tensor-engine GEMMs
int launch_tc2(const TcArgs& t, int passes, cudaStream_t st);   // rfn_gemm_tc2.cu
int launch_tc2p(const TcArgs& t, cudaStream_t st);               // rfn_gemm_tc2p.cu (persistent, 3xTF32)
static bool use_cluster(const GemmArgs& a) { return g_tc_cluster.load() != 0 && a.N >= 256 && a.M >= 256; }

bool gemm_tc_supported(const GemmArgs& a) {
  if (a.N % 4 != 0 || a.ldy % 4 != 0 || ((uintptr_t)a.y % 16) != 0) return false;
  for (int s = 0; s < a.nsrc; ++s) {
    const GemmSrc& g = a.src[s];
    if (g.K % 4 != 0 || g.ldx % 4 != 0 || g.ldw % 4 != 0) return false;
    if (((uintptr_t)g.x % 16) != 0 || ((uintptr_t)g.w % 16) != 0) return false;
    if (g.bias && ((uintptr_t)g.bias % 16) != 0) return false;
  }
  return true;
}

// passes: 3 = 3xTF32 (fp32-equivalent), 1 = single-pass TF32.  score != nullptr selects the fused
// attention-score epilogue (then y is unused); it receives tc_score_slices(N) partial rows of M.
int tc_score_slices(int N) { return ((N + 255) / 256) * 2; }

int gemm_tc(const GemmArgs& a, int passes, const float* g, int ldg, const float* wv, float* score, int natt,
            cudaStream_t st) {
  ProfScope prof__(TAG_GEMM_OTHER, st);
  const bool bf16x = (passes == 2);   // engine mode 3: BF16 cross terms, persistent kernel only
  if (bf16x) passes = 3;
  RFN_CHECK_ARG(gemm_tc_supported(a), "gemm_tc: operands must be 16-byte aligned with K, N, ld multiples of 4");
  if (a.M == 0) return RFN_OK;
  TcArgs t{};
  t.nsrc = a.nsrc;
  // 128 x 256 tiles unless that leaves most of the 148 SMs idle (few rows): then 128 x 128 doubles the CTA count
  const long tiles256 = (long)((a.M + TC_BM - 1) / TC_BM) * ((a.N + 255) / 256);
  // 2-CTA pairs (each CTA lands 128 rows of W) whenever the problem has >= 256 rows and columns: even when one
  // launch does not fill the SMs, the encoder side streams run several such GEMMs concurrently
  const bool cluster = use_cluster(a);
  const int bn = (score || cluster) ? 256 : ((a.N > 128 && tiles256 >= 148) ? 256 : 128);
  for (int s = 0; s < a.nsrc; ++s) {
    RFN_TRY(tc_make_map(&t.tm_x[s], a.src[s].x, a.M, a.src[s].K, a.src[s].ldx, TC_BM));
    RFN_TRY(tc_make_map(&t.tm_w[s], a.src[s].w, a.N, a.src[s].K, a.src[s].ldw, cluster ? 128 : bn));
    t.K[s] = a.src[s].K;
    t.bias[s] = a.src[s].bias;
  }
  t.y = a.y; t.ldy = a.ldy; t.M = a.M; t.N = a.N; t.accumulate = a.accumulate;
  t.epi = score ? 1 : 0;
  t.g = g; t.ldg = ldg; t.wv = wv; t.score = score; t.natt = natt > 0 ? natt : 1;
  t.dbg = (g_tc_dbg_epi.load() < 0 || g_tc_dbg_epi.load() == t.epi) ? g_tc_dbg.load() : nullptr;
  // few output tiles and a long contraction (weight gradients dU = dP^T . A over rows x locations): split K over
  // clusters so that one wave of 74 cluster slots is filled; partial tiles are added with red.global.add
  if (cluster && t.epi == 0 && a.nsrc == 1 && a.splitk_ok) {
    const int tiles = ((a.N + 255) / 256) * ((a.M + 255) / 256);
    const int nkb = (a.src[0].K + TC_BK - 1) / TC_BK;
    int ks = std::min(74 / std::max(tiles, 1), nkb / 16);
    if (ks > 1) {
      const int kper = (nkb + ks - 1) / ks;
      ks = (nkb + kper - 1) / kper;             // every slice non-empty
      if (ks > 1) {
        t.ksplit = ks;
        if (!a.accumulate)
          RFN_CUDA(cudaMemset2DAsync(a.y, (size_t)a.ldy * sizeof(float), 0, (size_t)a.N * sizeof(float), (size_t)a.M, st));
      }
    }
  }
  // persistent clusters for the fused-epilogue GEMMs; the plain-store GEMMs keep the 3-stage one-tile kernel
  // (the persistent store epilogue needs staging memory that costs a pipeline stage: measured slower)
  if (cluster && g_tc_cluster.load() >= 2 && passes == 3 && (t.epi != 0 || (bf16x && !t.ksplit && a.nsrc == 1))) {
    void* scratch = nullptr;
    if (bf16x && a.nsrc == 1) {   // engine mode 3 (single-source GEMMs): bf16 cross terms
      t.bf16x = 1;
      RFN_TRY(tc2p_prepare_bf16_w(t, a.src[0].w, a.src[0].ldw, a.N, a.src[0].K, &scratch, st));
    }
    const int rc = launch_tc2p(t, st);
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
  }
  if (cluster) return launch_tc2(t, passes, st);
  if (passes == 3) return bn == 256 ? launch_tc<256, 2, 3, 4>(t, st) : launch_tc<128, 3, 3, 4>(t, st);
  return bn == 256 ? launch_tc<256, 4, 1, 1>(t, st) : launch_tc<128, 6, 1, 1>(t, st);
}

// Small-row / general-layout route (training): 128 x 128 tiles of the 1-CTA kernel with the contraction split over
// blockIdx.z so that ~one wave of CTAs streams the weights (a GEMM with <= 127 rows has only N / 128 output tiles);
// partial tiles are added atomically.  b_mn: W given as (K, N) row-major (dX = dY . W on nn.Linear weights).
int gemm_tc_splitk(const GemmArgs& a, bool b_mn, int passes, cudaStream_t st) {
  ProfScope prof__(TAG_GEMM_OTHER, st);
  if (passes == 2) passes = 3;
  if (a.M == 0 || a.N == 0) return RFN_OK;
  TcArgs t{};
  t.nsrc = a.nsrc;
  int nkb = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const GemmSrc& g = a.src[s];
    RFN_TRY(tc_make_map(&t.tm_x[s], g.x, a.M, g.K, g.ldx, TC_BM));
    if (b_mn) RFN_TRY(tc_make_map(&t.tm_w[s], g.w, g.K, a.N, g.ldw, 32, true));   // rows = contraction, inner = N
    else RFN_TRY(tc_make_map(&t.tm_w[s], g.w, a.N, g.K, g.ldw, 128));
    t.K[s] = g.K;
    t.bias[s] = g.bias;
    nkb += (g.K + TC_BK - 1) / TC_BK;
  }
  t.y = a.y; t.ldy = a.ldy; t.M = a.M; t.N = a.N; t.accumulate = a.accumulate;
  t.epi = 0;
  t.b_mn = b_mn ? 1 : 0;
  const int tiles = ((a.M + TC_BM - 1) / TC_BM) * ((a.N + 127) / 128);
  int ks = std::max(1, std::min(148 / std::max(tiles, 1), nkb / 4));
  if (ks > 1) {
    const int kper = (nkb + ks - 1) / ks;
    ks = (nkb + kper - 1) / kper;               // every slice non-empty
  }
  if (ks > 1) {
    t.ksplit = ks;
    if (!a.accumulate)
      RFN_CUDA(cudaMemset2DAsync(a.y, (size_t)a.ldy * sizeof(float), 0, (size_t)a.N * sizeof(float), (size_t)a.M, st));
  }
  return passes == 3 ? launch_tc<128, 3, 3, 4>(t, st) : launch_tc<128, 6, 1, 1>(t, st);
}

// logits GEMM with the fused vocabulary epilogue: returns per-slice statistics instead of logits
int gemm_tc_vocab(const GemmArgs& a, int passes, float* st_max, float* st_sum, float* st_val, int32_t* st_idx, int ktop,
                  cudaStream_t st) {
  ProfScope prof__(TAG_GEMM_LOGIT, st);
  const bool bf16x = (passes == 2);
  if (bf16x) passes = 3;
  RFN_CHECK_ARG(gemm_tc_supported(a) && a.nsrc == 1 && ktop >= 1 && ktop <= RFN_MAX_BEAM, "gemm_tc_vocab: unsupported arguments");
  if (a.M == 0) return RFN_OK;
  TcArgs t{};
  t.nsrc = 1;
  const bool cluster = use_cluster(a);
  RFN_TRY(tc_make_map(&t.tm_x[0], a.src[0].x, a.M, a.src[0].K, a.src[0].ldx, TC_BM));
  RFN_TRY(tc_make_map(&t.tm_w[0], a.src[0].w, a.N, a.src[0].K, a.src[0].ldw, cluster ? 128 : 256));
  t.K[0] = a.src[0].K;
  t.bias[0] = a.src[0].bias;
  t.M = a.M; t.N = a.N;
  t.epi = 2;
  t.st_max = st_max; t.st_sum = st_sum; t.st_val = st_val; t.st_idx = st_idx; t.ktop = ktop;
  t.dbg = (g_tc_dbg_epi.load() < 0 || g_tc_dbg_epi.load() == 2) ? g_tc_dbg.load() : nullptr;
  if (cluster && g_tc_cluster.load() >= 2 && passes == 3) {
    void* scratch = nullptr;
    if (bf16x) {
      t.bf16x = 1;
      RFN_TRY(tc2p_prepare_bf16_w(t, a.src[0].w, a.src[0].ldw, a.N, a.src[0].K, &scratch, st));
    }
    const int rc = launch_tc2p(t, st);
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
  }
  if (cluster) return launch_tc2(t, passes, st);
  return passes == 3 ? launch_tc<256, 2, 3, 4>(t, st) : launch_tc<256, 4, 1, 1>(t, st);
}

}  // namespace rfn

extern "C" int rfn_set_tc_cluster(int on) {
  rfn::g_tc_cluster.store(on < 0 ? 0 : (on > 2 ? 2 : on));
  return RFN_OK;
}
extern "C" int rfn_get_tc_cluster(void) { return rfn::g_tc_cluster.load(); }
// debugging aid: device buffer of 8 clock64 stamps per CTA filled by the 2-CTA kernel (NULL = off)
extern "C" int rfn_debug_set_timeline(long long* d_buf, int epilogue_kind) {
  rfn::g_tc_dbg_epi.store(epilogue_kind);
  rfn::g_tc_dbg.store(d_buf);
  return RFN_OK;
}
