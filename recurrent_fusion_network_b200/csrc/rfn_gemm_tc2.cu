// 2-CTA (cta_group::2) variant of the tcgen05 GEMM engine: a cluster of two CTAs on one TPC computes a
// 256 x 256 output tile.  Each CTA loads its own 128 rows of x and HALF of the W tile (128 of the 256
// output columns' rows); one tcgen05.mma.cta_group::2 issued by the leader CTA reads both CTAs' shared
// memory and writes rows 0..127 into the leader's TMEM and rows 128..255 into the peer's.  Per CTA the
// tensor core therefore fetches 8 KB instead of 12 KB of operands per MMA, TMA lands 32 KB instead of
// 48 KB per k-block and the 3xTF32 splitter touches a third less shared memory -- shared-memory bandwidth
// is what bounds the 1-CTA kernel (profiles/r1_gemm_tc_ncu_v1.txt).
//
// Cross-CTA synchronisation:
//   ready[st]   (leader)  <- 16 worker warps (8 local arrives + 8 remote arrives from the peer) once their
//                            half of the stage is landed and split
//   empty[st]   (both)    <- tcgen05.commit multicast: the MMAs that read the stage have retired
//   cfull[b]    (both)    <- tcgen05.commit multicast: accumulator chunk complete
//   drained[b]  (leader)  <- 16 worker warps after moving their TMEM half into fp32 registers
// plus cluster barriers after initialisation and before teardown.
#include <cuda.h>

#include "rfn_internal.cuh"
#include "rfn_tc_args.cuh"
#include "rfn_tc_ptx.cuh"
#include "rfn_tc_epilogue.cuh"

namespace rfn {

constexpr int T2_BN = 256;                              // output columns per cluster tile
constexpr int T2_BH = T2_BN / 2;                        // W rows held by each CTA
constexpr int T2_TILE_BYTES = TC_A_BYTES + T2_BH * 128; // 32 KB landed per CTA per k-block

template <int STAGES, int PASSES, int CH, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ TcArgs a, int n_tiles) {
  constexpr int NWORK = 8;
  constexpr int COLS = T2_BN / 2;
  constexpr bool DRAIN = (PASSES == 3);
  constexpr int NBUF = DRAIN ? 2 : 1;
  constexpr int TMEM_COLS = (T2_BN * NBUF <= 256) ? 256 : 512;
  constexpr int STAGE_BYTES = T2_TILE_BYTES * (PASSES == 3 ? 2 : 1);
  constexpr uint32_t READY_COUNT = (PASSES == 3) ? 2 * NWORK : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // local: TMA landed
  uint64_t* ready = bars + STAGES;       // leader: both halves landed (and split)
  uint64_t* empty = bars + 2 * STAGES;   // both: stage free
  uint64_t* cfull = bars + 3 * STAGES;   // [2] both: accumulator chunk complete
  uint64_t* drained = cfull + 2;         // [2] leader: both halves drained
  uint32_t* tmem_slot = (uint32_t*)(drained + 2);
  float* s_bias = (float*)((uint8_t*)bars + 512);   // [BN] bias summed over the sources (16-byte aligned)
  float* s_wv = s_bias + 256;                // [BN] att_h_2_out weights (score epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int tiles_all = n_tiles * ((a.M + 2 * TC_BM - 1) / (2 * TC_BM));
  const int cid = (int)(blockIdx.x >> 1) % tiles_all;
  const int kz = (int)(blockIdx.x >> 1) / tiles_all;                  // split-K slice (0 when ksplit == 1)
  const int n0 = (cid % n_tiles) * T2_BN;
  const int m0 = (cid / n_tiles) * (2 * TC_BM) + (int)rank * TC_BM;   // this CTA's 128 rows

  int total_kb = 0;
  for (int s = 0; s < a.nsrc; ++s) total_kb += (a.K[s] + TC_BK - 1) / TC_BK;
  int kb_first = 0;                                                   // split-K: one source, k-blocks [kb_first, +total_kb)
  if (EPI == 0 && a.ksplit > 1) {
    const int kper = (total_kb + a.ksplit - 1) / a.ksplit;
    kb_first = kz * kper;
    total_kb = min(total_kb, kb_first + kper) - kb_first;
  }
  const int nchunk = DRAIN ? (total_kb + CH - 1) / CH : 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nsrc; ++s) { tma_prefetch_desc(&a.tm_x[s]); tma_prefetch_desc(&a.tm_w[s]); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&ready[s]), READY_COUNT);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&cfull[b]), 1);
      mbar_init(smem_u32(&drained[b]), 2 * NWORK);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its own half) =====================
    if (lane == 0) {
      int it = 0;
      for (int s = 0; s < a.nsrc; ++s) {
        const int nkb = (EPI == 0 && a.ksplit > 1) ? kb_first + total_kb : (a.K[s] + TC_BK - 1) / TC_BK;
        for (int kb = kb_first; kb < nkb; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(smem_u32(&empty[st]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full[st]);
          mbar_arrive_expect_tx(fb, (uint32_t)T2_TILE_BYTES);
          uint8_t* stage = smem + st * STAGE_BYTES;
          tma_load_2d(&a.tm_x[s], fb, smem_u32(stage), kb * TC_BK, m0);
          tma_load_2d(&a.tm_w[s], fb, smem_u32(stage + TC_A_BYTES), kb * TC_BK, n0 + (int)rank * T2_BH);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32_m(2 * TC_BM, T2_BN);
      for (int it = 0; it < total_kb; ++it) {
        const int st = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int c = DRAIN ? it / CH : 0;
        const int b = c & 1;
        const bool chunk_start = DRAIN ? (it % CH == 0) : (it == 0);
        if (DRAIN && chunk_start && c >= 2) mbar_wait(smem_u32(&drained[b]), (uint32_t)((c >> 1) - 1) & 1u);
        mbar_wait(smem_u32(&ready[st]), ph);
        tc_fence_after();
        if (dbg && lane == 0) { if (it == 0) dbg[2] = clock64(); if (it == total_kb - 1) dbg[3] = clock64(); }
        if (lane == 0) {
          const uint32_t td = tmem_base + (uint32_t)(b * T2_BN);
          const uint32_t sa = smem_u32(smem + st * STAGE_BYTES);
          const uint64_t da_hi = make_desc_sw128(sa);
          const uint64_t db_hi = make_desc_sw128(sa + TC_A_BYTES);
          const uint64_t da_lo = make_desc_sw128(sa + T2_TILE_BYTES);
          const uint64_t db_lo = make_desc_sw128(sa + T2_TILE_BYTES + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            umma2_tf32(td, da_hi + adv, db_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
            if (PASSES == 3) {
              umma2_tf32(td, da_lo + adv, db_hi + adv, idesc, 1u);
              umma2_tf32(td, da_hi + adv, db_lo + adv, idesc, 1u);
            }
          }
          umma2_commit(smem_u32(&empty[st]));
          const bool chunk_end = DRAIN ? (it % CH == CH - 1 || it == total_kb - 1) : (it == total_kb - 1);
          if (chunk_end) umma2_commit(smem_u32(&cfull[b]));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== workers: splitter + accumulator drain + epilogue (both CTAs) =====================
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int wt = threadIdx.x - 64;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * COLS);
    float acc[COLS];
#pragma unroll
    for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
    tc_stage_bias<T2_BN>(a, n0, wt, s_bias, EPI == 1 ? s_wv : nullptr, kz == 0);

    auto signal = [&](uint64_t* bar) {   // arrive on the LEADER's barrier
      if (leader) mbar_arrive(smem_u32(bar)); else mbar_arrive_remote(smem_u32(bar), 0);
    };
    auto drain = [&](int d) {
      const int b = d & 1;
      mbar_wait(smem_u32(&cfull[b]), (uint32_t)(d >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)(b * T2_BN + c0), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c0 + i] += v[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) signal(&drained[b]);
    };

    int next_drain = 0;
    for (int it = 0; it < total_kb; ++it) {
      const int st = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      if (PASSES == 3) {
        mbar_wait(smem_u32(&full[st]), ph);
        const uint4* hi = reinterpret_cast<const uint4*>(smem + st * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + st * STAGE_BYTES + T2_TILE_BYTES);
#pragma unroll 4
        for (int i = wt; i < T2_TILE_BYTES / 16; i += NWORK * 32) {
          const uint4 v = hi[i];
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u));
          lo[i] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) signal(&ready[st]);
      } else if (warp == 2) {
        mbar_wait(smem_u32(&full[st]), ph);   // forward "my half has landed" to the leader
        if (lane == 0) signal(&ready[st]);
      }
      if (DRAIN)
        while (next_drain < nchunk && min((next_drain + 1) * CH, total_kb) - 1 <= it - STAGES) drain(next_drain++);
    }
    while (next_drain < nchunk) drain(next_drain++);
    if (dbg && threadIdx.x == 64) dbg[4] = clock64();

    // ----- epilogue (identical to the 1-CTA kernel; this CTA owns rows m0 .. m0+127) -----
    const int m = m0 + wq * 32 + lane;
    const int nb = n0 + half * COLS;
    if (EPI == 0) {
      // all MMAs have retired (last chunk drained), so the operand ring is free: stage the tile through it
      float* stage = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * (COLS + 4);
      tc_epilogue_store<COLS>(acc, stage, a, m0 + wq * 32, nb, lane, s_bias + half * COLS, a.ksplit > 1);
    } else if (EPI == 1) {
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
      const int slice = (n0 / T2_BN) * 2 + half;
      if (m < a.M) a.score[(size_t)slice * a.M + m] = part;
    } else {
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
      const int slice = (n0 / T2_BN) * 2 + half;
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
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();
  }
  __syncthreads();
  cluster_sync_all();   // the peer may still be signalling our barriers / the leader reading our smem until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
  if (dbg && threadIdx.x == 32) dbg[6] = clock64();
}

template <int STAGES, int PASSES, int CH, int EPI>
static int launch_tc2_epi(const TcArgs& t, cudaStream_t st) {
  constexpr int STAGE_BYTES = T2_TILE_BYTES * (PASSES == 3 ? 2 : 1);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 4096;
  static bool configured = false;
  if (!configured) {
    RFN_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<STAGES, PASSES, CH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int n_tiles = (t.N + T2_BN - 1) / T2_BN;
  const int n_pairs = (t.M + 2 * TC_BM - 1) / (2 * TC_BM);
  cudaLaunchConfig_t cfg{};
  const int ksplit = (EPI == 0 && t.ksplit > 1) ? t.ksplit : 1;
  cfg.gridDim = dim3((unsigned)(2 * n_tiles * n_pairs * ksplit), 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RFN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<STAGES, PASSES, CH, EPI>, t, n_tiles));
  RFN_LAUNCH_CHECK();
  count_engine(ENG_TC2);
  return RFN_OK;
}

// the args were prepared by gemm_tc (tensor maps with 128-row boxes for both operands)
int launch_tc2(const TcArgs& t, int passes, cudaStream_t st) {
  if (passes == 3) {
    if (t.epi == 0) return launch_tc2_epi<3, 3, 4, 0>(t, st);
    if (t.epi == 1) return launch_tc2_epi<3, 3, 4, 1>(t, st);
    return launch_tc2_epi<3, 3, 4, 2>(t, st);
  }
  if (t.epi == 0) return launch_tc2_epi<6, 1, 1, 0>(t, st);
  if (t.epi == 1) return launch_tc2_epi<6, 1, 1, 1>(t, st);
  return launch_tc2_epi<6, 1, 1, 2>(t, st);
}

}  // namespace rfn
