// Split-fp16 ("fp16x3") / single-pass bf16 tcgen05 GEMM engine: persistent 2-CTA clusters, all operands by TMA.
//   y[M,N] = sum_i x_i[M,K_i] . W_i[N,K_i]^T + bias     (see rfn_h3.cuh for the numerics)
//
// Roles per CTA (384 threads = 3 warpgroups; setmaxnreg moves registers from warpgroup 0 (56 each) to the drain warpgroups (224)):
//   warp 0      TMA producer: per 64-element k-block the two pieces of this CTA's 128 rows of x and of its 128 rows of W
//               (4 x 16 KB, SWIZZLE_128B) into a 3-stage ring; runs ahead through the tiles
//   warp 1      leader CTA: MMA issuer, 4 x {x0.w0} + 4 x {x1.w0, x0.w1} tcgen05.mma.cta_group::2.kind::f16 per k-block into
//               one of two TMEM accumulators (alternating every CH k-blocks, across tiles).  Both CTAs' TMA loads count
//               their bytes on the LEADER's full barrier (cp.async.bulk.tensor.cta_group::2), so the issuer waits on one
//               barrier per stage; the peer's warp 1 is idle (4-CTA clusters keep the relay of the first design)
//   warps 4..11 drain (tcgen05.ld of finished chunks into round-to-nearest fp32 registers) + epilogue of the tile
//               (scaled store / fused attention score / fused vocabulary statistics) while the MMA warp works ahead
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <atomic>

#include "rfn_h3.cuh"
#include "rfn_tc_ptx.cuh"
#include "rfn_tc_epilogue.cuh"

namespace rfn {

constexpr int H3_THREADS = 384;
constexpr int H3_BN = 256;
constexpr int H3_BH = H3_BN / 2;
constexpr int H3_BK = 64;                       // 16-bit elements per 128-byte swizzled row
constexpr int H3_TILE = TC_BM * 128;            // 16 KB: 128 rows x 64 halves
constexpr int H3_GSLOTS = 5;
constexpr int H3_EPI_BYTES = 8 * H3_GSLOTS * 128 * 4;   // 20 KB: g rows (score) / 8 warps x 32 rows x 16 floats (store)

struct H3Args {
  CUtensorMap tm_x[3][2];
  CUtensorMap tm_w[3][2];
  CUtensorMap tm_w64[3][2];   // the same W pieces with a 64-row box (4-CTA clusters: each pair loads half a W tile and multicasts it)
  int K[3];
  const float* bias[3];
  int nsrc;
  const float* row_inv;
  const float* col_inv;
  float* y;
  int ldy;
  int M, N;
  int accumulate;
  const float* g;
  int ldg;
  const float* wv;
  float* score;
  int natt;
  float* st_max;
  float* st_sum;
  float* st_val;
  int32_t* st_idx;
  int ktop;
};

// cluster-scope variants: the peer's relay releases to / the leader's MMA warp acquires from the other CTA
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// TMA load delivered to the same shared-memory offset (and mbarrier offset) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// TMA load issued by either CTA of a pair whose completion bytes are counted on the LEADER's mbarrier (`bar` = the barrier's
// shared-memory address with the peer bit cleared): the leader's MMA warp then waits on ONE barrier per stage and no thread
// has to relay "my half has landed" across the pair (that relay was an mbarrier.arrive.release.cluster per k-block, which
// compiles to MEMBAR.ALL.GPU + ERRBAR, with a CCTL.IVALL on the acquiring side: ~1.2 k cycles per k-block, the bound of the
// single-pass bf16 kernel, profiles/r2_h3_bf16_score_ncu.txt)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mask(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// instruction descriptor, kind::f16: D fp32, A/B fp16 (format 0) or bf16 (format 1), both K-major
__host__ __device__ constexpr uint32_t make_idesc_16(int bm, int bn, int bf16) {
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(bm >> 4) << 24);
}

template <int NPROD>
struct H3Cfg {
  static constexpr int NP = (NPROD == 3) ? 2 : 1;            // pieces per operand
  static constexpr int STAGE_BYTES = 2 * NP * H3_TILE;       // x pieces + W pieces of this CTA
  static constexpr int STAGES = (NPROD == 3) ? 3 : 6;
  static constexpr int CH = (NPROD == 3) ? 4 : 8;            // k-blocks per accumulator chunk (48 / 32 MMAs)
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + H3_EPI_BYTES + 4096 + 1024;
};

// CL = 2: one CTA pair per cluster.  CL = 4: two pairs that work on vertically adjacent 256-row tiles of the SAME 256 output
// columns and share the W tile: each CTA loads half of its 128 W rows and multicasts it to its counterpart in the other pair
// (L2 -> SM traffic per CTA and k-block: x 128 rows + W 64 rows instead of 128 + 128).  The stage is released only when both
// pairs' MMAs have retired (the empty barriers count two commits, multicast to all four CTAs).
template <int EPI, int NPROD, int CL>
__global__ void __launch_bounds__(H3_THREADS, 1) gemm_h3_kernel(const __grid_constant__ H3Args a, int n_tiles, int total_tiles,
                                                                int bf16) {
  using Cfg = H3Cfg<NPROD>;
  constexpr int STAGES = Cfg::STAGES, NP = Cfg::NP, CH = Cfg::CH;
  constexpr int COLS = H3_BN / 2;
  constexpr bool DIRECT = (CL == 2);   // both CTAs' TMA bytes land on the leader's full barrier (no relay warp)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(epi_smem + H3_EPI_BYTES);
  uint64_t* full = bars;                   // DIRECT: leader only, both CTAs' TMA landed; else local: this CTA's TMA landed
  uint64_t* pfull = bars + STAGES;         // !DIRECT: leader: the peer's TMA landed (relayed)
  uint64_t* empty = bars + 2 * STAGES;     // both: stage free (tcgen05.commit multicast)
  uint64_t* cfull = bars + 3 * STAGES;     // [2] both: accumulator chunk complete
  uint64_t* drained = cfull + 2;           // [2] leader: both halves drained
  uint32_t* tmem_slot = (uint32_t*)(drained + 2);
  float* s_bias = (float*)((uint8_t*)bars + 512);
  float* s_wv = s_bias + 256;
  float* s_cs = s_wv + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;                 // rank inside the CTA pair
  const uint32_t pr = (CL == 4) ? (crank >> 1) : 0u;  // which pair of the cluster
  const uint32_t lead = pr * 2u;                    // cluster rank of this pair's leader CTA
  const bool leader = (rank == 0);
  const int cluster = blockIdx.x / CL;
  const int n_clusters = gridDim.x / CL;
  constexpr int MP = CL / 2;                        // 256-row tiles per cluster step
  const uint16_t pair_mask = (uint16_t)(3u << (2u * pr));
  pdl_trigger();

  int total_kb = 0;
  for (int s = 0; s < a.nsrc; ++s) total_kb += (a.K[s] + H3_BK - 1) / H3_BK;
  const int nchunk = (total_kb + CH - 1) / CH;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nsrc; ++s)
      for (int p = 0; p < NP; ++p) { tma_prefetch_desc(&a.tm_x[s][p]); tma_prefetch_desc(&a.tm_w[s][p]); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&pfull[s]), 1);
      mbar_init(smem_u32(&empty[s]), MP);   // one commit per pair of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&cfull[b]), 1);
      mbar_init(smem_u32(&drained[b]), 16);   // 8 drain warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched only shared memory, TMEM and the kernel parameters

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
        const int n0 = (tile % n_tiles) * H3_BN + (int)rank * H3_BH;
        const int m0 = ((tile / n_tiles) * MP + (int)pr) * (2 * TC_BM) + (int)rank * TC_BM;
        for (int s = 0; s < a.nsrc; ++s) {
          const int nkb = (a.K[s] + H3_BK - 1) / H3_BK;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(smem_u32(&empty[st]), ph ^ 1u);
            const uint32_t stage = smem_u32(smem + st * Cfg::STAGE_BYTES);
            if (DIRECT) {
              const uint32_t fb = smem_u32(&full[st]) & PEER_BIT_MASK;   // the leader's barrier, from either CTA
              if (leader) mbar_arrive_expect_tx(fb, (uint32_t)(2 * Cfg::STAGE_BYTES));
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                tma_load_2d_pair(&a.tm_x[s][p], fb, stage + p * H3_TILE, kb * H3_BK, m0);
                tma_load_2d_pair(&a.tm_w[s][p], fb, stage + (NP + p) * H3_TILE, kb * H3_BK, n0);
              }
            } else {
              const uint32_t fb = smem_u32(&full[st]);
              mbar_arrive_expect_tx(fb, (uint32_t)Cfg::STAGE_BYTES);
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                tma_load_2d(&a.tm_x[s][p], fb, stage + p * H3_TILE, kb * H3_BK, m0);
                // my half (64 rows) of this CTA's 128 W rows, to me and to the same-rank CTA of the other pair
                tma_load_2d_mc(&a.tm_w64[s][p], fb, stage + (NP + p) * H3_TILE + pr * (H3_TILE / 2), kb * H3_BK, n0 + (int)pr * (H3_BH / 2),
                               (uint16_t)((1u << rank) | (1u << (rank + 2u))));
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_idesc_16(2 * TC_BM, H3_BN, bf16);
      int it = 0, gc = 0;
      for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          const bool chunk_start = (kb % CH == 0);
          const bool chunk_end = (kb % CH == CH - 1) || (kb == total_kb - 1);
          const int b = gc & 1;
          if (chunk_start && gc >= 2) mbar_wait(smem_u32(&drained[b]), (uint32_t)((gc >> 1) - 1) & 1u);
          mbar_wait(smem_u32(&full[st]), ph);
          if (!DIRECT) mbar_wait_cluster(smem_u32(&pfull[st]), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t td = tmem_base + (uint32_t)(b * H3_BN);
            const uint32_t sa = smem_u32(smem + st * Cfg::STAGE_BYTES);
            const uint64_t dx0 = make_desc_sw128(sa);
            const uint64_t dw0 = make_desc_sw128(sa + NP * H3_TILE);
            // main product first, then the two corrections (2^-11 of it): one 16-element k-step = 32 bytes = 2 descriptor units
#pragma unroll
            for (int k = 0; k < H3_BK / 16; ++k)
              umma2_bf16(td, dx0 + (uint64_t)(k * 2), dw0 + (uint64_t)(k * 2), idesc, (chunk_start && k == 0) ? 0u : 1u);
            if (NPROD == 3) {
              const uint64_t dx1 = make_desc_sw128(sa + H3_TILE);
              const uint64_t dw1 = make_desc_sw128(sa + (NP + 1) * H3_TILE);
#pragma unroll
              for (int k = 0; k < H3_BK / 16; ++k) {
                umma2_bf16(td, dx1 + (uint64_t)(k * 2), dw0 + (uint64_t)(k * 2), idesc, 1u);
                umma2_bf16(td, dx0 + (uint64_t)(k * 2), dw1 + (uint64_t)(k * 2), idesc, 1u);
              }
            }
            umma2_commit_mask(smem_u32(&empty[st]), CL == 4 ? (uint16_t)0xF : (uint16_t)0x3);
            if (chunk_end) umma2_commit_mask(smem_u32(&cfull[b]), pair_mask);
          }
          __syncwarp();
          if (chunk_end) ++gc;
        }
      }
    } else if (!DIRECT) {
      // ===================== relay (4-CTA clusters only): the peer's half of the stage has landed =====================
      int it = 0;
      for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(smem_u32(&full[st]), ph);
          if (lane == 0) mbar_arrive_remote_release(smem_u32(&pfull[st]), lead);
          __syncwarp();
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== drain + epilogue (warps 4..11, 256 threads) =====================
    const int wq = warp & 3;                   // TMEM lane quarter this warp may access
    const int ew = warp - 4;                   // 0..7
    const int half = ew >> 2;
    const int et = threadIdx.x - 128;          // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * COLS);
    float acc[COLS];
    int gc = 0;
    for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
      const int n0 = (tile % n_tiles) * H3_BN;
      const int m0 = ((tile / n_tiles) * MP + (int)pr) * (2 * TC_BM) + (int)rank * TC_BM;
      const int m = m0 + wq * 32 + lane;
      const float rs = (a.row_inv && m < a.M) ? __ldg(a.row_inv + m) : 1.f;
#pragma unroll
      for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunk; ++c, ++gc) {
        const int b = gc & 1;
        mbar_wait(smem_u32(&cfull[b]), (uint32_t)(gc >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)(b * H3_BN + c0), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c0 + i] += v[i];   // round-to-nearest fp32 add
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(smem_u32(&drained[b])); else mbar_arrive_remote(smem_u32(&drained[b]), lead);
        }
      }
      // ----- epilogue of this tile (the MMA warp is already working on the next one) -----
      asm volatile("bar.sync 2, 256;" ::: "memory");     // previous tile's readers of s_bias / s_wv / s_cs are done
      if (et < H3_BN) {
        const int n = n0 + et;
        float bsum = 0.f;
        if (n < a.N)
          for (int s = 0; s < a.nsrc; ++s)
            if (a.bias[s]) bsum += __ldg(a.bias[s] + n);
        s_bias[et] = bsum;
        s_cs[et] = (n < a.N && a.col_inv) ? __ldg(a.col_inv + n) : 1.f;
        if (EPI == 1) s_wv[et] = (n < a.N) ? __ldg(a.wv + n) : 0.f;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const int nb = n0 + half * COLS;
      const float* cb = s_bias + half * COLS;
      const float* cc = s_cs + half * COLS;
      if (EPI == 0) {
        // coalesced store, 16 columns at a time through this warp's 32 x 16 staging block (XOR-swizzled 16-byte chunks:
        // both the one-row-per-lane writes and the four-lanes-per-row reads are bank-conflict free)
        float* stage = reinterpret_cast<float*>(epi_smem) + (size_t)ew * 32 * 16;
        const int wsw = (lane >> 1) & 3;
#pragma unroll
        for (int r16 = 0; r16 < COLS / 16; ++r16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c0 = r16 * 16 + q * 4;
            const float4 bb = *reinterpret_cast<const float4*>(cb + c0);
            const float4 ss = *reinterpret_cast<const float4*>(cc + c0);
            *reinterpret_cast<float4*>(stage + lane * 16 + ((q ^ wsw) * 4)) =
                make_float4(fmaf(acc[c0] * rs, ss.x, bb.x), fmaf(acc[c0 + 1] * rs, ss.y, bb.y), fmaf(acc[c0 + 2] * rs, ss.z, bb.z),
                            fmaf(acc[c0 + 3] * rs, ss.w, bb.w));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2), seg = lane & 3;
            const int mm = m0 + wq * 32 + row;
            const int n = nb + r16 * 16 + seg * 4;
            if (mm < a.M && n + 3 < a.N) {
              float4 o = *reinterpret_cast<const float4*>(stage + row * 16 + ((seg ^ ((row >> 1) & 3)) * 4));
              float* yp = a.y + (size_t)mm * a.ldy + n;
              if (a.accumulate) {
                const float4 t4 = *reinterpret_cast<const float4*>(yp);
                o.x += t4.x; o.y += t4.y; o.z += t4.z; o.w += t4.w;
              }
              *reinterpret_cast<float4*>(yp) = o;
            }
          }
          __syncwarp();
        }
      } else if (EPI == 1) {
        // fused additive-attention score (misc/AttentionModelCore.py:37-42): partial over this thread's 128 columns
        const int mg = (m < a.M ? m : a.M - 1) / a.natt;
        const int mg_first = __shfl_sync(0xffffffffu, mg, 0);
        const int nslots = __shfl_sync(0xffffffffu, mg, 31) - mg_first + 1;
        float* gst = reinterpret_cast<float*>(epi_smem) + (size_t)ew * H3_GSLOTS * COLS;
        const bool staged = nslots <= H3_GSLOTS;
        if (staged) {
          for (int sl = 0; sl < nslots; ++sl) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nb + lane * 4 + 3 < a.N) v = *reinterpret_cast<const float4*>(a.g + (size_t)(mg_first + sl) * a.ldg + nb + lane * 4);
            *reinterpret_cast<float4*>(gst + sl * COLS + lane * 4) = v;
          }
          __syncwarp();
        }
        // (a generic pointer: shared-memory slots when the warp's rows span few images, else straight from global memory)
        const float* gr = staged ? gst + (mg - mg_first) * COLS : a.g + (size_t)mg * a.ldg + nb;
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < COLS / 4; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(cb + q * 4);
          const float4 s4 = *reinterpret_cast<const float4*>(cc + q * 4);
          float4 gg;
          if (staged) gg = *reinterpret_cast<const float4*>(gr + q * 4);
          else gg = (nb + q * 4 + 3 < a.N) ? *reinterpret_cast<const float4*>(gr + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 ww = *reinterpret_cast<const float4*>(s_wv + half * COLS + q * 4);
          part = fmaf(ww.x, tc_tanh(fmaf(acc[q * 4 + 0] * rs, s4.x, b4.x) + gg.x), part);
          part = fmaf(ww.y, tc_tanh(fmaf(acc[q * 4 + 1] * rs, s4.y, b4.y) + gg.y), part);
          part = fmaf(ww.z, tc_tanh(fmaf(acc[q * 4 + 2] * rs, s4.z, b4.z) + gg.z), part);
          part = fmaf(ww.w, tc_tanh(fmaf(acc[q * 4 + 3] * rs, s4.w, b4.w) + gg.w), part);
        }
        const int slice = (n0 / H3_BN) * 2 + half;
        if (m < a.M) a.score[(size_t)slice * a.M + m] = part;
        __syncwarp();
      } else {
        // fused vocabulary epilogue: per (column slice, row) max, sum exp(x - max) and top-k of x = acc + bias
        float mx = -INFINITY;
        if (n0 + H3_BN > a.N) {   // last column tile: columns past N are masked
#pragma unroll
          for (int i = 0; i < COLS; ++i) {
            const float v = fmaf(acc[i] * rs, cc[i], cb[i]);
            acc[i] = (nb + i < a.N) ? v : -INFINITY;
            mx = fmaxf(mx, acc[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < COLS; ++i) {
            acc[i] = fmaf(acc[i] * rs, cc[i], cb[i]);
            mx = fmaxf(mx, acc[i]);
          }
        }
        const float mref = (mx == -INFINITY) ? 0.f : mx;
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < COLS; ++i) se += expf(acc[i] - mref);   // (ex2.approx instead was measured: 0.3 ms of a 142 ms step)
        const int slice = (n0 / H3_BN) * 2 + half;
        if (m < a.M) {
          a.st_max[(size_t)slice * a.M + m] = mx;
          a.st_sum[(size_t)slice * a.M + m] = se;
        }
        // top-k in the order (value descending, column ascending).  k <= 3 (greedy, beam 2 / 3): ONE pass that keeps the best
        // three in registers -- the columns are visited in ascending order and a later column replaces an earlier one only
        // if it is strictly larger, which is that order (13 instructions per column against 8 per column and rank for the
        // k-pass selection below; with the K = 512 of the logit projection this epilogue, not the MMAs, bounds the kernel)
        if (a.ktop <= 3) {
          float t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY;
          int i0 = -1, i1 = -1, i2 = -1;
#pragma unroll
          for (int i = 0; i < COLS; ++i) {
            const float v = acc[i];
            const bool p0 = v > t0, p1 = v > t1, p2 = v > t2;
            t2 = p1 ? t1 : (p2 ? v : t2);
            i2 = p1 ? i1 : (p2 ? i : i2);
            t1 = p0 ? t0 : (p1 ? v : t1);
            i1 = p0 ? i0 : (p1 ? i : i1);
            t0 = p0 ? v : t0;
            i0 = p0 ? i : i0;
          }
          if (m < a.M) {
            float* ov = a.st_val + ((size_t)slice * a.M + m) * a.ktop;
            int32_t* oi = a.st_idx + ((size_t)slice * a.M + m) * a.ktop;
            ov[0] = t0; oi[0] = i0 < 0 ? 0x7fffffff : nb + i0;   // (index 0x7fffffff: fewer than k columns in this slice)
            if (a.ktop > 1) { ov[1] = t1; oi[1] = i1 < 0 ? 0x7fffffff : nb + i1; }
            if (a.ktop > 2) { ov[2] = t2; oi[2] = i2 < 0 ? 0x7fffffff : nb + i2; }
          }
        } else {
          float pv = INFINITY;
          int pi = -1;
          for (int r = 0; r < a.ktop; ++r) {
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
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- operand split: fp32 rows -> power-of-two row scale + two fp16 pieces (or one bf16 piece) --------------------------
struct SplitArgs {
  const float* x[3];
  int ldx[3];
  int K[3];
  void* p0[3];
  void* p1[3];
  int ldp[3];
  int nsrc;
  int rows;
  float* inv;
  int bf16;
};

constexpr int SPLIT_WARPS = 8;

__global__ void __launch_bounds__(SPLIT_WARPS * 32) split_rows_kernel(const SplitArgs a) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * SPLIT_WARPS + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();
  if (row >= a.rows) return;
  float scale = 1.f;
  if (!a.bf16) {
    float mx = 0.f;
    for (int s = 0; s < a.nsrc; ++s) {
      const float4* xr = reinterpret_cast<const float4*>(a.x[s] + (size_t)row * a.ldx[s]);
      for (int i = lane; i < a.K[s] / 4; i += 32) {
        const float4 v = __ldg(xr + i);
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // largest magnitude -> [2^14, 2^15): a power of two, so the scaling itself is exact.  Rows that are all zero (or
    // denormal / non-finite) keep scale 1.
    int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;
    float inv = 1.f;
    if (e > -100 && e < 100) {
      scale = __uint_as_float((uint32_t)(127 + 14 - e) << 23);
      inv = __uint_as_float((uint32_t)(127 - 14 + e) << 23);
    }
    if (lane == 0 && a.inv) a.inv[row] = inv;
  } else if (lane == 0 && a.inv) {
    a.inv[row] = 1.f;
  }
  for (int s = 0; s < a.nsrc; ++s) {
    const float4* xr = reinterpret_cast<const float4*>(a.x[s] + (size_t)row * a.ldx[s]);
    if (a.bf16) {
      uint2* o0 = reinterpret_cast<uint2*>((__nv_bfloat16*)a.p0[s] + (size_t)row * a.ldp[s]);
      for (int i = lane; i < a.K[s] / 4; i += 32) {
        const float4 v = __ldg(xr + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        o0[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    } else {
      uint2* o0 = reinterpret_cast<uint2*>((__half*)a.p0[s] + (size_t)row * a.ldp[s]);
      uint2* o1 = reinterpret_cast<uint2*>((__half*)a.p1[s] + (size_t)row * a.ldp[s]);
      for (int i = lane; i < a.K[s] / 4; i += 32) {
        const float4 v = __ldg(xr + i);
        const float f[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
        __half h0[4], h1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h0[e] = __float2half_rn(f[e]);
          h1[e] = __float2half_rn(f[e] - __half2float(h0[e]));   // exact fp32 subtraction, then 11 more bits
        }
        o0[i] = make_uint2((uint32_t)__half_as_ushort(h0[0]) | ((uint32_t)__half_as_ushort(h0[1]) << 16),
                           (uint32_t)__half_as_ushort(h0[2]) | ((uint32_t)__half_as_ushort(h0[3]) << 16));
        o1[i] = make_uint2((uint32_t)__half_as_ushort(h1[0]) | ((uint32_t)__half_as_ushort(h1[1]) << 16),
                           (uint32_t)__half_as_ushort(h1[2]) | ((uint32_t)__half_as_ushort(h1[3]) << 16));
      }
    }
  }
}

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int ld16(int K) { return (K + 7) & ~7; }

size_t h3_split_bytes(int rows, const int* K, int nsrc, bool bf16) {
  size_t n = up256((size_t)rows * sizeof(float));
  for (int s = 0; s < nsrc; ++s) n += (bf16 ? 1 : 2) * up256((size_t)rows * ld16(K[s]) * 2);
  return n;
}

int h3_split(const float* const* x, const int* ldx, const int* K, int nsrc, int rows, bool bf16, void* scratch, H3Operand* out,
             cudaStream_t st) {
  ProfScope prof__(TAG_SPLIT, st, true);
  RFN_CHECK_ARG(nsrc >= 1 && nsrc <= 3 && rows >= 0 && scratch, "h3_split: bad arguments");
  if (rows == 0) return RFN_OK;
  SplitArgs a{};
  char* p = (char*)scratch;
  a.inv = (float*)p;
  p += up256((size_t)rows * sizeof(float));
  for (int s = 0; s < nsrc; ++s) {
    RFN_CHECK_ARG(K[s] % 4 == 0 && ldx[s] % 4 == 0 && ((uintptr_t)x[s] % 16) == 0, "h3_split: source %d must be 16-byte aligned, K %% 4 == 0", s);
    a.x[s] = x[s]; a.ldx[s] = ldx[s]; a.K[s] = K[s]; a.ldp[s] = ld16(K[s]);
    a.p0[s] = p;
    p += up256((size_t)rows * a.ldp[s] * 2);
    a.p1[s] = nullptr;
    if (!bf16) {
      a.p1[s] = p;
      p += up256((size_t)rows * a.ldp[s] * 2);
    }
    out[s] = H3Operand{a.p0[s], a.p1[s], a.ldp[s], a.inv};
  }
  a.nsrc = nsrc; a.rows = rows; a.bf16 = bf16 ? 1 : 0;
  RFN_CUDA(launch_pdl(split_rows_kernel, dim3((unsigned)((rows + SPLIT_WARPS - 1) / SPLIT_WARPS)), dim3(SPLIT_WARPS * 32), 0, st, a));
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn16)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn16 get_encode16() {
  static EncodeTiledFn16 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn16)p;
  }
  return fn;
}

// 2-D 16-bit tensor map over a row-major (rows, K) matrix with pitch ld elements: box = 64 elements (128 bytes) x 128 rows
static int make_map16(CUtensorMap* tm, const void* base, int rows, int K, int ld, bool bf16, int box_rows = TC_BM) {
  EncodeTiledFn16 enc = get_encode16();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RFN_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)H3_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (16-bit) failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
    return RFN_ERR_CUDA;
  }
  return RFN_OK;
}

bool h3_shape_ok(int M, int N) { return M >= 256 && N >= 256; }

// operand layout of a split buffer: [inv (rows floats)] [piece 0 of source 0] [piece 1 of source 0] [piece 0 of source 1] ...
void h3_view(const void* buf, int rows, const int* K, int nsrc, bool bf16, H3Operand* out) {
  const char* p = (const char*)buf;
  const float* inv = (const float*)p;
  p += up256((size_t)rows * sizeof(float));
  for (int s = 0; s < nsrc; ++s) {
    const int ld = ld16(K[s]);
    const void* p0 = p;
    p += up256((size_t)rows * ld * 2);
    const void* p1 = nullptr;
    if (!bf16) {
      p1 = p;
      p += up256((size_t)rows * ld * 2);
    }
    out[s] = H3Operand{p0, p1, ld, inv};
  }
}

// rfn_set_h3_cluster(4): 4-CTA clusters (two pairs sharing the W tile by TMA multicast) for GEMMs with at least two 256-row tiles
static std::atomic<int> g_h3_cluster{2};

template <int EPI, int NPROD, int CL>
static int launch_h3(const H3Args& t, int bf16, cudaStream_t st) {
  using Cfg = H3Cfg<NPROD>;
  static bool configured = false;
  static int n_sm = 0;
  static int max_clusters = 0;
  if (!configured) {
    RFN_CUDA(cudaFuncSetAttribute(gemm_h3_kernel<EPI, NPROD, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    int dev = 0;
    RFN_CUDA(cudaGetDevice(&dev));
    RFN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    max_clusters = n_sm / CL;
    if (CL > 2) {   // how many 4-CTA clusters of this kernel the GPCs can hold at once
      cudaLaunchConfig_t q{};
      q.gridDim = dim3((unsigned)(CL * (n_sm / CL)), 1, 1);
      q.blockDim = dim3(H3_THREADS, 1, 1);
      q.dynamicSmemBytes = Cfg::SMEM;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, gemm_h3_kernel<EPI, NPROD, CL>, &q) == cudaSuccess && nc > 0) max_clusters = nc;
      else cudaGetLastError();
    }
    configured = true;
  }
  const int n_tiles = (t.N + H3_BN - 1) / H3_BN;
  const int n_pairs = ((t.M + 2 * TC_BM - 1) / (2 * TC_BM) + CL / 2 - 1) / (CL / 2);   // cluster steps along M
  const int total = n_tiles * n_pairs;
  // as few clusters as finish in the same number of waves: a launch of 314 tiles takes 5 waves on 74 cluster slots and on 63,
  // and the 22 SMs it then leaves alone run the other encoder streams' kernels meanwhile (matters for small shards)
  int slots = max_clusters;
  // short launches issued side by side with other encoder streams (concurrency_hint() > 1) take half of the cluster slots:
  // 314 tiles are 5 waves on 74 slots but 9 on 37, i.e. 4.5 waves' worth of the machine, with a sibling launch in the other half
  if (concurrency_hint() > 1 && total <= 6 * slots) slots = (slots + 1) / 2;
  const int waves = (total + slots - 1) / slots;
  const int clusters = (total + waves - 1) / waves;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(CL * clusters), 1, 1);
  cfg.blockDim = dim3(H3_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  RFN_CUDA(cudaLaunchKernelEx(&cfg, gemm_h3_kernel<EPI, NPROD, CL>, t, n_tiles, total, bf16));
  RFN_LAUNCH_CHECK();
  count_engine(bf16 ? ENG_BF16 : ENG_H3);
  return RFN_OK;
}

int gemm_h3(const H3Gemm& a, cudaStream_t st) {
  ProfScope prof__(a.epi == 2 ? TAG_GEMM_LOGIT : TAG_GEMM_OTHER, st);
  RFN_CHECK_ARG(a.nsrc >= 1 && a.nsrc <= 3 && h3_shape_ok(a.M, a.N), "gemm_h3: needs 1..3 sources and M, N >= 256 (M=%d N=%d)", a.M, a.N);
  RFN_CHECK_ARG(a.epi != 0 || (a.N % 4 == 0 && a.ldy % 4 == 0 && ((uintptr_t)a.y % 16) == 0), "gemm_h3: y must be 16-byte aligned, N %% 4 == 0");
  RFN_CHECK_ARG(a.epi != 2 || (a.ktop >= 1 && a.st_max && a.st_sum && a.st_val && a.st_idx),
                "gemm_h3: the vocabulary epilogue needs k >= 1 and its four statistics buffers");
  H3Args t{};
  t.nsrc = a.nsrc;
  const int np = a.bf16 ? 1 : 2;
  const bool cl4 = g_h3_cluster.load() == 4 && a.M > 2 * TC_BM;
  for (int s = 0; s < a.nsrc; ++s) {
    const H3Src& g = a.src[s];
    RFN_CHECK_ARG(g.x.p0 && g.w.p0 && (a.bf16 || (g.x.p1 && g.w.p1)) && g.x.ld % 8 == 0 && g.w.ld % 8 == 0, "gemm_h3: bad operand %d", s);
    for (int p = 0; p < np; ++p) {
      RFN_TRY(make_map16(&t.tm_x[s][p], p ? g.x.p1 : g.x.p0, a.M, g.K, g.x.ld, a.bf16));
      RFN_TRY(make_map16(&t.tm_w[s][p], p ? g.w.p1 : g.w.p0, a.N, g.K, g.w.ld, a.bf16));
      if (cl4) RFN_TRY(make_map16(&t.tm_w64[s][p], p ? g.w.p1 : g.w.p0, a.N, g.K, g.w.ld, a.bf16, TC_BM / 2));
    }
    t.K[s] = g.K;
    t.bias[s] = g.bias;
  }
  t.row_inv = a.src[0].x.inv;
  t.col_inv = a.src[0].w.inv;
  t.y = a.y; t.ldy = a.ldy; t.M = a.M; t.N = a.N; t.accumulate = a.accumulate;
  t.g = a.g; t.ldg = a.ldg; t.wv = a.wv; t.score = a.score; t.natt = a.natt > 0 ? a.natt : 1;
  t.st_max = a.st_max; t.st_sum = a.st_sum; t.st_val = a.st_val; t.st_idx = a.st_idx; t.ktop = a.ktop;
  if (cl4) {
    if (a.bf16) {
      if (a.epi == 0) return launch_h3<0, 1, 4>(t, 1, st);
      if (a.epi == 1) return launch_h3<1, 1, 4>(t, 1, st);
      return launch_h3<2, 1, 4>(t, 1, st);
    }
    if (a.epi == 0) return launch_h3<0, 3, 4>(t, 0, st);
    if (a.epi == 1) return launch_h3<1, 3, 4>(t, 0, st);
    return launch_h3<2, 3, 4>(t, 0, st);
  }
  if (a.bf16) {
    if (a.epi == 0) return launch_h3<0, 1, 2>(t, 1, st);
    if (a.epi == 1) return launch_h3<1, 1, 2>(t, 1, st);
    return launch_h3<2, 1, 2>(t, 1, st);
  }
  if (a.epi == 0) return launch_h3<0, 3, 2>(t, 0, st);
  if (a.epi == 1) return launch_h3<1, 3, 2>(t, 0, st);
  return launch_h3<2, 3, 2>(t, 0, st);
}

size_t h3_auto_bytes(const GemmArgs& a, bool bf16) {
  int K[3];
  for (int s = 0; s < a.nsrc; ++s) K[s] = a.src[s].K;
  return h3_split_bytes(a.M, K, a.nsrc, bf16) + h3_split_bytes(a.N, K, a.nsrc, bf16);
}

int gemm_h3_auto(const GemmArgs& a, bool bf16, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  RFN_CHECK_ARG(scratch && scratch_bytes >= h3_auto_bytes(a, bf16), "gemm_h3_auto: scratch %zu < %zu bytes", scratch_bytes, h3_auto_bytes(a, bf16));
  const float* xs[3]; const float* ws[3];
  int ldx[3], ldw[3], K[3];
  for (int s = 0; s < a.nsrc; ++s) {
    xs[s] = a.src[s].x; ws[s] = a.src[s].w; ldx[s] = a.src[s].ldx; ldw[s] = a.src[s].ldw; K[s] = a.src[s].K;
  }
  H3Operand xo[3], wo[3];
  RFN_TRY(h3_split(xs, ldx, K, a.nsrc, a.M, bf16, scratch, xo, st));
  RFN_TRY(h3_split(ws, ldw, K, a.nsrc, a.N, bf16, (char*)scratch + h3_split_bytes(a.M, K, a.nsrc, bf16), wo, st));
  H3Gemm g{};
  g.nsrc = a.nsrc; g.bf16 = bf16 ? 1 : 0;
  for (int s = 0; s < a.nsrc; ++s) g.src[s] = H3Src{xo[s], wo[s], K[s], a.src[s].bias};
  g.y = a.y; g.ldy = a.ldy; g.M = a.M; g.N = a.N; g.accumulate = a.accumulate;
  return gemm_h3(g, st);
}

}  // namespace rfn

extern "C" int rfn_set_h3_cluster(int ctas) {
  RFN_CHECK_ARG(ctas == 2 || ctas == 4, "rfn_set_h3_cluster: 2 or 4 CTAs per cluster");
  rfn::g_h3_cluster.store(ctas);
  return RFN_OK;
}
extern "C" int rfn_get_h3_cluster(void) { return rfn::g_h3_cluster.load(); }

extern "C" size_t rfn_split_bytes(int rows, int n_src, const int* K, int bf16) {
  if (rows < 0 || n_src < 1 || n_src > 3 || !K) return 0;
  return rfn::h3_split_bytes(rows, K, n_src, bf16 != 0);
}

extern "C" int rfn_split_rows_f32(int n_src, const float* const* x, const int* ldx, const int* K, int rows, int bf16, void* out,
                                  size_t out_bytes, rfn_stream_t stream) {
  RFN_CHECK_ARG(n_src >= 1 && n_src <= 3 && x && ldx && K && out, "rfn_split_rows_f32: bad arguments");
  RFN_CHECK_ARG(out_bytes >= rfn::h3_split_bytes(rows, K, n_src, bf16 != 0), "rfn_split_rows_f32: out buffer %zu < %zu bytes", out_bytes,
                rfn::h3_split_bytes(rows, K, n_src, bf16 != 0));
  rfn::H3Operand ops[3];
  return rfn::h3_split(x, ldx, K, n_src, rows, bf16 != 0, out, ops, (cudaStream_t)stream);
}

extern "C" int rfn_linear_split(int bf16, int n_src, const void* x_split, const void* w_split, const int* K, const float* const* bias,
                                float* y, int ldy, int M, int N, int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(n_src >= 1 && n_src <= 3 && x_split && w_split && K && y, "rfn_linear_split: bad arguments");
  rfn::H3Operand xo[3], wo[3];
  rfn::h3_view(x_split, M, K, n_src, bf16 != 0, xo);
  rfn::h3_view(w_split, N, K, n_src, bf16 != 0, wo);
  rfn::H3Gemm g{};
  g.nsrc = n_src; g.bf16 = bf16 ? 1 : 0;
  for (int s = 0; s < n_src; ++s) g.src[s] = rfn::H3Src{xo[s], wo[s], K[s], bias ? bias[s] : nullptr};
  g.y = y; g.ldy = ldy; g.M = M; g.N = N; g.accumulate = accumulate & 1;
  return rfn::gemm_h3(g, (cudaStream_t)stream);
}
