// Persistent 2-CTA tcgen05 GEMM (3xTF32 mode): one cluster per SM pair loops over output tiles so that a
// tile's prologue (barrier init, TMEM allocation, pipeline fill) and epilogue overlap the next tile's
// mainloop.  The per-tile timeline of the one-tile-per-cluster kernel (rfn_gemm_tc2.cu) showed 15 % of an
// att_2_att_h tile and ~60 % of a K = 512 logits tile outside the tensor-bound mainloop.
//
// Roles per CTA (512 threads = 4 warpgroups; setmaxnreg moves registers from warpgroups 0 and 3 (72 each) to the drain
// warpgroups 1 and 2 (184 each)):
//   warp 0        TMA producer: runs ahead through the tiles, bounded only by the stage ring
//   warp 1        MMA issuer (leader CTA): alternates the two TMEM accumulators per chunk ACROSS tiles, so it can
//                 be two chunks (8 k-blocks) into the next tile while the previous tile's epilogue runs
//   warps 2..3    splitters (3xTF32: lo = x - trunc_tf32(x); mode 3: bf16(x), bf16(x_lo)), never blocked by an epilogue
//   warps 4..11   drain + epilogue: round-to-nearest accumulation of the finished chunks in registers, then the
//                 store / attention-score / vocabulary epilogue of the tile
//   warps 12..15  four more splitter warps in engine mode 3 (idle in 3xTF32 mode), where the x split bounds the kernel
// Tiles are assigned round-robin (tile = cluster, cluster + #clusters, ...) with the n-tile index fastest, so
// every role derives the same sequence without communication.
#include <cuda.h>
#include <algorithm>
#include <cuda_bf16.h>

#include "rfn_internal.cuh"
#include "rfn_tc_args.cuh"
#include "rfn_tc_ptx.cuh"
#include "rfn_tc_epilogue.cuh"

namespace rfn {

constexpr int TP_THREADS = 512;   // 4 warpgroups: {TMA, MMA, 2 splitters}, 2 x drain, {4 more splitters (mode 3)}
constexpr int TP_BN = 256;
constexpr int TP_BH = TP_BN / 2;
constexpr int TP_TILE_BYTES = TC_A_BYTES + TP_BH * 128;   // 32 KB landed per CTA per k-block
constexpr int TP_STAGE_BYTES = 2 * TP_TILE_BYTES;         // raw (= hi) + lo
constexpr int TP_CH = 4;                                   // k-blocks per accumulator chunk
constexpr int TP_GSLOTS = 5;                               // g rows staged per epilogue warp (score epilogue)

template <int EPI>
struct TpCfg {
  static constexpr int STAGES = (EPI == 0) ? 2 : 3;
  // epilogue staging: store epilogue = 8 warps x 32 rows x (64 + 4) floats; score epilogue = 8 x GSLOTS x 128 floats
  static constexpr int EPI_BYTES = (EPI == 0) ? 8 * 32 * 68 * 4 : 8 * TP_GSLOTS * 128 * 4;
  static constexpr size_t SMEM = (size_t)STAGES * TP_STAGE_BYTES + EPI_BYTES + 4096 + 1024;
};

template <int EPI>
__global__ void __launch_bounds__(TP_THREADS, 1) gemm_tc2p_kernel(const __grid_constant__ TcArgs a, int n_tiles, int total_tiles) {
  using Cfg = TpCfg<EPI>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int COLS = TP_BN / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = smem + STAGES * TP_STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(epi_smem + Cfg::EPI_BYTES);
  uint64_t* full = bars;                 // local: TMA landed
  uint64_t* ready = bars + STAGES;       // leader: both halves landed and split
  uint64_t* empty = bars + 2 * STAGES;   // both: stage free
  uint64_t* cfull = bars + 3 * STAGES;   // [2] both: accumulator chunk complete
  uint64_t* drained = cfull + 2;         // [2] leader: both halves drained
  uint32_t* tmem_slot = (uint32_t*)(drained + 2);
  float* s_bias = (float*)((uint8_t*)bars + 512);
  float* s_wv = s_bias + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  int total_kb = 0;
  for (int s = 0; s < a.nsrc; ++s) total_kb += (a.K[s] + TC_BK - 1) / TC_BK;
  const int nchunk = (total_kb + TP_CH - 1) / TP_CH;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nsrc; ++s) { tma_prefetch_desc(&a.tm_x[s]); tma_prefetch_desc(&a.tm_w[s]); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&ready[s]), a.bf16x ? 12 : 4);   // splitter warps x 2 CTAs (mode 3: warps 2, 3 and 12..15)
      mbar_init(smem_u32(&empty[s]), 1);
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

  // 512 threads start with 128 registers each; the drain / epilogue warpgroups hold a 128-column fp32 tile per thread and
  // take 184, the producer / MMA / splitter warpgroups keep 72 (128 x (72 + 72 + 184 + 184) = 65,536)
  if (warp < 4 || warp >= 12) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
        const int n0 = (tile % n_tiles) * TP_BN;
        const int m0 = (tile / n_tiles) * (2 * TC_BM) + (int)rank * TC_BM;
        for (int s = 0; s < a.nsrc; ++s) {
          const int nkb = (a.K[s] + TC_BK - 1) / TC_BK;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(smem_u32(&empty[st]), ph ^ 1u);
            const uint32_t fb = smem_u32(&full[st]);
            mbar_arrive_expect_tx(fb, (uint32_t)(TP_TILE_BYTES + (a.bf16x ? 16384 : 0)));
            uint8_t* stage = smem + st * TP_STAGE_BYTES;
            tma_load_2d(&a.tm_x[s], fb, smem_u32(stage), kb * TC_BK, m0);
            tma_load_2d(&a.tm_w[s], fb, smem_u32(stage + TC_A_BYTES), kb * TC_BK, n0 + (int)rank * TP_BH);
            if (a.bf16x) {   // this CTA's 128 W rows as 16 groups x 512 bytes of bf16 core matrices, twice (w, w_lo)
              const int grp = (n0 + (int)rank * TP_BH) >> 3;
              tma_load_2d(&a.tm_wb[0], fb, smem_u32(stage + TP_TILE_BYTES + 16384), kb * 256, grp);
              tma_load_2d(&a.tm_wb[1], fb, smem_u32(stage + TP_TILE_BYTES + 24576), kb * 256, grp);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32_m(2 * TC_BM, TP_BN);
      int it = 0, gc = 0;
      for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          const bool chunk_start = (kb % TP_CH == 0);
          const int b = gc & 1;
          if (chunk_start && gc >= 2) mbar_wait(smem_u32(&drained[b]), (uint32_t)((gc >> 1) - 1) & 1u);
          mbar_wait(smem_u32(&ready[st]), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t td = tmem_base + (uint32_t)(b * TP_BN);
            const uint32_t sa = smem_u32(smem + st * TP_STAGE_BYTES);
            const uint64_t da_hi = make_desc_sw128(sa);
            const uint64_t db_hi = make_desc_sw128(sa + TC_A_BYTES);
            if (a.bf16x) {
              // hi.hi in TF32 (raw tile, truncated by the tensor core); x_lo.w and x.w_lo in BF16 from the four 8 KB
              // tiles the splitter wrote: [x | x_lo | w | w_lo]
              constexpr uint32_t idesc_b = make_idesc_bf16_m(2 * TC_BM, TP_BN);
              const uint32_t sb = sa + TP_TILE_BYTES;
              const uint64_t dxb = make_desc_bf16_interleaved(sb), dxl = make_desc_bf16_interleaved(sb + 8192);
              const uint64_t dwb = make_desc_bf16_interleaved(sb + 16384), dwl = make_desc_bf16_interleaved(sb + 24576);
#pragma unroll
              for (int k = 0; k < TC_BK / 8; ++k)
                umma2_tf32(td, da_hi + (uint64_t)(k * 2), db_hi + (uint64_t)(k * 2), idesc, (chunk_start && k == 0) ? 0u : 1u);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                umma2_bf16(td, dxl + (uint64_t)(k * 16), dwb + (uint64_t)(k * 16), idesc_b, 1u);
                umma2_bf16(td, dxb + (uint64_t)(k * 16), dwl + (uint64_t)(k * 16), idesc_b, 1u);
              }
            } else {
              const uint64_t da_lo = make_desc_sw128(sa + TP_TILE_BYTES);
              const uint64_t db_lo = make_desc_sw128(sa + TP_TILE_BYTES + TC_A_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BK / 8; ++k) {
                const uint64_t adv = (uint64_t)(k * 2);
                umma2_tf32(td, da_hi + adv, db_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
                umma2_tf32(td, da_lo + adv, db_hi + adv, idesc, 1u);
                umma2_tf32(td, da_hi + adv, db_lo + adv, idesc, 1u);
              }
            }
            umma2_commit(smem_u32(&empty[st]));
            if (kb % TP_CH == TP_CH - 1 || kb == total_kb - 1) umma2_commit(smem_u32(&cfull[b]));
          }
          __syncwarp();
          if (kb % TP_CH == TP_CH - 1 || kb == total_kb - 1) ++gc;
        }
      }
    }
  } else {
    // ===================== splitters: warps 2, 3 (both modes) and 12..15 (mode 3 only) =====================
    const int t = (warp < 4) ? (int)threadIdx.x - 64 : 64 + ((int)threadIdx.x - 384);
    int it = 0;
    const int first_tile = (warp >= 12 && !a.bf16x) ? total_tiles : cluster;   // the extra warpgroup idles in 3xTF32 mode
    for (int tile = first_tile; tile < total_tiles; tile += n_clusters) {
      for (int kb = 0; kb < total_kb; ++kb, ++it) {
        const int st = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(smem_u32(&full[st]), ph);
        const uint4* hi = reinterpret_cast<const uint4*>(smem + st * TP_STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + st * TP_STAGE_BYTES + TP_TILE_BYTES);
        if (a.bf16x) {
          // x tile only (the W side arrives as bf16 by TMA).  item = (row r, 8-element k-chunk c): two swizzled 16-byte
          // pieces of the raw fp32 row in, one 16-byte core-matrix row of bf16(v) and one of bf16(v - trunc_tf32(v)) out
          uint8_t* base = smem + st * TP_STAGE_BYTES + TP_TILE_BYTES;
          // rows fastest across the lanes: a quarter-warp reads 8 distinct swizzled chunks and writes one whole 128-byte
          // core matrix (both bank-conflict free)
#pragma unroll 3
          for (int i = t; i < 128 * 4; i += 192) {
            const int r = i & 127, c = i >> 7;
            const uint4* row = hi + r * 8;
            const uint4 v0 = row[(2 * c) ^ (r & 7)], v1 = row[(2 * c + 1) ^ (r & 7)];
            const float f[8] = {__uint_as_float(v0.x), __uint_as_float(v0.y), __uint_as_float(v0.z), __uint_as_float(v0.w),
                                __uint_as_float(v1.x), __uint_as_float(v1.y), __uint_as_float(v1.z), __uint_as_float(v1.w)};
            uint32_t pb[4], pl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a0 = f[2 * e], a1 = f[2 * e + 1];
              const float l0 = a0 - __uint_as_float(__float_as_uint(a0) & 0xffffe000u);
              const float l1 = a1 - __uint_as_float(__float_as_uint(a1) & 0xffffe000u);
              __nv_bfloat162 b2 = __floats2bfloat162_rn(a0, a1), l2 = __floats2bfloat162_rn(l0, l1);
              pb[e] = *reinterpret_cast<uint32_t*>(&b2);
              pl[e] = *reinterpret_cast<uint32_t*>(&l2);
            }
            const int off = (r >> 3) * 512 + c * 128 + (r & 7) * 16;
            *reinterpret_cast<uint4*>(base + off) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
            *reinterpret_cast<uint4*>(base + off + 8192) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          }
        } else
#pragma unroll 8
        for (int i = t; i < TP_TILE_BYTES / 16; i += 64) {
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
        if (lane == 0) {
          if (leader) mbar_arrive(smem_u32(&ready[st])); else mbar_arrive_remote(smem_u32(&ready[st]), 0);
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    // ===================== drain + epilogue (256 threads) =====================
    const int wq = warp & 3;
    const int ew = warp - 4;                   // 0..7
    const int half = ew >> 2;
    const int et = threadIdx.x - 128;          // 0..255
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * COLS);
    float acc[COLS];
    int gc = 0;
    for (int tile = cluster; tile < total_tiles; tile += n_clusters) {
      const int n0 = (tile % n_tiles) * TP_BN;
      const int m0 = (tile / n_tiles) * (2 * TC_BM) + (int)rank * TC_BM;
#pragma unroll
      for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunk; ++c, ++gc) {
        const int b = gc & 1;
        mbar_wait(smem_u32(&cfull[b]), (uint32_t)(gc >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)(b * TP_BN + c0), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c0 + i] += v[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(smem_u32(&drained[b])); else mbar_arrive_remote(smem_u32(&drained[b]), 0);
        }
      }
      // ----- epilogue of this tile (the MMA warp is already working on the next one) -----
      asm volatile("bar.sync 2, 256;" ::: "memory");     // previous tile's readers of s_bias / s_wv are done
      if (et < TP_BN) {
        const int n = n0 + et;
        float bsum = 0.f;
        if (n < a.N)
          for (int s = 0; s < a.nsrc; ++s)
            if (a.bias[s]) bsum += __ldg(a.bias[s] + n);
        s_bias[et] = bsum;
        if (EPI == 1) s_wv[et] = (n < a.N) ? __ldg(a.wv + n) : 0.f;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const int m = m0 + wq * 32 + lane;
      const int nb = n0 + half * COLS;
      if (EPI == 0) {
        // coalesced store, 64 columns at a time through this warp's 32 x 68 staging block
        float* stage = reinterpret_cast<float*>(epi_smem) + (size_t)ew * 32 * 68;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(s_bias + half * COLS + h * 64 + q * 4);
            *reinterpret_cast<float4*>(stage + lane * 68 + q * 4) =
                make_float4(acc[h * 64 + q * 4] + bb.x, acc[h * 64 + q * 4 + 1] + bb.y, acc[h * 64 + q * 4 + 2] + bb.z,
                            acc[h * 64 + q * 4 + 3] + bb.w);
          }
          __syncwarp();
          // two rows per instruction: lanes 0..15 -> row r, lanes 16..31 -> row r + 1; 256 contiguous bytes each
          const int rsel = lane >> 4, cq = (lane & 15) * 4;
#pragma unroll 4
          for (int r = 0; r < 32; r += 2) {
            const int mm = m0 + wq * 32 + r + rsel;
            const int n = nb + h * 64 + cq;
            if (mm < a.M && n + 3 < a.N) {
              float4 o = *reinterpret_cast<const float4*>(stage + (r + rsel) * 68 + cq);
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
        const int mg = (m < a.M ? m : a.M - 1) / a.natt;
        const int mg_first = __shfl_sync(0xffffffffu, mg, 0);
        const int nslots = __shfl_sync(0xffffffffu, mg, 31) - mg_first + 1;
        float* gst = reinterpret_cast<float*>(epi_smem) + (size_t)ew * TP_GSLOTS * COLS;
        const bool staged = nslots <= TP_GSLOTS;
        if (staged) {
          for (int sl = 0; sl < nslots; ++sl) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nb + lane * 4 + 3 < a.N) v = *reinterpret_cast<const float4*>(a.g + (size_t)(mg_first + sl) * a.ldg + nb + lane * 4);
            *reinterpret_cast<float4*>(gst + sl * COLS + lane * 4) = v;
          }
          __syncwarp();
        }
        const float* gr = staged ? gst + (mg - mg_first) * COLS : a.g + (size_t)mg * a.ldg + nb;
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < COLS / 4; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + half * COLS + q * 4);
          float4 gg;
          if (staged) gg = *reinterpret_cast<const float4*>(gr + q * 4);
          else gg = (nb + q * 4 + 3 < a.N) ? *reinterpret_cast<const float4*>(gr + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 ww = *reinterpret_cast<const float4*>(s_wv + half * COLS + q * 4);
          part = fmaf(ww.x, tc_tanh(acc[q * 4 + 0] + b4.x + gg.x), part);
          part = fmaf(ww.y, tc_tanh(acc[q * 4 + 1] + b4.y + gg.y), part);
          part = fmaf(ww.z, tc_tanh(acc[q * 4 + 2] + b4.z + gg.z), part);
          part = fmaf(ww.w, tc_tanh(acc[q * 4 + 3] + b4.w + gg.w), part);
        }
        const int slice = (n0 / TP_BN) * 2 + half;
        if (m < a.M) a.score[(size_t)slice * a.M + m] = part;
        __syncwarp();
      } else {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < COLS; ++i) {
          const float v = acc[i] + s_bias[half * COLS + i];
          acc[i] = (nb + i < a.N) ? v : -INFINITY;
          mx = fmaxf(mx, acc[i]);
        }
        const float mref = (mx == -INFINITY) ? 0.f : mx;
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < COLS; ++i) se += expf(acc[i] - mref);
        const int slice = (n0 / TP_BN) * 2 + half;
        if (m < a.M) {
          a.st_max[(size_t)slice * a.M + m] = mx;
          a.st_sum[(size_t)slice * a.M + m] = se;
        }
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
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// bf16(W) and bf16(W - trunc_tf32(W)) in the layout the bf16 MMA descriptors expect, so that one plain TMA box lands a
// 128-row x 32-element tile as 16 groups x 4 core matrices: element (n, k) at ((n/8) * (Kp/8) + k/8) * 64 + (n%8) * 8 + k%8
__global__ void __launch_bounds__(256) w_bf16_tiles_kernel(const float* __restrict__ W, int ldw, int N, int K, int Ng, int Kp,
                                                           __nv_bfloat16* __restrict__ Wb, __nv_bfloat16* __restrict__ Wl) {
  const long total = (long)Ng * 8 * (Kp / 8);
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int kc = (int)(i % (Kp / 8));
    const int n = (int)(i / (Kp / 8));
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      f[e] = (n < N && k < K) ? __ldg(W + (size_t)n * ldw + k) : 0.f;
    }
    uint32_t pb[4], pl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float l0 = f[2 * e] - __uint_as_float(__float_as_uint(f[2 * e]) & 0xffffe000u);
      const float l1 = f[2 * e + 1] - __uint_as_float(__float_as_uint(f[2 * e + 1]) & 0xffffe000u);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]), l2 = __floats2bfloat162_rn(l0, l1);
      pb[e] = *reinterpret_cast<uint32_t*>(&b2);
      pl[e] = *reinterpret_cast<uint32_t*>(&l2);
    }
    const size_t off = ((size_t)(n >> 3) * (Kp / 8) + kc) * 64 + (size_t)(n & 7) * 8;
    *reinterpret_cast<uint4*>(Wb + off) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
    *reinterpret_cast<uint4*>(Wl + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// prepares the bf16 W tiles of engine mode 3 in stream-ordered scratch memory (freed after the GEMM on the same stream)
int tc2p_prepare_bf16_w(TcArgs& t, const float* W, int ldw, int N, int K, void** scratch, cudaStream_t st) {
  const int Ng = (N + 7) / 8, Kp = (K + TC_BK - 1) / TC_BK * TC_BK;
  const size_t elems = (size_t)Ng * 8 * Kp;
  // keep freed scratch in the device's default stream-ordered pool (its default release threshold of 0 hands the memory
  // back to the driver at every synchronisation, which would make each cudaMallocAsync below a real allocation)
  static bool pool_ready[64] = {};
  int dev = 0;
  RFN_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
    cudaMemPool_t pool;
    RFN_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;
    RFN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_ready[dev] = true;
  }
  void* p = nullptr;
  RFN_CUDA(cudaMallocAsync(&p, 2 * elems * sizeof(__nv_bfloat16), st));
  *scratch = p;
  __nv_bfloat16* Wb = (__nv_bfloat16*)p;
  __nv_bfloat16* Wl = Wb + elems;
  const long items = (long)Ng * 8 * (Kp / 8);
  w_bf16_tiles_kernel<<<(unsigned)std::min<long>((items + 255) / 256, 4096), 256, 0, st>>>(W, ldw, N, K, Ng, Kp, Wb, Wl);
  RFN_LAUNCH_CHECK();
  for (int i = 0; i < 2; ++i)
    RFN_TRY(tc_make_map_bf16_tiles(&t.tm_wb[i], i == 0 ? Wb : Wl, Ng, Kp));
  return RFN_OK;
}

template <int EPI>
static int launch_tc2p_epi(const TcArgs& t, cudaStream_t st) {
  using Cfg = TpCfg<EPI>;
  static bool configured = false;
  static int n_sm = 0;
  if (!configured) {
    RFN_CUDA(cudaFuncSetAttribute(gemm_tc2p_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    int dev = 0;
    RFN_CUDA(cudaGetDevice(&dev));
    RFN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int n_tiles = (t.N + TP_BN - 1) / TP_BN;
  const int n_pairs = (t.M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int total = n_tiles * n_pairs;
  const int clusters = total < n_sm / 2 ? total : n_sm / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters), 1, 1);
  cfg.blockDim = dim3(TP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RFN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2p_kernel<EPI>, t, n_tiles, total));
  RFN_LAUNCH_CHECK();
  count_engine(EPI == 0 ? ENG_TC2P_STORE : (EPI == 1 ? ENG_TC2P_SCORE : ENG_TC2P_VOCAB));
  return RFN_OK;
}

// persistent variant, 3xTF32 only; args prepared by gemm_tc / gemm_tc_vocab (128-row TMA boxes for both operands)
int launch_tc2p(const TcArgs& t, cudaStream_t st) {
  if (t.epi == 0) return launch_tc2p_epi<0>(t, st);
  if (t.epi == 1) return launch_tc2p_epi<1>(t, st);
  return launch_tc2p_epi<2>(t, st);
}

}  // namespace rfn
