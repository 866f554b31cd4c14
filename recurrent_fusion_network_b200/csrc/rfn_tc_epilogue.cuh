// Epilogues shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.  Each worker warp owns a 32-row x COLS-column
// block of the output tile, one row per lane, COLS fp32 values in registers.
#pragma once
#include "rfn_tc_args.cuh"

namespace rfn {

// y = acc + bias (+ y): staged through shared memory (the operand ring is dead once the last accumulator chunk
// has been drained) so that every global store instruction writes 512 contiguous bytes of one output row.
// Writing straight from the one-row-per-lane register layout costs ~30k cycles per tile (32 scattered
// 16-byte pieces per instruction); the staged version a tenth of that.
// The worker warps stage the tile's bias (summed over the sources; 0 beyond N) and, for the score epilogue, the
// att_h_2_out weights into shared memory during the prologue: per-thread global loads in the epilogue formed a
// serialized long-scoreboard chain (~10k cycles per tile, profiles/r1_epi0 source view).
template <int BN>
__device__ __forceinline__ void tc_stage_bias(const TcArgs& a, int n0, int t /* 0..255 */, float* s_bias, float* s_wv,
                                              bool with_bias = true) {
  if (t < BN) {
    const int n = n0 + t;
    float b = 0.f;
    if (n < a.N && with_bias)
      for (int s = 0; s < a.nsrc; ++s)
        if (a.bias[s]) b += __ldg(a.bias[s] + n);
    s_bias[t] = b;
    if (s_wv) s_wv[t] = (n < a.N) ? __ldg(a.wv + n) : 0.f;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight worker warps only
}

// tanh for the fused score epilogue: 1 - 2 / (exp(2x) + 1) on the SFU (ex2.approx + rcp.approx), absolute error
// <= ~2e-7 over the whole range (|tanh| <= 1), against ~35 instructions for tanhf: the epilogue evaluates 128 of
// them per thread and was compute-bound on tanhf (14k of 120k cycles per tile).
__device__ __forceinline__ float tc_tanh(float x) {
  const float e = exp2f(x * 2.8853900817779268f);       // exp(2x); saturates cleanly to 0 / inf
  return 1.f - __fdividef(2.f, e + 1.f);
}

template <int COLS>
__device__ __forceinline__ void tc_epilogue_store(float (&acc)[COLS], float* stage /* this warp's 32 x (COLS+4) floats */,
                                                  const TcArgs& a, int m_base, int nb, int lane, const float* s_bias_half,
                                                  bool atomic = false) {
  constexpr int LD = COLS + 4;
  // bias is the same for every row: add it on the way into shared memory
#pragma unroll
  for (int q = 0; q < COLS / 4; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(s_bias_half + q * 4);
    *reinterpret_cast<float4*>(stage + lane * LD + q * 4) =
        make_float4(acc[q * 4] + b.x, acc[q * 4 + 1] + b.y, acc[q * 4 + 2] + b.z, acc[q * 4 + 3] + b.w);
  }
  __syncwarp();
  // lanes now sweep the columns of one row at a time: fully coalesced 128-bit stores
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const int m = m_base + r;
    if (m >= a.M) break;
    float* yr = a.y + (size_t)m * a.ldy;
#pragma unroll
    for (int c = lane * 4; c < COLS; c += 128) {
      const int n = nb + c;
      if (n + 3 < a.N) {
        float4 o = *reinterpret_cast<const float4*>(stage + r * LD + c);
        if (atomic) {   // split-K partial tile: one 16-byte vector reduction per lane instead of four scalar atomics
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yr + n), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
          continue;
        }
        if (a.accumulate) {
          const float4 t = *reinterpret_cast<const float4*>(yr + n);
          o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        *reinterpret_cast<float4*>(yr + n) = o;
      }
    }
  }
}

}  // namespace rfn
