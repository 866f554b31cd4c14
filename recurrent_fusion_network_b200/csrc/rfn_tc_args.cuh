// Kernel argument block shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace rfn {

constexpr int TC_THREADS = 320;             // warp 0 TMA, warp 1 MMA, warps 2..9 workers

struct TcArgs {
  CUtensorMap tm_x[3];
  CUtensorMap tm_w[3];
  int K[3];
  const float* bias[3];
  int nsrc;
  float* y;
  int ldy;
  int M, N;
  int accumulate;
  // split-K (2-CTA store kernel only, one source): cluster z of ksplit owns a contiguous range of k-blocks and adds its
  // partial tile with red.global.add (y pre-zeroed or accumulated into); used by the weight-gradient GEMMs whose output
  // has few tiles and whose contraction runs over rows x attention locations
  int ksplit;
  // 1-CTA store kernel only: W is given as (K, N) row-major (contraction index slow), i.e. the B operand is MN-major;
  // dX = dY . W reads nn.Linear weights (out, in) this way without a transposed copy
  int b_mn;
  // persistent kernel only: the two cross terms of the split product run as BF16 MMAs (x_lo . w and x . w_lo with bf16
  // operands written by the splitter) instead of TF32 MMAs: 8 instead of 12 tensor-core units per k-block, relative error
  // ~1.3e-6 instead of ~4e-7 (engine mode 3)
  int bf16x;
  CUtensorMap tm_wb[2];   // bf16x: bf16(W) and bf16(W - trunc_tf32(W)) in the core-matrix-tiled layout written by w_bf16_tiles_kernel
  // fused attention-score epilogue (epi == 1)
  int epi;
  const float* g;   // (rows, ldg)   h_2_att_h(h)
  const float* wv;  // (N)           att_h_2_out.weight
  float* score;     // (N / (BN/2), M) partial scores, one slice per worker column range
  int natt;         // attention locations per feature row: g row = m / natt
  int ldg;
  // fused vocabulary epilogue (epi == 2): per (column slice, row) max, sum exp(x - max) and top-k of x = acc + bias
  float* st_max;    // (slices, M)
  float* st_sum;    // (slices, M)
  float* st_val;    // (slices, M, ktop)
  int32_t* st_idx;  // (slices, M, ktop)
  int ktop;
  // optional per-CTA timeline stamps (clock64): [cta][8] = start, init done, first MMA issue, last MMA issue,
  // last chunk drained, epilogue done, exit
  long long* dbg;
};

// host helpers (rfn_gemm_tc.cu)
// atom32: 128-byte swizzle with 32-byte atoms (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), the layout an MN-major tf32 operand needs
int tc_make_map(CUtensorMap* tm, const float* base, int rows, int K, int ld, int box_rows, bool atom32 = false);

// 2-D bf16 map over the core-matrix-tiled W copy: rows = Ng groups of 8 W rows, Kp * 8 elements each; box = 256 x 16
int tc_make_map_bf16_tiles(CUtensorMap* tm, const void* base, int Ng, int Kp);
int tc2p_prepare_bf16_w(TcArgs& t, const float* W, int ldw, int N, int K, void** scratch, cudaStream_t st);

}  // namespace rfn
