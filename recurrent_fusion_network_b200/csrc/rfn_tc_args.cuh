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

}  // namespace rfn
