// Split-fp16 tensor engine ("fp16x3", engine mode 4) and its single-pass bf16 sibling (engine mode 5).
//
// fp32-grade products from kind::f16 MMAs: every operand row is scaled by a power of two so that its largest
// magnitude lands in [2^14, 2^15) and split into two fp16 pieces x = (x0 + x1) / s (22 mantissa bits); the
// product is x0.w0 + x1.w0 + x0.w1 -- three fp16 MMAs, i.e. 1.5 TF32-MMA units against 2 for engine mode 3 and
// 3 for 3xTF32 -- accumulated in one fp32 TMEM tile that is drained into round-to-nearest registers every few
// k-blocks exactly as the 3xTF32 engine does.  The split happens in a separate HBM-bound pass
// (split_rows_kernel), so the GEMM kernel has no splitter warps and touches shared memory only through TMA
// and the tensor core: the mode-3 kernel was bound by the shared-memory pipe (512 tensor + 431 LSU + 384 TMA
// wavefronts per k-block against 1024 MMA cycles, profiles/r2_tc2p_score_shipped_ncu.txt).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "rfn_internal.cuh"

namespace rfn {

// a (rows, K) matrix as one or two 16-bit pieces, row pitch ld elements (a multiple of 8), and 1 / row scale
struct H3Operand {
  const void* p0;      // fp16 (split) or bf16 (single pass)
  const void* p1;      // fp16 residual, nullptr in bf16 mode
  int ld;
  const float* inv;    // (rows) reciprocal of the power-of-two row scale; nullptr = 1
};

struct H3Src {
  H3Operand x;         // (M, K)
  H3Operand w;         // (N, K)   x.inv / w.inv of source 0 are used for every source (joint scales)
  int K;
  const float* bias;   // (N) or nullptr
};

struct H3Gemm {
  H3Src src[3];
  int nsrc;
  int bf16;            // 1: single-pass bf16 (p1 unused), 0: split fp16 (three products)
  float* y;
  int ldy;
  int M, N;
  int accumulate;
  // epilogue 1: fused attention score (as TcArgs); epilogue 2: fused vocabulary statistics
  int epi;
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

// bytes of scratch split_rows needs for a (rows, K_0 .. K_{n-1}) operand: pieces + reciprocal scales
size_t h3_split_bytes(int rows, const int* K, int nsrc, bool bf16);
// splits the fp32 sources (joint power-of-two row scale over all of them) into `scratch`; fills out[s] (all share inv)
int h3_split(const float* const* x, const int* ldx, const int* K, int nsrc, int rows, bool bf16, void* scratch,
             H3Operand* out, cudaStream_t st);
// the operands a split buffer written by h3_split holds (same layout computation, no launch)
void h3_view(const void* buf, int rows, const int* K, int nsrc, bool bf16, H3Operand* out);
bool h3_shape_ok(int M, int N);
int gemm_h3(const H3Gemm& a, cudaStream_t st);
// convenience: split x and W of a GemmArgs into scratch (scratch_bytes must cover both) and run the engine
size_t h3_auto_bytes(const GemmArgs& a, bool bf16);
int gemm_h3_auto(const GemmArgs& a, bool bf16, void* scratch, size_t scratch_bytes, cudaStream_t st);

}  // namespace rfn
