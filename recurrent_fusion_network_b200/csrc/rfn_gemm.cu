// GEMM engine selection.  The north star asks for tensor cores "only when batch x beam is large
// enough to be a real dense contraction; otherwise warp-level FMA": small-M problems always take
// the SIMT kernel, large-M problems take the engine chosen with rfn_set_gemm_mode().
#include "rfn_internal.cuh"

namespace rfn {

int gemm(const GemmArgs& a, cudaStream_t st) {
  return gemm_simt(a, st);
}

}  // namespace rfn
