// GEMM engine selection.  The north star asks for tensor cores "only when batch x beam is large
// enough to be a real dense contraction; otherwise warp-level FMA": problems with fewer than
// 128 rows always take the SIMT kernel; larger ones take the engine chosen with
// rfn_set_gemm_mode() (0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 single-pass TF32).
#include "rfn_internal.cuh"
#include "rfn_h3.cuh"

namespace rfn {

int gemm_engine(const GemmArgs& a, int engine, cudaStream_t st) {
  if (engine == 0) return gemm_simt(a, st);
  if (engine >= 4) {
    // split-fp16 / bf16 engine on explicitly given fp32 operands (engine parity tests): the operand split needs scratch,
    // taken from the stream-ordered pool here; the path code hands the engine its caller-owned workspace instead
    RFN_CHECK_ARG(h3_shape_ok(a.M, a.N) && gemm_tc_supported(a), "engine %d needs M, N >= 256 and 16-byte aligned operands", engine);
    const size_t bytes = h3_auto_bytes(a, engine == 5);
    void* scratch = nullptr;
    RFN_CUDA(cudaMallocAsync(&scratch, bytes, st));
    const int rc = gemm_h3_auto(a, engine == 5, scratch, bytes, st);
    cudaFreeAsync(scratch, st);
    return rc;
  }
  return gemm_tc(a, tc_passes(engine), nullptr, 0, nullptr, nullptr, 0, st);
}

int gemm(const GemmArgs& a, cudaStream_t st) {
  const int mode = gemm_mode();
  // plain-store GEMMs of mode 3 stay on the 3xTF32 kernels (3-stage ring); its BF16 cross terms live in the persistent
  // kernel, which the fused-epilogue GEMMs of the path use (and rfn_linear_f32_engine(3, ...) for the parity tests)
  // (modes 4 / 5: the path code routes large GEMMs to the split-fp16 / bf16 engine with its workspace, rfn_path.cu; what
  // arrives here are the shapes that engine does not take: 3xTF32 / single-pass TF32 respectively)
  const int eng = (mode == 3 || mode == 4) ? 1 : (mode == 5 ? 2 : mode);
  const bool splitk_ok = a.splitk_ok || splitk_all();
  if (mode >= 1 && a.M >= 128 && splitk_ok && gemm_tc_supported(a)) {
    // a few hundred rows (the 250 sampled captions of an RL iteration): the 128 x 128 output tiles alone cover a fraction of
    // the SMs and each would walk the whole contraction; split it over CTAs when the caller accepts an atomic sum
    long nkb = 0;
    for (int s = 0; s < a.nsrc; ++s) nkb += (a.src[s].K + 31) / 32;
    const long tiles = (long)((a.M + 127) / 128) * ((a.N + 127) / 128);
    if (tiles * 2 <= 148 && nkb >= 8) return gemm_tc_splitk(a, false, tc_passes(mode), st);
  }
  if (mode >= 1 && a.M >= 128 && gemm_tc_supported(a)) return gemm_engine(a, eng, st);
  if (mode >= 1 && splitk_ok && gemm_tc_supported(a)) {   // training: few rows, the weights are streamed once
    long wk = 0;
    for (int s = 0; s < a.nsrc; ++s) wk += a.src[s].K;
    if (wk * a.N >= (a.M > 32 ? (1L << 17) : (1L << 20))) return gemm_tc_splitk(a, false, tc_passes(mode), st);
  }
  return gemm_simt(a, st);
}

}  // namespace rfn

extern "C" int rfn_linear_f32_engine(int engine, int n_src, const float* const* x, const int* ldx, const float* const* W,
                                     const int* K, const float* const* bias, float* y, int ldy, int M, int N,
                                     int accumulate, rfn_stream_t stream) {
  RFN_CHECK_ARG(engine >= 0 && engine <= 5, "rfn_linear_f32_engine: engine %d not in 0..5", engine);
  RFN_CHECK_ARG(n_src >= 1 && n_src <= 3 && x && ldx && W && K, "rfn_linear_f32_engine: bad source arrays");
  rfn::GemmArgs a{};
  a.nsrc = n_src;
  for (int s = 0; s < n_src; ++s) a.src[s] = rfn::GemmSrc{x[s], W[s], bias ? bias[s] : nullptr, ldx[s], K[s], K[s]};
  a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.accumulate = accumulate & 1; a.splitk_ok = (accumulate & RFN_GEMM_SPLITK) ? 1 : 0;
  return rfn::gemm_engine(a, engine, (cudaStream_t)stream);
}
