// Internal declarations shared by the translation units of librfn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "rfn_b200.h"

namespace rfn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// per-engine launch accounting (rfn_engine_launch_counts): which GEMM kernel family actually ran
enum Engine {
  ENG_SIMT_SKINNY = 0, ENG_SIMT_TILED = 1, ENG_TC1 = 2, ENG_TC1_SPLITK = 3, ENG_TC2 = 4, ENG_TC2P_STORE = 5,
  ENG_TC2P_SCORE = 6, ENG_TC2P_VOCAB = 7, ENG_H3 = 8, ENG_BF16 = 9, ENG_PERSIST_DECODER = 10,
  ENG_COUNT = 11
};
void count_engine(int engine);
int gemm_mode();
bool splitk_all();
// tensor-engine pass count of an engine mode: 1 -> 3 (3xTF32), 2 -> 1 (single-pass TF32), 3 -> 2 (TF32 hi.hi + two BF16 cross
// terms in the persistent kernel; 3xTF32 wherever another kernel runs)
// modes 4 (split fp16, rfn_h3.cuh) and 5 (single-pass bf16) have their own engine; GEMMs too small for it fall back to
// 3xTF32 (mode 4) / single-pass TF32 (mode 5)
inline int tc_passes(int mode) { return (mode == 1 || mode == 4) ? 3 : (mode == 3 ? 2 : 1); }

// kernel classes for the optional CUDA-event profile (rfn_profile_*)
enum Tag {
  TAG_MISC = 0, TAG_GEMM_ATT2ATT = 1, TAG_ATTN_S1 = 2, TAG_GEMM_GATES = 3, TAG_GEMM_LOGIT = 4, TAG_GEMM_OTHER = 5,
  TAG_ATTN_SMALL = 6, TAG_CELL = 7, TAG_VOCAB = 8, TAG_BEAM = 9, TAG_GEMM_BWD = 10, TAG_ATTN_BWD = 11, TAG_SPLIT = 12,
  TAG_COUNT = 13
};
bool prof_enabled();
// how many independent kernel chains the caller is issuing side by side (the J encoder streams of a fusion step): a GEMM
// launch that would take only a few waves then sizes its grid for its share of the SMs, so that two such launches run
// concurrently instead of one after the other with a partly empty last wave each
int concurrency_hint();
struct ConcurrencyScope {
  int prev;
  explicit ConcurrencyScope(int n);
  ~ConcurrencyScope();
};
struct TagScope {  // call-site override of a launcher's default class
  int prev;
  explicit TagScope(int tag);
  ~TagScope();
};
struct ProfScope {  // records start/stop events around one launcher call when profiling is on
  cudaStream_t st_;
  int idx_;
  ProfScope(int default_tag, cudaStream_t st, bool fixed_tag = false);   // fixed_tag: ignore a call-site TagScope
  ~ProfScope();
};

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl may be scheduled while its predecessor in the
// stream is still running; it must execute pdl_wait() before its first global-memory access (then the predecessor has
// completed and its writes are visible), and every kernel calls pdl_trigger() at its start so that a PDL successor can be
// scheduled early.  What this hides is the launch latency and the prologue (barrier init, TMEM allocation, cluster sync) of
// the ~700 dependent kernels of a decode call -- it matters for small shards.  rfn_set_pdl(0) turns it off.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define RFN_CHECK_ARG(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      rfn::set_error(__VA_ARGS__);            \
      return RFN_ERR_INVALID;                 \
    }                                         \
  } while (0)

#define RFN_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      rfn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return RFN_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define RFN_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    rfn::count_launch();                                                            \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      rfn::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return RFN_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define RFN_TRY(call)            \
  do {                           \
    int s__ = (call);            \
    if (s__ != RFN_OK) return s__; \
  } while (0)

// ---- GEMM (y = sum_i x_i W_i^T + bias) -------------------------------------------------------
struct GemmSrc {
  const float* x;
  const float* w;
  const float* bias;
  int ldx;
  int ldw;
  int K;
};
struct GemmArgs {
  GemmSrc src[3];
  int nsrc;
  float* y;
  int ldy;
  int M;
  int N;
  int accumulate;
  int splitk_ok;   // the caller accepts a split-K sum (atomic adds: summation order not reproducible); weight gradients only
};
int gemm_simt(const GemmArgs& a, cudaStream_t st);
// tcgen05 engine: passes 3 = 3xTF32 (fp32-equivalent), 1 = single-pass TF32; score != nullptr selects
// the fused attention-score epilogue score[m] += sum_n wv[n] tanh(acc + bias + g[m / natt, n])
bool gemm_tc_supported(const GemmArgs& a);
int gemm_tc(const GemmArgs& a, int passes, const float* g, int ldg, const float* wv, float* score, int natt,
            cudaStream_t st);
int gemm_engine(const GemmArgs& a, int engine, cudaStream_t st);
// 1-CTA tensor kernel with the contraction split over CTAs (atomic partial sums); b_mn: W stored (K, N) row-major
int gemm_tc_splitk(const GemmArgs& a, bool b_mn, int passes, cudaStream_t st);
int tc_score_slices(int N);
// logits GEMM whose epilogue keeps only per-(slice,row) max / sum-exp / top-k (slices = tc_score_slices(N))
int gemm_tc_vocab(const GemmArgs& a, int passes, float* st_max, float* st_sum, float* st_val, int32_t* st_idx, int ktop,
                  cudaStream_t st);
// combines the slices: rowmax, log-sum-exp and the top-k LOG-PROBS with ties -> lower index
int vocab_merge(const float* st_max, const float* st_sum, const float* st_val, const int32_t* st_idx, int slices, int rows,
                int k, float* rowmax, float* logsum, float* top_val, int32_t* top_idx, cudaStream_t st);
// dispatches on rfn_set_gemm_mode() and problem shape
int gemm(const GemmArgs& a, cudaStream_t st);
inline GemmArgs gemm1(const float* x, int ldx, const float* w, const float* bias, int K, float* y,
                      int ldy, int M, int N) {
  GemmArgs a{};
  a.src[0] = GemmSrc{x, w, bias, ldx, K, K};
  a.nsrc = 1;
  a.y = y;
  a.ldy = ldy;
  a.M = M;
  a.N = N;
  a.accumulate = 0;
  return a;
}

// ---- attention / pointwise --------------------------------------------------------------------
int attention_step(const float* A, const float* P, const float* g, const float* w, const float* d_wb,
                   float* z, int ldz, float* alpha, int rows, int N, int D, int Ah, int div,
                   cudaStream_t st);
// same, from pre-reduced scores e[r,n] (without the att_h_2_out bias) produced by the fused GEMM epilogue
// A_bf16 != nullptr: the context sum reads this bf16 copy of A (row pitch lda_bf16 elements) instead of the fp32 map
int attention_from_scores(const float* A, const float* scores, int nslices, const float* d_wb, float* z, int ldz,
                          float* alpha, int rows, int N, int D, int div, cudaStream_t st, const void* A_bf16 = nullptr,
                          int lda_bf16 = 0);
int lstm_cell(const float* G, const float* c_prev, float* h_out, float* c_out, float* h_out2,
              int ldh2, float* h_out3, int ldh3, int rows, int R, cudaStream_t st, const float* mask = nullptr,
              float scale = 1.f, int maxout = 0);
int embed_gather_i64(const int64_t* tok, int ld_tok, const float* embed, float* x, int rows, int E,
                     int V1, cudaStream_t st);
int embed_gather_i32(const int32_t* tok, const float* embed, float* x, int rows, int E, int V1,
                     cudaStream_t st);
// dst[r,:] = src[idx ? idx[r] : r / div, :]
int gather_rows(const float* src, const int32_t* idx, int div, float* dst, int rows, int R,
                cudaStream_t st);
// out = (((0 + in_0) + in_1) + ...) / n  over n tensors spaced by `stride` floats
int mean_tensors(const float* in, size_t stride, int n, float* out, int ld_out, size_t count, int R,
                 int ld_in, cudaStream_t st);
// out[r,k] = max_s in[r,s,k]
int max_over_steps(const float* in, float* out, int rows, int S, int K, cudaStream_t st);

// ---- vocab-side kernels -----------------------------------------------------------------------
// per row: max, log(sum exp(x-max)), and the top-k log-probs (ties -> lower index)
int vocab_stats_topk(const float* logits, int ld, int rows, int V, int k, float* rowmax,
                     float* logsum, float* top_val, int32_t* top_idx, cudaStream_t st);
int vocab_write_lp(const float* logits, int ld, const float* rowmax, const float* logsum, float* lp,
                   size_t ld_out, int rows, int V, cudaStream_t st);
// logits_out = (((0 + l_0) + l_1) ...) / n   (eval_utils.py:282-287)
struct PtrList8 {
  const float* p[8];
};
int mean_logits8(const PtrList8& ptrs, int n, float* out, size_t count, cudaStream_t st);

}  // namespace rfn
