// Library-level entry points: error string, version, device check, launch accounting.
#include <stdarg.h>
#include <atomic>

#include "rfn_internal.cuh"

namespace rfn {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_gemm_mode{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int gemm_mode() { return g_gemm_mode.load(std::memory_order_relaxed); }
}  // namespace rfn

extern "C" {

const char* rfn_last_error(void) { return rfn::g_err; }
int rfn_version(void) { return 100; }
uint64_t rfn_launch_count(void) { return rfn::g_launches.load(); }

int rfn_check_device(void) {
  int dev = 0;
  RFN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  RFN_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    rfn::set_error("device %d is sm_%d%d; librfn_b200 is built for sm_100a (B200) only", dev, p.major,
                   p.minor);
    return RFN_ERR_UNSUPPORTED;
  }
  return RFN_OK;
}

int rfn_set_gemm_mode(int mode) {
  RFN_CHECK_ARG(mode >= 0 && mode <= 2, "gemm mode %d not in {0,1,2}", mode);
  rfn::g_gemm_mode.store(mode);
  return RFN_OK;
}
int rfn_get_gemm_mode(void) { return rfn::gemm_mode(); }

int rfn_num_params(const rfn_dims* d) {
  if (!d) return RFN_ERR_INVALID;
  const int J = d->J;
  return 2 * J + 3 + 10 * d->num_review_steps_0 * J + 2 * J + d->num_review_steps * (2 + 8 * J) + 2 + 12;
}
}
