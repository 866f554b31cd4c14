// Library-level entry points: error string, version, device check, launch accounting.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "rfn_internal.cuh"

namespace rfn {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_gemm_mode{4};  // default: split-fp16 tcgen05 engine (fp32-grade; rfn_h3.cuh), 3xTF32 for the small shapes

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
static std::atomic<uint64_t> g_engine_launches[ENG_COUNT];
void count_engine(int engine) {
  if (engine >= 0 && engine < ENG_COUNT) g_engine_launches[engine].fetch_add(1, std::memory_order_relaxed);
}
int gemm_mode() { return g_gemm_mode.load(std::memory_order_relaxed); }
static std::atomic<int> g_splitk_all{0};   // rfn_set_splitk: every GEMM may use the atomically summed split-K route
bool splitk_all() { return g_splitk_all.load(std::memory_order_relaxed) != 0; }
static std::atomic<int> g_pdl{0};   // measured: no gain at 625 or 5000 images (profiles/r2_pdl_625img.json), so off by default
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

// ---- optional per-kernel-class timing with CUDA events on the launching stream -------------------
static std::atomic<int> g_prof_on{0};
static thread_local int g_tag_override = -1;
struct ProfRec { int tag; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
static thread_local int g_conc_hint = 1;
int concurrency_hint() { return g_conc_hint; }
ConcurrencyScope::ConcurrencyScope(int n) : prev(g_conc_hint) { g_conc_hint = n < 1 ? 1 : n; }
ConcurrencyScope::~ConcurrencyScope() { g_conc_hint = prev; }
TagScope::TagScope(int tag) : prev(g_tag_override) { g_tag_override = tag; }
TagScope::~TagScope() { g_tag_override = prev; }
ProfScope::ProfScope(int default_tag, cudaStream_t st, bool fixed_tag) : st_(st), idx_(-1) {
  if (!prof_enabled()) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = (g_tag_override >= 0 && !fixed_tag) ? g_tag_override : default_tag;
  r.e0 = prof_event();
  r.e1 = prof_event();
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
  idx_ = (int)g_prof_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof_recs[idx_].e1, st_);
}
}  // namespace rfn

extern "C" {

const char* rfn_last_error(void) { return rfn::g_err; }
int rfn_version(void) { return 100; }
uint64_t rfn_launch_count(void) { return rfn::g_launches.load(); }
int rfn_engine_num(void) { return rfn::ENG_COUNT; }
const char* rfn_engine_name(int e) {
  static const char* names[rfn::ENG_COUNT] = {"simt_skinny", "simt_tiled", "tcgen05_1cta", "tcgen05_1cta_splitk", "tcgen05_2cta",
                                              "tcgen05_2cta_persistent_store", "tcgen05_2cta_persistent_score",
                                              "tcgen05_2cta_persistent_vocab", "tcgen05_2cta_persistent_fp16x3",
                                              "tcgen05_2cta_persistent_bf16", "persistent_decoder"};
  return (e >= 0 && e < rfn::ENG_COUNT) ? names[e] : "?";
}
int rfn_engine_launch_counts(uint64_t* out, int n) {
  RFN_CHECK_ARG(out && n >= rfn::ENG_COUNT, "rfn_engine_launch_counts: need %d slots", rfn::ENG_COUNT);
  for (int i = 0; i < rfn::ENG_COUNT; ++i) out[i] = rfn::g_engine_launches[i].load();
  return RFN_OK;
}

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
  RFN_CHECK_ARG(mode >= 0 && mode <= 5, "gemm mode %d not in 0..5", mode);
  rfn::g_gemm_mode.store(mode);
  return RFN_OK;
}
int rfn_get_gemm_mode(void) { return rfn::gemm_mode(); }
int rfn_set_splitk(int on) {
  rfn::g_splitk_all.store(on ? 1 : 0);
  return RFN_OK;
}
int rfn_get_splitk(void) { return rfn::splitk_all() ? 1 : 0; }
int rfn_set_pdl(int on) {
  rfn::g_pdl.store(on ? 1 : 0);
  return RFN_OK;
}
int rfn_get_pdl(void) { return rfn::pdl_enabled() ? 1 : 0; }

int rfn_profile_enable(int on) {
  rfn::g_prof_on.store(on ? 1 : 0);
  return RFN_OK;
}
int rfn_profile_num_tags(void) { return rfn::TAG_COUNT; }
const char* rfn_profile_tag_name(int tag) {
  static const char* names[rfn::TAG_COUNT] = {"misc", "gemm_att2att_stage1", "attention_step_stage1", "gemm_gates",
                                              "gemm_logit", "gemm_other", "attention_step_small", "lstm_cell",
                                              "vocab_stats_select", "beam_merge", "gemm_backward", "attention_backward",
                                              "operand_split"};
  return (tag >= 0 && tag < rfn::TAG_COUNT) ? names[tag] : "?";
}
int rfn_profile_read(float* ms, uint64_t* launches, int n) {
  RFN_CHECK_ARG(ms && launches && n >= rfn::TAG_COUNT, "rfn_profile_read: need %d slots", rfn::TAG_COUNT);
  for (int i = 0; i < n; ++i) { ms[i] = 0.f; launches[i] = 0; }
  std::lock_guard<std::mutex> lk(rfn::g_prof_mu);
  for (auto& r : rfn::g_prof_recs) {
    RFN_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    RFN_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.tag] += t;
    launches[r.tag] += 1;
    rfn::g_prof_pool.push_back(r.e0);
    rfn::g_prof_pool.push_back(r.e1);
  }
  rfn::g_prof_recs.clear();
  return RFN_OK;
}

int rfn_num_params(const rfn_dims* d) {
  if (!d) return RFN_ERR_INVALID;
  const int J = d->J;
  return 2 * J + 3 + 10 * d->num_review_steps_0 * J + 2 * J + d->num_review_steps * (2 + 8 * J) + 2 + 12;
}
}
