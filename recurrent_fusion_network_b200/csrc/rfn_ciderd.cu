// CIDEr-D reward scorer on the device (reference: cider/pyciderevalcap/ciderD/ciderD_scorer.py:114-199
// driven by get_rewards.py:39-112).  Captions are integer token sequences of at most 32 tokens (the tokens up to and
// including the first 0, get_rewards.py:20-27); n-grams (n = 1..4) are compared token-wise, document frequencies come
// from an open-addressing hash table built on the host.  All arithmetic is fp64 and every sum runs in the order the
// reference's dict iteration produces (first occurrence of each n-gram), so scores agree to the last few ulps.
#include "rfn_internal.cuh"

namespace rfn {

constexpr int CD_MAXLEN = 32;
constexpr int CD_EMPTY = (int)0x80000000;

__host__ __device__ inline unsigned long long cd_hash(const int* k) {
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < 4; ++i) {
    h ^= (unsigned long long)(unsigned int)k[i];
    h *= 1099511628211ull;
  }
  return h;
}

__device__ __forceinline__ double cd_df(const int* __restrict__ keys, const double* __restrict__ vals, int cap, const int* tok, int n) {
  if (cap == 0) return 0.0;
  int k[4] = {-1, -1, -1, -1};
  for (int i = 0; i < n; ++i) k[i] = tok[i];
  unsigned long long s = cd_hash(k) & (unsigned long long)(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    const int* e = keys + 4 * s;
    if (e[0] == CD_EMPTY) return 0.0;
    if (e[0] == k[0] && e[1] == k[1] && e[2] == k[2] && e[3] == k[3]) return vals[s];
    s = (s + 1) & (unsigned long long)(cap - 1);
  }
  return 0.0;
}

__device__ __forceinline__ bool cd_same(const int* a, const int* b, int n) {
  for (int i = 0; i < n; ++i)
    if (a[i] != b[i]) return false;
  return true;
}

__device__ __forceinline__ int cd_caption_len(const int* t, int L) {
  for (int i = 0; i < L; ++i)
    if (t[i] == 0) return i + 1;   // the terminating 0 is part of the caption string
  return L;
}

// one thread per (hypothesis, reference slot): sim(vec_hyp, vec_ref, ...) for the four n-gram orders
__global__ void ciderd_pair_kernel(const int* __restrict__ hyp, int ld_h, int n_hyp, const int* __restrict__ hyp_img,
                                   const int* __restrict__ refs, const int* __restrict__ n_refs, int R, int Lr,
                                   const int* __restrict__ df_keys, const double* __restrict__ df_vals, int cap,
                                   double ref_len, double sigma, double* __restrict__ sims) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_hyp * R) return;
  const int hi = idx / R, ri = idx % R;
  const int img = hyp_img[hi];
  double* out = sims + (size_t)idx * 4;
  if (ri >= n_refs[img]) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
  int h[CD_MAXLEN], r[CD_MAXLEN];
  const int lh = cd_caption_len(hyp + (size_t)hi * ld_h, min(ld_h, CD_MAXLEN));
  const int lr = cd_caption_len(refs + ((size_t)img * R + ri) * Lr, min(Lr, CD_MAXLEN));
  for (int i = 0; i < lh; ++i) h[i] = hyp[(size_t)hi * ld_h + i];
  for (int i = 0; i < lr; ++i) r[i] = refs[((size_t)img * R + ri) * Lr + i];
  // "length" = sum of bigram term frequencies (ciderD_scorer.py:138-139)
  const double delta = (double)(max(lh - 1, 0) - max(lr - 1, 0));
  const double pen = pow(2.718281828459045, -(delta * delta) / (2.0 * sigma * sigma));
  for (int n = 1; n <= 4; ++n) {
    const int nh = lh - n + 1, nr = lr - n + 1;
    double norm_h = 0.0, norm_r = 0.0, val = 0.0;
    for (int i = 0; i < nh; ++i) {
      bool first = true;
      for (int j = 0; j < i; ++j)
        if (cd_same(h + i, h + j, n)) { first = false; break; }
      if (!first) continue;
      int tf = 1;
      for (int j = i + 1; j < nh; ++j) tf += cd_same(h + i, h + j, n) ? 1 : 0;
      const double d = log(fmax(1.0, cd_df(df_keys, df_vals, cap, h + i, n)));
      const double vh = (double)tf * (ref_len - d);
      norm_h += vh * vh;
      int tfr = 0;
      for (int j = 0; j < nr; ++j) tfr += cd_same(h + i, r + j, n) ? 1 : 0;
      const double vr = tfr ? (double)tfr * (ref_len - d) : 0.0;
      val += fmin(vh, vr) * vr;
    }
    for (int i = 0; i < nr; ++i) {
      bool first = true;
      for (int j = 0; j < i; ++j)
        if (cd_same(r + i, r + j, n)) { first = false; break; }
      if (!first) continue;
      int tf = 1;
      for (int j = i + 1; j < nr; ++j) tf += cd_same(r + i, r + j, n) ? 1 : 0;
      const double d = log(fmax(1.0, cd_df(df_keys, df_vals, cap, r + i, n)));
      const double vr = (double)tf * (ref_len - d);
      norm_r += vr * vr;
    }
    norm_h = sqrt(norm_h);
    norm_r = sqrt(norm_r);
    if (norm_h != 0.0 && norm_r != 0.0) val /= (norm_h * norm_r);
    out[n - 1] = val * pen;
  }
}

// score = mean_n(sum_refs sim) / n_refs * 10; reward[b, :] = w * (score[b] - score[rows + b])   (get_rewards.py:96-110)
__global__ void ciderd_finalize_kernel(const double* __restrict__ sims, const int* __restrict__ hyp_img,
                                       const int* __restrict__ n_refs, int n_hyp, int R, double* __restrict__ scores) {
  const int hi = blockIdx.x * blockDim.x + threadIdx.x;
  if (hi >= n_hyp) return;
  const int nr = n_refs[hyp_img[hi]];
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int ri = 0; ri < nr; ++ri)
    for (int n = 0; n < 4; ++n) s[n] += sims[((size_t)hi * R + ri) * 4 + n];
  double m = ((s[0] + s[1]) + s[2]) + s[3];
  m /= 4.0;
  m /= (double)nr;
  m *= 10.0;
  scores[hi] = m;
}

__global__ void ciderd_reward_kernel(const double* __restrict__ scores, int rows, int T, double weight, int use_baseline,
                                     float* __restrict__ reward) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * T) return;
  const int b = i / T;
  const double v = use_baseline ? scores[b] - scores[rows + b] : scores[b];
  reward[i] = (float)(v * weight);
}

}  // namespace rfn

using namespace rfn;
extern "C" {

int rfn_ciderd_scores_f64(const int32_t* hyp, int ld_h, int n_hyp, const int32_t* hyp_img, const int32_t* refs,
                          const int32_t* n_refs, int R, int Lr, const int32_t* df_keys, const double* df_vals, int df_cap,
                          double ref_len, double sigma, double* sims_scratch, double* scores, rfn_stream_t stream) {
  RFN_CHECK_ARG(hyp && hyp_img && refs && n_refs && sims_scratch && scores, "rfn_ciderd_scores_f64: null pointer");
  RFN_CHECK_ARG(ld_h >= 1 && ld_h <= CD_MAXLEN && Lr >= 1 && Lr <= CD_MAXLEN && R >= 1, "rfn_ciderd_scores_f64: captions longer than %d tokens", CD_MAXLEN);
  RFN_CHECK_ARG(df_cap == 0 || ((df_cap & (df_cap - 1)) == 0 && df_keys && df_vals), "rfn_ciderd_scores_f64: df table capacity must be a power of two");
  if (n_hyp == 0) return RFN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof__(TAG_MISC, st);
    ciderd_pair_kernel<<<(n_hyp * R + 63) / 64, 64, 0, st>>>(hyp, ld_h, n_hyp, hyp_img, refs, n_refs, R, Lr, df_keys, df_vals, df_cap,
                                                             ref_len, sigma, sims_scratch);
    RFN_LAUNCH_CHECK();
  }
  ProfScope prof__(TAG_MISC, st);
  ciderd_finalize_kernel<<<(n_hyp + 63) / 64, 64, 0, st>>>(sims_scratch, hyp_img, n_refs, n_hyp, R, scores);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

int rfn_ciderd_reward_f32(const double* scores, int rows, int T, double weight, int use_baseline, float* reward,
                          rfn_stream_t stream) {
  RFN_CHECK_ARG(scores && reward, "rfn_ciderd_reward_f32: null pointer");
  if (rows * T == 0) return RFN_OK;
  ciderd_reward_kernel<<<(rows * T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scores, rows, T, weight, use_baseline, reward);
  RFN_LAUNCH_CHECK();
  return RFN_OK;
}

// host helper: slot of an n-gram key (4 ints, unused = -1) in a table of `cap` (power of two) entries, before probing
uint64_t rfn_ciderd_hash(const int32_t* key4) { return (uint64_t)cd_hash(key4); }

}  // extern "C"
