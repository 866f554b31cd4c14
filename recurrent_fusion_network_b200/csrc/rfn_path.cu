// Path-level host orchestration: stage 1 (fusion), stage 2 (review), decoder loops (teacher forced,
// greedy / multinomial, batched beam, ensemble beam).  Pure launch sequencing on one stream: no
// host<->device synchronisation, no allocation; everything lives in the caller's workspace.
#include <algorithm>
#include <atomic>
#include <initializer_list>
#include <mutex>
#include <vector>

#include "rfn_decoder_persist.cuh"
#include "rfn_h3.cuh"
#include "rfn_internal.cuh"
#include "rfn_vocab.cuh"

namespace rfn {

// ---- parameter indexing: the reference's state_dict registration order ---------------------------
// (misc/RecurrentFusionModel.py:153-184)
struct PIdx {
  int J, S0, S1;
  explicit PIdx(const rfn_dims& d) : J(d.J), S0(d.num_review_steps_0), S1(d.num_review_steps) {}
  int fc2h(int j, int o) const { return 2 * j + o; }
  int embed() const { return 2 * J; }
  int logit(int o) const { return 2 * J + 1 + o; }
  // o: 0 att_2_att_h.w 1 .b 2 h_2_att_h.w 3 .b 4 att_h_2_out.w 5 .b 6 H2h.w 7 .b 8 z2h.w 9 .b
  int s1(int s, int j, int o) const { return 2 * J + 3 + (s * J + j) * 10 + o; }
  int reason_ind(int j, int o) const { return 2 * J + 3 + 10 * S0 * J + 2 * j + o; }
  int s2base(int s) const { return 2 * J + 3 + 10 * S0 * J + 2 * J + s * (2 + 8 * J); }
  int s2_h2h(int s, int o) const { return s2base(s) + o; }
  int s2_z2h(int s, int j, int o) const { return s2base(s) + 2 + 2 * j + o; }
  int s2_att(int s, int j, int o) const { return s2base(s) + 2 + 2 * J + 6 * j + o; }
  int reason(int o) const { return s2base(S1) + o; }
  // o: 0 i2h.w 1 .b 2 h2h.w 3 .b 4 z2h.w 5 .b 6 att_2_att_h.w 7 .b 8 h_2_att_h.w 9 .b 10 att_h_2_out.w 11 .b
  int dec(int o) const { return s2base(S1) + 2 + o; }
};

// ---- bump allocator over the caller's workspace (dry run computes the requirement) ---------------
struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base((char*)b) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// gate-row widths: 4R, or 5R with the maxout input transform (misc/LSTMSoftAttentionCore.py:25, ...FeatArrayNoInputCore.py:25)
static inline int g2w(const rfn_dims& d) { return (4 + (d.review_maxout ? 1 : 0)) * d.rnn_size; }
static inline int gdw(const rfn_dims& d) { return (4 + (d.decoder_maxout ? 1 : 0)) * d.rnn_size; }

static int check_dims(const rfn_dims* d) {
  RFN_CHECK_ARG(d != nullptr, "dims is null");
  RFN_CHECK_ARG((d->review_maxout == 0 || d->review_maxout == 1) && (d->decoder_maxout == 0 || d->decoder_maxout == 1),
                "review_maxout / decoder_maxout must be 0 or 1");
  RFN_CHECK_ARG(d->J >= 1 && d->J <= RFN_MAX_ENCODERS, "J=%d not in 1..%d", d->J, RFN_MAX_ENCODERS);
  RFN_CHECK_ARG(d->rnn_size % 4 == 0 && d->att_hid_size % 4 == 0 && d->input_encoding_size % 4 == 0,
                "rnn_size/att_hid_size/input_encoding_size must be multiples of 4");
  for (int j = 0; j < d->J; ++j)
    RFN_CHECK_ARG(d->att_feat_size[j] % 4 == 0 && d->fc_feat_size[j] % 4 == 0 && d->att_num[j] >= 1,
                  "encoder %d: feature sizes must be multiples of 4", j);
  RFN_CHECK_ARG(d->num_review_steps_0 >= 1 && d->num_review_steps >= 1 && d->seq_length >= 1 && d->seq_length <= 64,
                "review steps >= 1 and 1 <= seq_length <= 64 required");
  RFN_CHECK_ARG(d->vocab_plus1 >= 2 && d->top_words_count >= 1, "bad vocab / top_words_count");
  return RFN_OK;
}

// ---- side streams: the J encoder cells of a fusion step are independent (Jacobi update), so their
// kernel chains run concurrently, forked from and joined back into the caller's stream with events ----
struct SidePool {
  int dev = -1;
  cudaStream_t s[RFN_MAX_ENCODERS];
  cudaEvent_t fork;
  cudaEvent_t join[RFN_MAX_ENCODERS];
};
static std::atomic<int> g_concurrency{1};
static SidePool* side_pool() {
  static std::mutex mu;
  static SidePool pools[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  SidePool& p = pools[dev];
  if (p.dev != dev) {
    for (int i = 0; i < RFN_MAX_ENCODERS; ++i) {
      if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    p.dev = dev;
  }
  return &p;
}

template <typename F>
static int for_each_encoder(int J, cudaStream_t st, F&& fn) {
  SidePool* sp = (g_concurrency.load() && J > 1) ? side_pool() : nullptr;
  if (!sp) {
    for (int j = 0; j < J; ++j) RFN_TRY(fn(j, st));
    return RFN_OK;
  }
  RFN_CUDA(cudaEventRecord(sp->fork, st));
  ConcurrencyScope conc(J);
  for (int j = 0; j < J; ++j) {
    RFN_CUDA(cudaStreamWaitEvent(sp->s[j], sp->fork, 0));
    RFN_TRY(fn(j, sp->s[j]));
    RFN_CUDA(cudaEventRecord(sp->join[j], sp->s[j]));
  }
  for (int j = 0; j < J; ++j) RFN_CUDA(cudaStreamWaitEvent(st, sp->join[j], 0));
  return RFN_OK;
}

// ---- split-fp16 / bf16 engine plumbing (engine modes 4 / 5, rfn_h3.cuh) -----------------------------
// Large GEMMs (M, N >= 256) run on the split engine: their fp32 operands are split into scaled fp16 pairs in the
// caller's workspace first -- the image features once per call (reused by the S0 fusion steps), the thought vectors
// once for stage 2, the decoder's weights once per decode loop, everything else right before its GEMM.
struct H3Ws {   // per-stream scratch for on-the-fly operand splits
  void* x = nullptr;
  size_t xcap = 0;
  void* w = nullptr;
  size_t wcap = 0;
};
static bool h3_mode() { return gemm_mode() >= 4; }
static bool h3_bf16() { return gemm_mode() == 5; }
static size_t split1(int rows, int K) { return h3_split_bytes(rows, &K, 1, h3_bf16()); }
static size_t split2(int rows, int K0, int K1) { const int K[2] = {K0, K1}; return h3_split_bytes(rows, K, 2, h3_bf16()); }
static size_t split3(int rows, int K0, int K1, int K2) { const int K[3] = {K0, K1, K2}; return h3_split_bytes(rows, K, 3, h3_bf16()); }
static bool h3_take(const GemmArgs& a) { return h3_mode() && h3_shape_ok(a.M, a.N) && gemm_tc_supported(a); }

// pre-split one operand (rows, K) into `buf` when the GEMM that will consume it (its other dimension being `other`) takes the engine
static int h3_presplit(const float* x, int ld, int K, int rows, void* buf, size_t cap, H3Operand* out, cudaStream_t st) {
  if (split1(rows, K) > cap) {
    set_error("operand split: scratch %zu < %zu bytes", cap, split1(rows, K));
    return RFN_ERR_WORKSPACE;
  }
  return h3_split(&x, &ld, &K, 1, rows, h3_bf16(), buf, out, st);
}

// fills g's sources: x from xpre or split into sc.x, W from wpre or split into sc.w
static int h3_prepare(const GemmArgs& a, const H3Ws& sc, const H3Operand* xpre, const H3Operand* wpre, H3Gemm& g, cudaStream_t st) {
  const bool bf = h3_bf16();
  const float *xs[3], *ws[3];
  int K[3], ldx[3], ldw[3];
  for (int s = 0; s < a.nsrc; ++s) {
    xs[s] = a.src[s].x; ws[s] = a.src[s].w; ldx[s] = a.src[s].ldx; ldw[s] = a.src[s].ldw; K[s] = a.src[s].K;
  }
  H3Operand xo[3], wo[3];
  if (xpre) {
    for (int s = 0; s < a.nsrc; ++s) xo[s] = xpre[s];
  } else {
    const size_t need = h3_split_bytes(a.M, K, a.nsrc, bf);
    if (need > sc.xcap) { set_error("operand split (x): scratch %zu < %zu bytes", sc.xcap, need); return RFN_ERR_WORKSPACE; }
    RFN_TRY(h3_split(xs, ldx, K, a.nsrc, a.M, bf, sc.x, xo, st));
  }
  if (wpre) {
    for (int s = 0; s < a.nsrc; ++s) wo[s] = wpre[s];
  } else {
    const size_t need = h3_split_bytes(a.N, K, a.nsrc, bf);
    if (need > sc.wcap) { set_error("operand split (W): scratch %zu < %zu bytes", sc.wcap, need); return RFN_ERR_WORKSPACE; }
    RFN_TRY(h3_split(ws, ldw, K, a.nsrc, a.N, bf, sc.w, wo, st));
  }
  g = H3Gemm{};
  g.nsrc = a.nsrc; g.bf16 = bf ? 1 : 0;
  for (int s = 0; s < a.nsrc; ++s) g.src[s] = H3Src{xo[s], wo[s], K[s], a.src[s].bias};
  g.y = a.y; g.ldy = a.ldy; g.M = a.M; g.N = a.N; g.accumulate = a.accumulate;
  return RFN_OK;
}

// y = sum_i x_i W_i^T + b on the engine the mode selects
static int path_gemm(const GemmArgs& a, const H3Ws* sc, const H3Operand* xpre, const H3Operand* wpre, cudaStream_t st) {
  if (sc && h3_take(a)) {
    H3Gemm g;
    RFN_TRY(h3_prepare(a, *sc, xpre, wpre, g, st));
    return gemm_h3(g, st);
  }
  return gemm(a, st);
}

// ---- weight cache: the split of every weight matrix the engine consumes, made once per weight version --------------
// (inference: the weights do not change between calls; splitting them on the fly costs 2 x 1.96 GB of HBM traffic and ~300
// launches per call, which is what bounds small shards, e.g. 625 images per GPU at 8 GPUs)
struct WGroupDesc {
  int n_rows, nsrc;
  int K[3], pidx[3];
  size_t off;
};
struct WCache {
  const char* base = nullptr;
  bool bf16 = false;
  int J = 0, S0 = 0, S1 = 0, G2 = 0;      // G2: gate-GEMM groups of a stage-2 step (h2h + 2 z_2_h, then 3 z_2_h each)
  std::vector<WGroupDesc> g;
  size_t bytes = 0;

  int i_fc2h(int j) const { return j; }
  int i_s1(int s, int j, int what /* 0 att_2_att_h, 1 h_2_att_h, 2 H2h|z2h */) const { return J + (s * J + j) * 3 + what; }
  int i_reason_ind(int j) const { return J + 3 * S0 * J + j; }
  int i_s2(int s, int k /* 2j: att_2_att_h_j, 2j+1: h_2_att_h_j, 2J+grp: gates group */) const {
    return 2 * J + 3 * S0 * J + s * (2 * J + G2) + k;
  }
  int i_reason() const { return 2 * J + 3 * S0 * J + S1 * (2 * J + G2); }
  int i_dec(int what /* 0 att_2_att_h, 1 h_2_att_h, 2 i2h|h2h|z2h, 3 logit */) const { return i_reason() + 1 + what; }

  void add(int n_rows, std::initializer_list<int> K, std::initializer_list<int> pidx) {
    WGroupDesc w{};
    w.n_rows = n_rows;
    w.nsrc = (int)K.size();
    int i = 0;
    for (int k : K) w.K[i++] = k;
    i = 0;
    for (int q : pidx) w.pidx[i++] = q;
    w.off = bytes;
    bytes += (h3_split_bytes(n_rows, w.K, w.nsrc, bf16) + 255) & ~(size_t)255;
    g.push_back(w);
  }
  WCache(const rfn_dims& d, bool bf16_) : bf16(bf16_), J(d.J), S0(d.num_review_steps_0), S1(d.num_review_steps) {
    const PIdx ix(d);
    const int R = d.rnn_size, A = d.att_hid_size, E = d.input_encoding_size, V = d.vocab_plus1, K = d.top_words_count;
    G2 = 1 + (std::max(0, J - 2) + 2) / 3;
    for (int j = 0; j < J; ++j) add(R, {d.fc_feat_size[j]}, {ix.fc2h(j, 0)});
    for (int s = 0; s < S0; ++s)
      for (int j = 0; j < J; ++j) {
        add(A, {d.att_feat_size[j]}, {ix.s1(s, j, 0)});
        add(A, {R}, {ix.s1(s, j, 2)});
        add(4 * R, {J * R, d.att_feat_size[j]}, {ix.s1(s, j, 6), ix.s1(s, j, 8)});
      }
    for (int j = 0; j < J; ++j) add(K, {R}, {ix.reason_ind(j, 0)});
    for (int s = 0; s < S1; ++s) {
      for (int j = 0; j < J; ++j) {
        add(A, {R}, {ix.s2_att(s, j, 0)});
        add(A, {R}, {ix.s2_att(s, j, 2)});
      }
      int jn = 0;
      for (int grp = 0; grp < G2; ++grp) {   // the grouping of thought_vectors()'s stage-2 gate GEMMs
        WGroupDesc w{};
        w.n_rows = g2w(d);
        int n = 0;
        if (grp == 0) { w.K[n] = R; w.pidx[n++] = ix.s2_h2h(s, 0); }
        while (n < 3 && jn < J) { w.K[n] = R; w.pidx[n++] = ix.s2_z2h(s, jn, 0); ++jn; }
        w.nsrc = n;
        w.off = bytes;
        bytes += (h3_split_bytes(w.n_rows, w.K, w.nsrc, bf16) + 255) & ~(size_t)255;
        g.push_back(w);
      }
    }
    add(K, {R}, {ix.reason(0)});
    add(A, {R}, {ix.dec(6)});
    add(A, {R}, {ix.dec(8)});
    add(gdw(d), {E, R, R}, {ix.dec(0), ix.dec(2), ix.dec(4)});
    add(V, {R}, {ix.logit(0)});
  }
  // the cached operands of group i, or nullptr when there is no cache
  const H3Operand* get(int i, H3Operand* out) const {
    if (!base) return nullptr;
    const WGroupDesc& w = g[i];
    h3_view(base + w.off, w.n_rows, w.K, w.nsrc, bf16, out);
    return out;
  }
};

// ---- stages 1 + 2 ----------------------------------------------------------------------------------
struct TVWork {
  float *Hcat[2], *C, *TV, *hx, *rs;
  float *g[RFN_MAX_ENCODERS], *P[RFN_MAX_ENCODERS], *z[RFN_MAX_ENCODERS], *G[RFN_MAX_ENCODERS];
  size_t p_cap[RFN_MAX_ENCODERS];        // floats available in P[j]
  H3Ws h3[RFN_MAX_ENCODERS];             // engine modes 4 / 5: split scratch of encoder j's stream
  void* fsplit[RFN_MAX_ENCODERS];        // pre-split features of encoder j (stage 1), then its thought vectors (stage 2)
  size_t fsplit_cap[RFN_MAX_ENCODERS];
  void* hsplit;                          // stage 2: the review state h, split once per step for the J query projections
  size_t hsplit_cap;
};
static size_t p_floats(const rfn_dims& d, int rows, int N) {
  // SIMT engine: P = att_2_att_h(A) is materialised (rows*N, A); tensor engine: partial scores only
  const size_t full = (size_t)rows * N * d.att_hid_size;
  if (gemm_mode() == 0) return full;
  return std::max((size_t)tc_score_slices(d.att_hid_size) * rows * N, std::min(full, (size_t)128 * d.att_hid_size));
}
static size_t carve_tv(const rfn_dims& d, int rows, bool need_tv, bool need_reason, Bump& b, TVWork& w) {
  const int J = d.J, R = d.rnn_size, A = d.att_hid_size, S0 = d.num_review_steps_0, S1 = d.num_review_steps;
  w.Hcat[0] = b.take<float>((size_t)rows * J * R);
  w.Hcat[1] = b.take<float>((size_t)rows * J * R);
  w.C = b.take<float>((size_t)rows * J * R);
  w.hx = b.take<float>((size_t)rows * R);
  for (int j = 0; j < J; ++j) {
    w.g[j] = b.take<float>((size_t)rows * A);
    w.p_cap[j] = p_floats(d, rows, std::max(d.att_num[j], S0));
    w.P[j] = b.take<float>(w.p_cap[j]);
    w.z[j] = b.take<float>((size_t)rows * std::max(d.att_feat_size[j], R));
    w.G[j] = b.take<float>((size_t)rows * std::max(4 * R, g2w(d)));
    w.h3[j] = H3Ws{};
    w.fsplit[j] = nullptr;
    w.fsplit_cap[j] = 0;
    if (h3_mode()) {
      const int D = d.att_feat_size[j], F = d.fc_feat_size[j], K = d.top_words_count;
      w.fsplit_cap[j] = std::max(split1(rows * d.att_num[j], D), split1(rows * S0, R));
      w.fsplit[j] = b.take<char>(w.fsplit_cap[j]);
      w.h3[j].xcap = std::max({split2(rows, J * R, D), split1(rows, R), split1(rows, F), split3(rows, R, R, R), split1(rows * S1, R)});
      w.h3[j].x = b.take<char>(w.h3[j].xcap);
      w.h3[j].wcap = std::max({split2(4 * R, J * R, D), split1(A, D), split1(A, R), split1(R, F), split3(g2w(d), R, R, R), split1(K, R)});
      w.h3[j].w = b.take<char>(w.h3[j].wcap);
    }
  }
  w.hsplit = nullptr;
  w.hsplit_cap = 0;
  if (h3_mode()) {
    w.hsplit_cap = split1(rows, R);
    w.hsplit = b.take<char>(w.hsplit_cap);
  }
  w.TV = need_tv ? b.take<float>((size_t)J * rows * S0 * R) : nullptr;
  w.rs = need_reason ? b.take<float>((size_t)rows * std::max(S0, S1) * d.top_words_count) : nullptr;
  return b.off;
}

// z = Att(h, A) for one attention module (misc/AttentionModelCore.py:31-48): g = h_2_att_h(h); then either
// the tensor engine with the fused tanh-score epilogue (U_a A never reaches HBM) or GEMM + fused step kernel
static int attention_module(const rfn_dims& d, const float* h, int ldh, const float* Afeat, int N, int D,
                            const float* U_w, const float* U_b, const float* Wh_w, const float* Wh_b, const float* v_w,
                            const float* v_b, float* g, float* P, size_t p_cap, float* z, int ldz, int rows, int tag_gemm,
                            int tag_attn, const H3Ws* sc, const H3Operand* feat_split, const H3Operand* w_u, const H3Operand* w_h,
                            cudaStream_t st, const H3Operand* h_split = nullptr) {
  const int R = d.rnn_size, A = d.att_hid_size;
  RFN_TRY(path_gemm(gemm1(h, ldh, Wh_w, Wh_b, R, g, A, rows, A), sc, h_split, w_h, st));      // :36
  GemmArgs pa = gemm1(Afeat, D, U_w, U_b, D, P, A, rows * N, A);                            // :32-34
  if (sc && h3_take(pa)) {
    // split engine with the fused tanh-score epilogue: the features arrive pre-split, U is split here
    H3Gemm hg;
    RFN_TRY(h3_prepare(pa, *sc, feat_split, w_u, hg, st));
    hg.epi = 1; hg.g = g; hg.ldg = A; hg.wv = v_w; hg.score = P; hg.natt = N;
    {
      TagScope ts(tag_gemm);
      RFN_TRY(gemm_h3(hg, st));
    }
    TagScope ts(tag_attn);
    // engine mode 5: the weighted sum streams the bf16 copy of the features the GEMM just used (half the bytes)
    const bool a16 = h3_bf16() && feat_split != nullptr;
    return attention_from_scores(Afeat, P, tc_score_slices(A), v_b, z, ldz, nullptr, rows, N, D, 1, st,
                                 a16 ? feat_split->p0 : nullptr, a16 ? feat_split->ld : 0);
  }
  if (gemm_mode() >= 1 && rows * N >= 128 && gemm_tc_supported(pa)) {
    {
      TagScope ts(tag_gemm);
      RFN_TRY(gemm_tc(pa, tc_passes(gemm_mode()), g, A, v_w, P, N, st));
    }
    TagScope ts(tag_attn);
    return attention_from_scores(Afeat, P, tc_score_slices(A), v_b, z, ldz, nullptr, rows, N, D, 1, st);
  }
  // P = att_2_att_h(A) materialised (SIMT engine, or operands the tensor engine cannot take, e.g. unaligned views): the
  // workspace reserved only the partial-score slices for the tensor route, so check before writing rows * N * A floats
  if ((size_t)rows * N * A > p_cap) {
    set_error("attention: operands are not 16-byte aligned for the tensor engine and the workspace holds %zu of the %zu floats "
              "the fp32 SIMT route needs; pass aligned tensors or select rfn_set_gemm_mode(0) before sizing the workspace",
              p_cap, (size_t)rows * N * A);
    return RFN_ERR_UNSUPPORTED;
  }
  {
    TagScope ts(tag_gemm);
    RFN_TRY(gemm(pa, st));
  }
  TagScope ts(tag_attn);
  return attention_step(Afeat, P, g, v_w, v_b, z, ldz, nullptr, rows, N, D, A, 1, st);
}

static int thought_vectors(const rfn_dims& d, const float* const* prm, const float* const* fc,
                           const float* const* init_h, const float* const* init_c, const float* const* att, int rows,
                           float* TVc, float* h_out, float* c_out, float* TV_user, float* reason_pred, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
  WCache wc(d, h3_bf16());
  if (h3_mode()) wc.base = (const char*)prm[rfn_num_params(&d)];   // optional trailing slot: rfn_wcache_build's buffer
  const int J = d.J, R = d.rnn_size, S0 = d.num_review_steps_0, S1 = d.num_review_steps;
  const int K = d.top_words_count;
  const PIdx ix(d);
  Bump b(ws);
  TVWork w;
  const size_t need = carve_tv(d, rows, TV_user == nullptr, reason_pred != nullptr, b, w);
  if (need > ws_bytes) {
    set_error("rfn_thought_vectors: workspace %zu < %zu bytes", ws_bytes, need);
    return RFN_ERR_WORKSPACE;
  }
  float* TV = TV_user ? TV_user : w.TV;
  const size_t tv_stride = (size_t)rows * S0 * R;

  const bool h3 = h3_mode();
  const H3Ws* sc0 = h3 ? &w.h3[0] : nullptr;
  H3Operand featop[RFN_MAX_ENCODERS], tvop[RFN_MAX_ENCODERS];
  bool have_feat[RFN_MAX_ENCODERS] = {};

  // A.0  h_j^0 = c_j^0 = fc2h_j(fc_j)                      (misc/RecurrentFusionModel.py:202-208)
  RFN_TRY(for_each_encoder(J, st, [&](int j, cudaStream_t sj) -> int {
    float* cj = w.C + (size_t)j * rows * R;
    const float* hj = cj;
    if (h3) {   // the features are the x operand of all S0 att_2_att_h GEMMs of this encoder: split them once
      const int N = d.att_num[j], D = d.att_feat_size[j];
      if (h3_take(gemm1(att[j], D, prm[ix.s1(0, j, 0)], nullptr, D, w.P[j], d.att_hid_size, rows * N, d.att_hid_size))) {
        RFN_TRY(h3_presplit(att[j], D, D, rows * N, w.fsplit[j], w.fsplit_cap[j], &featop[j], sj));
        have_feat[j] = true;
      }
    }
    if (fc) {
      H3Operand wo[3];
      RFN_TRY(path_gemm(gemm1(fc[j], d.fc_feat_size[j], prm[ix.fc2h(j, 0)], prm[ix.fc2h(j, 1)], d.fc_feat_size[j], cj, R, rows, R),
                        h3 ? &w.h3[j] : nullptr, nullptr, wc.get(wc.i_fc2h(j), wo), sj));
    } else {
      RFN_CUDA(cudaMemcpyAsync(cj, init_c[j], (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, sj));
      hj = init_h[j];
    }
    RFN_CUDA(cudaMemcpy2DAsync(w.Hcat[0] + (size_t)j * R, (size_t)J * R * sizeof(float), hj, (size_t)R * sizeof(float),
                               (size_t)R * sizeof(float), rows, cudaMemcpyDeviceToDevice, sj));
    return RFN_OK;
  }));
  // A.3  stage 1: S0 fusion steps, distinct weights per (s, j); Jacobi update over encoders
  for (int s = 0; s < S0; ++s) {
    const float* Hin = w.Hcat[s & 1];
    float* Hout = w.Hcat[(s + 1) & 1];
    RFN_TRY(for_each_encoder(J, st, [&](int j, cudaStream_t sj) -> int {
      const int N = d.att_num[j], D = d.att_feat_size[j];
      float* cj = w.C + (size_t)j * rows * R;
      // z = Att(h_j, A_j)   -- att_2_att_h is the 89%-of-FLOPs contraction
      const H3Ws* scj = h3 ? &w.h3[j] : nullptr;
      H3Operand wu[1], wh[1], wg[3];
      RFN_TRY(attention_module(d, Hin + (size_t)j * R, J * R, att[j], N, D, prm[ix.s1(s, j, 0)], prm[ix.s1(s, j, 1)],
                               prm[ix.s1(s, j, 2)], prm[ix.s1(s, j, 3)], prm[ix.s1(s, j, 4)], prm[ix.s1(s, j, 5)], w.g[j],
                               w.P[j], w.p_cap[j], w.z[j], D, rows, TAG_GEMM_ATT2ATT, TAG_ATTN_S1, scj,
                               have_feat[j] ? &featop[j] : nullptr, wc.get(wc.i_s1(s, j, 0), wu), wc.get(wc.i_s1(s, j, 1), wh), sj));
      // G = H2h(H) + z2h(z)                                 (misc/RecurrentFusionModel.py:53)
      GemmArgs ga{};
      ga.src[0] = GemmSrc{Hin, prm[ix.s1(s, j, 6)], prm[ix.s1(s, j, 7)], J * R, J * R, J * R};
      ga.src[1] = GemmSrc{w.z[j], prm[ix.s1(s, j, 8)], prm[ix.s1(s, j, 9)], D, D, D};
      ga.nsrc = 2; ga.y = w.G[j]; ga.ldy = 4 * R; ga.M = rows; ga.N = 4 * R;
      {
        TagScope ts(TAG_GEMM_GATES);
        RFN_TRY(path_gemm(ga, scj, nullptr, wc.get(wc.i_s1(s, j, 2), wg), sj));
      }
      return lstm_cell(w.G[j], cj, nullptr, cj, Hout + (size_t)j * R, J * R, TV + j * tv_stride + (size_t)s * R, S0 * R, rows, R, sj);
    }));
  }
  const float* Hfin = w.Hcat[S0 & 1];
  // the thought vectors TV_j are the x operand of the stage-2 attention projections (S1 steps) and of the reason heads
  const bool have_tv = h3 && h3_shape_ok(rows * S0, d.att_hid_size);
  if (have_tv)
    for (int j = 0; j < J; ++j)
      RFN_TRY(h3_presplit(TV + j * tv_stride, R, R, rows * S0, w.fsplit[j], w.fsplit_cap[j], &tvop[j], st));
  if (reason_pred) {
    for (int j = 0; j < J; ++j) {  // reason_pred_j = max_s reason_linear_individual_j(h_j^s)   (:291,:303)
      H3Operand wo[3];
      RFN_TRY(path_gemm(gemm1(TV + j * tv_stride, R, prm[ix.reason_ind(j, 0)], prm[ix.reason_ind(j, 1)], R, w.rs, K, rows * S0, K), sc0,
                        have_tv ? &tvop[j] : nullptr, wc.get(wc.i_reason_ind(j), wo), st));
      RFN_TRY(max_over_steps(w.rs, reason_pred + (size_t)j * rows * K, rows, S0, K, st));
    }
  }
  // A.4  bridge: mean over encoders of (h, c)               (:307-309)
  float* hb[2] = {h_out, w.hx};
  float* h0 = hb[S1 & 1];
  RFN_TRY(mean_tensors(Hfin, (size_t)R, J, h0, R, (size_t)rows * R, R, J * R, st));
  RFN_TRY(mean_tensors(w.C, (size_t)rows * R, J, c_out, R, (size_t)rows * R, R, R, st));
  // A.5  stage 2: S1 review steps over the J thought-vector sets
  for (int s = 0; s < S1; ++s) {
    const float* hin = hb[(S1 + s) & 1];
    float* hout = hb[(S1 + s + 1) & 1];
    // the J attention modules of a review step share the query h: split it once
    H3Operand hop[1];
    const bool have_h = h3 && h3_shape_ok(rows, d.att_hid_size);
    if (have_h) RFN_TRY(h3_presplit(hin, R, R, rows, w.hsplit, w.hsplit_cap, hop, st));
    RFN_TRY(for_each_encoder(J, st, [&](int j, cudaStream_t sj) -> int {
      H3Operand wu[1], wh[1];
      return attention_module(d, hin, R, TV + j * tv_stride, S0, R, prm[ix.s2_att(s, j, 0)], prm[ix.s2_att(s, j, 1)],
                              prm[ix.s2_att(s, j, 2)], prm[ix.s2_att(s, j, 3)], prm[ix.s2_att(s, j, 4)], prm[ix.s2_att(s, j, 5)],
                              w.g[j], w.P[j], w.p_cap[j], w.z[j], R, rows, TAG_GEMM_OTHER, TAG_ATTN_SMALL, h3 ? &w.h3[j] : nullptr,
                              have_tv ? &tvop[j] : nullptr, wc.get(wc.i_s2(s, 2 * j), wu), wc.get(wc.i_s2(s, 2 * j + 1), wh), sj,
                              have_h ? hop : nullptr);
    }));
    // G = h2h(h) + sum_j z_2_h[j](z_j)       (misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:50-52)
    int jn = 0, grp = 0;
    bool first = true;
    while (first || jn < J) {
      GemmArgs ga{};
      int n = 0;
      if (first) ga.src[n++] = GemmSrc{hin, prm[ix.s2_h2h(s, 0)], prm[ix.s2_h2h(s, 1)], R, R, R};
      while (n < 3 && jn < J) {
        ga.src[n++] = GemmSrc{w.z[jn], prm[ix.s2_z2h(s, jn, 0)], prm[ix.s2_z2h(s, jn, 1)], R, R, R};
        ++jn;
      }
      ga.nsrc = n; ga.y = w.G[0]; ga.ldy = g2w(d); ga.M = rows; ga.N = g2w(d); ga.accumulate = first ? 0 : 1;
      {
        TagScope ts(TAG_GEMM_GATES);
        H3Operand wo[3];
        RFN_TRY(path_gemm(ga, sc0, nullptr, wc.get(wc.i_s2(s, 2 * J + grp), wo), st));
      }
      first = false;
      ++grp;
    }
    RFN_TRY(lstm_cell(w.G[0], c_out, hout, c_out, TVc + (size_t)s * R, S1 * R, nullptr, 0, rows, R, st, nullptr, 1.f, d.review_maxout));
  }
  if (reason_pred) {
    H3Operand wo[3];
    RFN_TRY(path_gemm(gemm1(TVc, R, prm[ix.reason(0)], prm[ix.reason(1)], R, w.rs, K, rows * S1, K), sc0, nullptr, wc.get(wc.i_reason(), wo), st));
    RFN_TRY(max_over_steps(w.rs, reason_pred + (size_t)J * rows * K, rows, S1, K, st));
  }
  return RFN_OK;
}

// ---- decoder ---------------------------------------------------------------------------------------
struct DecWork {
  float *Pdec, *hA, *hB, *cA, *cB, *x, *g, *z, *G, *logits, *rowmax, *logsum, *top_val;
  float *st_max, *st_sum, *st_val;   // per-slice statistics of the fused logits epilogue
  int32_t *top_idx, *tok, *src, *any, *st_idx;
  uint8_t* unfinished;
  BeamState bs;
  // engine modes 4 / 5: activation-split scratch and the decoder's weights split once per decode loop
  H3Ws h3;
  void* wsplit[3];          // h_2_att_h | i2h,h2h,z2h (joint row scale) | logit
  size_t wsplit_cap[3];
  H3Operand w_att[1], w_gates[3], w_logit[1];
  bool h3_ready;
  float* pd_part;           // persistent decoder (rows <= PD_MAX_ROWS): per-slice vocabulary statistics
};
static size_t carve_dec(const rfn_dims& d, int rowsA, int rows, int beam, int n_logit_bufs, Bump& b, DecWork& w) {
  const int R = d.rnn_size, A = d.att_hid_size, E = d.input_encoding_size, S1 = d.num_review_steps, V = d.vocab_plus1;
  const int L = d.seq_length;
  w.Pdec = b.take<float>((size_t)rowsA * S1 * A);
  w.hA = b.take<float>((size_t)rows * R);
  w.hB = b.take<float>((size_t)rows * R);
  w.cA = b.take<float>((size_t)rows * R);
  w.cB = b.take<float>((size_t)rows * R);
  w.x = b.take<float>((size_t)rows * E);
  w.g = b.take<float>((size_t)rows * A);
  w.z = b.take<float>((size_t)rows * R);
  w.G = b.take<float>((size_t)rows * gdw(d));
  w.logits = b.take<float>((size_t)rows * V * n_logit_bufs);
  w.rowmax = b.take<float>(rows);
  w.logsum = b.take<float>(rows);
  const int k = std::max(1, beam);
  w.top_val = b.take<float>((size_t)rows * k);
  w.top_idx = b.take<int32_t>((size_t)rows * k);
  const size_t sl = (size_t)tc_score_slices(V);
  w.st_max = b.take<float>(sl * rows);
  w.st_sum = b.take<float>(sl * rows);
  w.st_val = b.take<float>(sl * rows * k);
  w.st_idx = b.take<int32_t>(sl * rows * k);
  w.tok = b.take<int32_t>(rows);
  w.src = b.take<int32_t>(rows);
  w.any = b.take<int32_t>(L + 2);
  w.unfinished = b.take<uint8_t>(rows);
  w.h3 = H3Ws{};
  w.h3_ready = false;
  w.pd_part = (rows >= 1 && rows <= PD_MAX_ROWS) ? b.take<float>(pd_part_floats(d, rows, k)) : nullptr;
  for (int i = 0; i < 3; ++i) { w.wsplit[i] = nullptr; w.wsplit_cap[i] = 0; }
  if (h3_mode()) {
    w.h3.xcap = std::max({split3(rows, E, R, R), split1(rows, R), split1(rowsA * S1, R)});
    w.h3.x = b.take<char>(w.h3.xcap);
    w.h3.wcap = split1(A, R);
    w.h3.w = b.take<char>(w.h3.wcap);
    w.wsplit_cap[0] = split1(A, R);
    w.wsplit_cap[1] = split3(gdw(d), E, R, R);
    w.wsplit_cap[2] = split1(V, R);
    for (int i = 0; i < 3; ++i) w.wsplit[i] = b.take<char>(w.wsplit_cap[i]);
  }
  if (beam > 0) {
    const int images = rowsA;
    const size_t cap = (size_t)beam * L;
    w.bs.images = images; w.bs.beam = beam; w.bs.L = L;
    w.bs.beam_seq = b.take<int32_t>((size_t)2 * images * beam * L);
    w.bs.beam_lp = b.take<float>((size_t)2 * images * beam * L);
    w.bs.beam_sum = b.take<float>((size_t)images * beam);
    w.bs.finished = b.take<uint8_t>(images);
    w.bs.done_seq = b.take<int32_t>((size_t)images * cap * L);
    w.bs.done_lp = b.take<float>((size_t)images * cap * L);
    w.bs.done_p = b.take<float>((size_t)images * cap);
    w.bs.n_done = b.take<int32_t>(images);
  }
  return b.off;
}

// hoisted loop invariant: P_dec = att_2_att_h(TVc)   (the reference recomputes it every step,
// misc/LSTMSoftAttentionCore.py:64-66)
static int decoder_prepare(const rfn_dims& d, const float* const* prm, const float* TVc, int rowsA, float* Pdec, cudaStream_t st,
                           DecWork* w = nullptr, int rows = 0) {
  const PIdx ix(d);
  const int R = d.rnn_size, A = d.att_hid_size, S1 = d.num_review_steps, E = d.input_encoding_size, V = d.vocab_plus1;
  const H3Ws* sc = (w && h3_mode() && w->h3.x) ? &w->h3 : nullptr;
  WCache wc(d, h3_bf16());
  if (sc) wc.base = (const char*)prm[rfn_num_params(&d)];
  H3Operand wu[1];
  RFN_TRY(path_gemm(gemm1(TVc, R, prm[ix.dec(6)], prm[ix.dec(7)], R, Pdec, A, rowsA * S1, A), sc, nullptr, wc.get(wc.i_dec(0), wu), st));
  if (sc && rows >= 256 && wc.base) {
    wc.get(wc.i_dec(1), w->w_att);
    wc.get(wc.i_dec(2), w->w_gates);
    wc.get(wc.i_dec(3), w->w_logit);
    w->h3_ready = true;
  } else if (sc && rows >= 256) {
    // the decoder's weights are shared by all L + 1 steps: split them once (misc/LSTMSoftAttentionCore.py:13-58, logit :156)
    const bool bf = h3_bf16();
    {
      const float* ws[1] = {prm[ix.dec(8)]};
      const int ld[1] = {R}, K[1] = {R};
      RFN_TRY(h3_split(ws, ld, K, 1, A, bf, w->wsplit[0], w->w_att, st));
    }
    {
      const float* ws[3] = {prm[ix.dec(0)], prm[ix.dec(2)], prm[ix.dec(4)]};
      const int ld[3] = {E, R, R}, K[3] = {E, R, R};
      RFN_TRY(h3_split(ws, ld, K, 3, gdw(d), bf, w->wsplit[1], w->w_gates, st));
    }
    {
      const float* ws[1] = {prm[ix.logit(0)]};
      const int ld[1] = {R}, K[1] = {R};
      RFN_TRY(h3_split(ws, ld, K, 1, V, bf, w->wsplit[2], w->w_logit, st));
    }
    w->h3_ready = true;
  }
  return RFN_OK;
}

// one LSTMSoftAttentionCore step + vocab projection  (misc/LSTMSoftAttentionCore.py:60-102, logit :349)
// fuse_topk > 0: the vocab projection keeps only row max / log-sum-exp / top-k (tensor engine, fused epilogue)
// and fills w.rowmax / w.logsum / w.top_val / w.top_idx; returns 1 through *fused when that path was taken.
static int decoder_step(const rfn_dims& d, const float* const* prm, const float* TVc, const float* Pdec, int div,
                        const float* x, const float* hin, const float* cin, float* hout, float* cout, float* logits,
                        DecWork& w, int rows, cudaStream_t st, int fuse_topk = 0, int* fused = nullptr) {
  const PIdx ix(d);
  const int R = d.rnn_size, A = d.att_hid_size, E = d.input_encoding_size, S1 = d.num_review_steps, V = d.vocab_plus1;
  const H3Ws* sc = w.h3_ready ? &w.h3 : nullptr;
  RFN_TRY(path_gemm(gemm1(hin, R, prm[ix.dec(8)], prm[ix.dec(9)], R, w.g, A, rows, A), sc, nullptr, w.w_att, st));
  RFN_TRY(attention_step(TVc, Pdec, w.g, prm[ix.dec(10)], prm[ix.dec(11)], w.z, R, nullptr, rows, S1, R, A, div, st));
  GemmArgs ga{};
  ga.src[0] = GemmSrc{x, prm[ix.dec(0)], prm[ix.dec(1)], E, E, E};
  ga.src[1] = GemmSrc{hin, prm[ix.dec(2)], prm[ix.dec(3)], R, R, R};
  ga.src[2] = GemmSrc{w.z, prm[ix.dec(4)], prm[ix.dec(5)], R, R, R};
  ga.nsrc = 3; ga.y = w.G; ga.ldy = gdw(d); ga.M = rows; ga.N = gdw(d);
  {
    TagScope ts(TAG_GEMM_GATES);
    RFN_TRY(path_gemm(ga, sc, nullptr, w.w_gates, st));
  }
  RFN_TRY(lstm_cell(w.G, cin, hout, cout, nullptr, 0, nullptr, 0, rows, R, st, nullptr, 1.f, d.decoder_maxout));
  if (fused) *fused = 0;
  if (logits) {
    GemmArgs la = gemm1(hout, R, prm[ix.logit(0)], prm[ix.logit(1)], R, logits, V, rows, V);
    if (sc && h3_take(la)) {
      H3Gemm hg;
      RFN_TRY(h3_prepare(la, *sc, nullptr, w.w_logit, hg, st));
      if (fuse_topk > 0) {   // logits never reach HBM: per-slice max / sum-exp / top-k, merged below
        hg.epi = 2; hg.st_max = w.st_max; hg.st_sum = w.st_sum; hg.st_val = w.st_val; hg.st_idx = w.st_idx; hg.ktop = fuse_topk;
        RFN_TRY(gemm_h3(hg, st));
        RFN_TRY(vocab_merge(w.st_max, w.st_sum, w.st_val, w.st_idx, tc_score_slices(V), rows, fuse_topk, w.rowmax, w.logsum,
                            w.top_val, w.top_idx, st));
        if (fused) *fused = 1;
      } else {
        TagScope ts(TAG_GEMM_LOGIT);
        RFN_TRY(gemm_h3(hg, st));
      }
    } else if (fuse_topk > 0 && gemm_mode() >= 1 && rows >= 128 && gemm_tc_supported(la)) {
      RFN_TRY(gemm_tc_vocab(la, tc_passes(gemm_mode()), w.st_max, w.st_sum, w.st_val, w.st_idx, fuse_topk, st));
      RFN_TRY(vocab_merge(w.st_max, w.st_sum, w.st_val, w.st_idx, tc_score_slices(V), rows, fuse_topk, w.rowmax, w.logsum,
                          w.top_val, w.top_idx, st));
      if (fused) *fused = 1;
    } else {
      TagScope ts(TAG_GEMM_LOGIT);
      RFN_TRY(gemm(la, st));
    }
  }
  return RFN_OK;
}

// the decoder loop as ONE cooperative launch (rfn_decoder_persist.cu); the caller has prepared Pdec, the initial state in
// w.hA / w.cA, tok = 0 and (greedy) the unfinished / any flags or (beam) the beam state
static int decoder_persistent(const rfn_dims& d, const float* const* prm, const float* TVc, DecWork& w, int rows, int div, int beam,
                              int64_t* seq, float* seq_lp, float* lp_all, cudaStream_t st) {
  const PIdx ix(d);
  PDArgs a{};
  a.i2h_w = prm[ix.dec(0)]; a.i2h_b = prm[ix.dec(1)]; a.h2h_w = prm[ix.dec(2)]; a.h2h_b = prm[ix.dec(3)];
  a.z2h_w = prm[ix.dec(4)]; a.z2h_b = prm[ix.dec(5)];
  a.hatt_w = prm[ix.dec(8)]; a.hatt_b = prm[ix.dec(9)]; a.out_w = prm[ix.dec(10)]; a.out_b = prm[ix.dec(11)];
  a.logit_w = prm[ix.logit(0)]; a.logit_b = prm[ix.logit(1)]; a.embed = prm[ix.embed()];
  a.TVc = TVc; a.Pdec = w.Pdec;
  a.rows = rows; a.div = div; a.R = d.rnn_size; a.A = d.att_hid_size; a.E = d.input_encoding_size; a.V = d.vocab_plus1;
  a.S1 = d.num_review_steps; a.L = d.seq_length;
  a.hbuf[0] = w.hA; a.hbuf[1] = w.hB; a.cbuf[0] = w.cA; a.cbuf[1] = w.cB;
  a.g = w.g; a.z = w.z; a.logits = lp_all ? w.logits : nullptr; a.rowmax = w.rowmax; a.logsum = w.logsum;
  a.tok = w.tok; a.src = w.src;
  a.ktop = beam > 0 ? beam : 1;
  a.steps = (beam > 0 || !lp_all) ? d.seq_length : d.seq_length + 1;
  a.seq = seq; a.seq_lp = seq_lp; a.unfinished = w.unfinished; a.any_unfinished = w.any; a.lp_all = lp_all;
  a.beam = beam; a.bs = w.bs;
  pd_bind_parts(d, a, w.pd_part);
  return pd_launch(d, a, st);
}

static int ws_fail(const char* who, size_t have, size_t need) {
  set_error("%s: workspace %zu < %zu bytes", who, have, need);
  return RFN_ERR_WORKSPACE;
}

}  // namespace rfn

using namespace rfn;

extern "C" {

int rfn_set_concurrency(int on) {
  g_concurrency.store(on ? 1 : 0);
  return RFN_OK;
}

int rfn_num_param_slots(const rfn_dims* dims) { return dims ? rfn_num_params(dims) + 1 : RFN_ERR_INVALID; }

size_t rfn_wcache_bytes(const rfn_dims* dims, int bf16) {
  if (check_dims(dims) != RFN_OK) return 0;
  return WCache(*dims, bf16 != 0).bytes + 256;
}

int rfn_wcache_build(const rfn_dims* dims, const float* const* params, int bf16, void* wcache, size_t wcache_bytes,
                     rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && wcache, "rfn_wcache_build: null pointer");
  WCache wc(*dims, bf16 != 0);
  if (wc.bytes > wcache_bytes) {
    set_error("rfn_wcache_build: buffer %zu < %zu bytes", wcache_bytes, wc.bytes);
    return RFN_ERR_WORKSPACE;
  }
  for (const WGroupDesc& g : wc.g) {
    const float* ws[3];
    int ld[3];
    for (int s = 0; s < g.nsrc; ++s) { ws[s] = params[g.pidx[s]]; ld[s] = g.K[s]; }
    H3Operand out[3];
    RFN_TRY(h3_split(ws, ld, g.K, g.nsrc, g.n_rows, bf16 != 0, (char*)wcache + g.off, out, (cudaStream_t)stream));
  }
  return RFN_OK;
}

size_t rfn_workspace_bytes(const rfn_dims* dims, int rows, int dec_rows) {
  if (check_dims(dims) != RFN_OK || rows < 0 || dec_rows < 0) return 0;
  Bump b1(nullptr);
  TVWork tw;
  const size_t a = carve_tv(*dims, rows, true, true, b1, tw);
  const int beam = rows > 0 ? std::max(1, (dec_rows + rows - 1) / rows) : 1;
  Bump b2(nullptr);
  DecWork dw;
  const size_t c = carve_dec(*dims, rows, dec_rows, std::min(beam, RFN_MAX_BEAM), 1, b2, dw);
  return std::max(a, c) + 1024;
}

size_t rfn_ensemble_workspace_bytes(const rfn_dims* dims, int n_models, int images, int beam) {
  if (check_dims(dims) != RFN_OK || n_models < 1 || images < 0 || beam < 1 || beam > RFN_MAX_BEAM) return 0;
  Bump b(nullptr);
  DecWork dw;
  carve_dec(*dims, images, images * beam, beam, n_models + 1, b, dw);
  const int R = dims->rnn_size, A = dims->att_hid_size, S1 = dims->num_review_steps;
  const size_t rows = (size_t)images * beam;
  // beam: 4 state buffers of rows x R + images x S1 x A; greedy (beam 1, rfn_ensemble_decode_greedy): 3 + rows x S1 x A
  const size_t per_model = (4 * rows * R + rows * S1 * A) * sizeof(float) + 5 * 256;
  return b.off + per_model * n_models + 1024;
}

int rfn_thought_vectors(const rfn_dims* dims, const float* const* params, const float* const* fc,
                        const float* const* init_h, const float* const* init_c, const float* const* att, int rows, float* TVc, float* h_out, float* c_out, float* TV, float* reason_pred, void* workspace,
                        size_t workspace_bytes, rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && att && TVc && h_out && c_out && workspace, "rfn_thought_vectors: null pointer");
  RFN_CHECK_ARG(fc || (init_h && init_c), "rfn_thought_vectors: need fc or init_h/init_c");
  RFN_CHECK_ARG(rows >= 1, "rfn_thought_vectors: rows=%d", rows);
  for (int j = 0; j < dims->J; ++j)
    RFN_CHECK_ARG(att[j] && (fc ? fc[j] != nullptr : (init_h[j] && init_c[j])), "rfn_thought_vectors: null feature pointer %d", j);
  return thought_vectors(*dims, params, fc, init_h, init_c, att, rows, TVc, h_out, c_out, TV, reason_pred, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

int rfn_one_time_step(const rfn_dims* dims, const float* const* params, const float* xt, const float* TVc, int div,
                      const float* h_in, const float* c_in, float* h_out, float* c_out, float* logits, int rows,
                      void* workspace, size_t workspace_bytes, rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && xt && TVc && h_in && c_in && h_out && c_out && logits && workspace, "rfn_one_time_step: null pointer");
  RFN_CHECK_ARG(rows >= 1 && div >= 1 && h_in != h_out, "rfn_one_time_step: rows/div invalid or h_in aliases h_out");
  cudaStream_t st = (cudaStream_t)stream;
  const int rowsA = (rows + div - 1) / div;
  Bump b(workspace);
  DecWork w{};
  w.Pdec = b.take<float>((size_t)rowsA * dims->num_review_steps * dims->att_hid_size);
  w.g = b.take<float>((size_t)rows * dims->att_hid_size);
  w.z = b.take<float>((size_t)rows * dims->rnn_size);
  w.G = b.take<float>((size_t)rows * 4 * dims->rnn_size);
  if (b.off > workspace_bytes) return ws_fail("rfn_one_time_step", workspace_bytes, b.off);
  RFN_TRY(decoder_prepare(*dims, params, TVc, rowsA, w.Pdec, st));
  return decoder_step(*dims, params, TVc, w.Pdec, div, xt, h_in, c_in, h_out, c_out, logits, w, rows, st);
}

int rfn_decode_teacher_forced(const rfn_dims* dims, const float* const* params, const float* TVc, const float* h0,
                              const float* c0, const int64_t* seq, int ld_seq, int T, int rows, float* logprobs,
                              void* workspace, size_t workspace_bytes, rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && TVc && h0 && c0 && seq && logprobs && workspace, "rfn_decode_teacher_forced: null pointer");
  RFN_CHECK_ARG(rows >= 1 && T >= 1 && T <= ld_seq, "rfn_decode_teacher_forced: rows=%d T=%d ld_seq=%d", rows, T, ld_seq);
  cudaStream_t st = (cudaStream_t)stream;
  const rfn_dims& d = *dims;
  const PIdx ix(d);
  const int R = d.rnn_size, V = d.vocab_plus1, E = d.input_encoding_size;
  Bump b(workspace);
  DecWork w{};
  if (carve_dec(d, rows, rows, 0, 1, b, w) > workspace_bytes) return ws_fail("rfn_decode_teacher_forced", workspace_bytes, b.off);
  RFN_TRY(decoder_prepare(d, params, TVc, rows, w.Pdec, st, &w, rows));
  RFN_CUDA(cudaMemcpyAsync(w.hA, h0, (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RFN_CUDA(cudaMemcpyAsync(w.cA, c0, (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  float* hb[2] = {w.hA, w.hB};
  for (int t = 0; t < T; ++t) {
    RFN_TRY(embed_gather_i64(seq + t, ld_seq, params[ix.embed()], w.x, rows, E, V, st));          // :276
    RFN_TRY(decoder_step(d, params, TVc, w.Pdec, 1, w.x, hb[t & 1], w.cA, hb[(t + 1) & 1], w.cA, w.logits, w, rows, st));
    RFN_TRY(log_softmax_rows(w.logits, V, logprobs + (size_t)t * V, T * V, rows, V, st));        // :278
  }
  return RFN_OK;
}

int rfn_decode_sample(const rfn_dims* dims, const float* const* params, const float* TVc, const float* h0,
                      const float* c0, int rows, const float* uniforms, float temperature, int64_t* seq,
                      float* seq_logprobs, float* lp_all, int32_t* d_T, void* workspace, size_t workspace_bytes,
                      rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && TVc && h0 && c0 && seq && seq_logprobs && d_T && workspace, "rfn_decode_sample: null pointer");
  RFN_CHECK_ARG(rows >= 1 && temperature > 0.f, "rfn_decode_sample: rows=%d temperature=%f", rows, temperature);
  cudaStream_t st = (cudaStream_t)stream;
  const rfn_dims& d = *dims;
  const PIdx ix(d);
  const int R = d.rnn_size, V = d.vocab_plus1, E = d.input_encoding_size, L = d.seq_length;
  Bump b(workspace);
  DecWork w{};
  if (carve_dec(d, rows, rows, 0, 1, b, w) > workspace_bytes) return ws_fail("rfn_decode_sample", workspace_bytes, b.off);
  RFN_TRY(decoder_prepare(d, params, TVc, rows, w.Pdec, st, &w, rows));
  RFN_CUDA(cudaMemcpyAsync(w.hA, h0, (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RFN_CUDA(cudaMemcpyAsync(w.cA, c0, (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RFN_CUDA(cudaMemsetAsync(w.tok, 0, (size_t)rows * sizeof(int32_t), st));                       // t == 0: BOS (:617-618)
  RFN_CUDA(cudaMemsetAsync(w.any, 0, (size_t)(L + 2) * sizeof(int32_t), st));
  RFN_CUDA(cudaMemsetAsync(w.unfinished, 0, (size_t)rows, st));
  if (!uniforms && w.pd_part && pd_supported(d, rows)) {   // greedy, small batch: the whole loop in one cooperative launch
    RFN_TRY(decoder_persistent(d, params, TVc, w, rows, 1, 0, seq, seq_logprobs, lp_all, st));
    return sample_finalize(w.any, L, d_T, st);
  }
  float* hb[2] = {w.hA, w.hB};
  for (int t = 0; t <= L; ++t) {
    if (t >= 1)
      RFN_TRY(sample_select(w.logits, V, V, w.rowmax, w.logsum, w.top_val, w.top_idx, uniforms, L, temperature, t, L,
                            w.tok, w.unfinished, w.any, seq, seq_logprobs, rows, st));
    RFN_TRY(embed_gather_i32(w.tok, params[ix.embed()], w.x, rows, E, V, st));                   // :637
    RFN_TRY(decoder_step(d, params, TVc, w.Pdec, 1, w.x, hb[t & 1], w.cA, hb[(t + 1) & 1], w.cA, w.logits, w, rows, st));
    RFN_TRY(vocab_stats_topk(w.logits, V, rows, V, 1, w.rowmax, w.logsum, w.top_val, w.top_idx, st));
    if (lp_all) RFN_TRY(vocab_write_lp(w.logits, V, w.rowmax, w.logsum, lp_all + (size_t)t * V, (size_t)(L + 1) * V, rows, V, st));
  }
  return sample_finalize(w.any, L, d_T, st);
}

static int beam_init(DecWork& w, int images, int beam, int L, cudaStream_t st) {
  const size_t rows = (size_t)images * beam;
  RFN_CUDA(cudaMemsetAsync(w.bs.beam_seq, 0, 2 * rows * L * sizeof(int32_t), st));
  RFN_CUDA(cudaMemsetAsync(w.bs.beam_lp, 0, 2 * rows * L * sizeof(float), st));
  RFN_CUDA(cudaMemsetAsync(w.bs.beam_sum, 0, rows * sizeof(float), st));
  RFN_CUDA(cudaMemsetAsync(w.bs.finished, 0, (size_t)images, st));
  RFN_CUDA(cudaMemsetAsync(w.bs.n_done, 0, (size_t)images * sizeof(int32_t), st));
  RFN_CUDA(cudaMemsetAsync(w.tok, 0, rows * sizeof(int32_t), st));
  return RFN_OK;
}

int rfn_decode_beam(const rfn_dims* dims, const float* const* params, const float* TVc, const float* h0, const float* c0,
                    int images, int beam, int64_t* seq, float* seq_logprobs, int32_t* done_seq, float* done_logps,
                    float* done_p, int32_t* n_done, void* workspace, size_t workspace_bytes, rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params && TVc && h0 && c0 && seq && seq_logprobs && workspace, "rfn_decode_beam: null pointer");
  RFN_CHECK_ARG(images >= 1 && beam >= 1 && beam <= RFN_MAX_BEAM && beam <= dims->vocab_plus1,
                "rfn_decode_beam: images=%d beam=%d (max %d)", images, beam, RFN_MAX_BEAM);
  cudaStream_t st = (cudaStream_t)stream;
  const rfn_dims& d = *dims;
  const PIdx ix(d);
  const int R = d.rnn_size, V = d.vocab_plus1, E = d.input_encoding_size, L = d.seq_length;
  const int rows = images * beam;
  Bump b(workspace);
  DecWork w{};
  if (carve_dec(d, images, rows, beam, 1, b, w) > workspace_bytes) return ws_fail("rfn_decode_beam", workspace_bytes, b.off);
  RFN_TRY(decoder_prepare(d, params, TVc, images, w.Pdec, st, &w, rows));
  RFN_TRY(beam_init(w, images, beam, L, st));
  if (w.pd_part && pd_supported(d, rows)) {   // few images: the whole beam search in one cooperative launch
    RFN_TRY(gather_rows(h0, nullptr, beam, w.hA, rows, R, st));
    RFN_TRY(gather_rows(c0, nullptr, beam, w.cA, rows, R, st));
    RFN_TRY(decoder_persistent(d, params, TVc, w, rows, beam, beam, nullptr, nullptr, nullptr, st));
    return beam_finalize(w.bs, seq, seq_logprobs, done_seq, done_logps, done_p, n_done, st);
  }
  // expand each image's stage-2 state to `beam` identical rows (:376-394)
  RFN_TRY(gather_rows(h0, nullptr, beam, w.hB, rows, R, st));
  RFN_TRY(gather_rows(c0, nullptr, beam, w.cB, rows, R, st));
  for (int t = 0; t <= L; ++t) {
    if (t >= 1) {
      RFN_TRY(beam_merge(w.bs, t, w.top_val, w.top_idx, w.src, w.tok, st));                      // :465-514
      if (t == L) break;  // the reference runs one more, unused, decoder step (:526)
      RFN_TRY(gather_rows(w.hA, w.src, 1, w.hB, rows, R, st));                                   // :499-501
      RFN_TRY(gather_rows(w.cA, w.src, 1, w.cB, rows, R, st));
    }
    RFN_TRY(embed_gather_i32(w.tok, params[ix.embed()], w.x, rows, E, V, st));                   // :517-521
    int fused = 0;   // vocab projection + log_softmax statistics + top-beam in one kernel when the tensor engine runs
    RFN_TRY(decoder_step(d, params, TVc, w.Pdec, beam, w.x, w.hB, w.cB, w.hA, w.cA, w.logits, w, rows, st, beam, &fused));
    if (!fused) RFN_TRY(vocab_stats_topk(w.logits, V, rows, V, beam, w.rowmax, w.logsum, w.top_val, w.top_idx, st));  // :463, :527
  }
  return beam_finalize(w.bs, seq, seq_logprobs, done_seq, done_logps, done_p, n_done, st);
}

int rfn_ensemble_decode_beam(const rfn_dims* dims, int n_models, const float* const* const* params_m,
                             const float* const* TVc_m, const float* const* h0_m, const float* const* c0_m, int images,
                             int beam, int64_t* seq, float* seq_logprobs, int32_t* done_seq, float* done_logps,
                             float* done_p, int32_t* n_done, void* workspace, size_t workspace_bytes,
                             rfn_stream_t stream) {
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params_m && TVc_m && h0_m && c0_m && seq && seq_logprobs && workspace, "rfn_ensemble_decode_beam: null pointer");
  RFN_CHECK_ARG(n_models >= 1 && n_models <= 8, "rfn_ensemble_decode_beam: n_models=%d not in 1..8", n_models);
  RFN_CHECK_ARG(images >= 1 && beam >= 1 && beam <= RFN_MAX_BEAM, "rfn_ensemble_decode_beam: images=%d beam=%d", images, beam);
  cudaStream_t st = (cudaStream_t)stream;
  const rfn_dims& d = *dims;
  const PIdx ix(d);
  const int R = d.rnn_size, V = d.vocab_plus1, E = d.input_encoding_size, L = d.seq_length, A = d.att_hid_size;
  const int S1 = d.num_review_steps;
  const int rows = images * beam;
  Bump b(workspace);
  DecWork w{};
  carve_dec(d, images, rows, beam, n_models + 1, b, w);
  float *hA[8], *hB[8], *cA[8], *cB[8], *Pd[8];
  for (int m = 0; m < n_models; ++m) {
    hA[m] = b.take<float>((size_t)rows * R); hB[m] = b.take<float>((size_t)rows * R);
    cA[m] = b.take<float>((size_t)rows * R); cB[m] = b.take<float>((size_t)rows * R);
    Pd[m] = b.take<float>((size_t)images * S1 * A);
  }
  if (b.off > workspace_bytes) return ws_fail("rfn_ensemble_decode_beam", workspace_bytes, b.off);
  float* mean = w.logits + (size_t)n_models * rows * V;
  RFN_TRY(beam_init(w, images, beam, L, st));
  for (int m = 0; m < n_models; ++m) {
    RFN_TRY(decoder_prepare(d, params_m[m], TVc_m[m], images, Pd[m], st));
    RFN_TRY(gather_rows(h0_m[m], nullptr, beam, hB[m], rows, R, st));
    RFN_TRY(gather_rows(c0_m[m], nullptr, beam, cB[m], rows, R, st));
  }
  PtrList8 pl{};
  for (int m = 0; m < n_models; ++m) pl.p[m] = w.logits + (size_t)m * rows * V;
  for (int t = 0; t <= L; ++t) {
    if (t >= 1) {
      RFN_TRY(beam_merge(w.bs, t, w.top_val, w.top_idx, w.src, w.tok, st));
      if (t == L) break;
      for (int m = 0; m < n_models; ++m) {  // every model's state is forked with the same q (eval_utils.py:604-611)
        RFN_TRY(gather_rows(hA[m], w.src, 1, hB[m], rows, R, st));
        RFN_TRY(gather_rows(cA[m], w.src, 1, cB[m], rows, R, st));
      }
    }
    for (int m = 0; m < n_models; ++m) {
      RFN_TRY(embed_gather_i32(w.tok, params_m[m][ix.embed()], w.x, rows, E, V, st));
      RFN_TRY(decoder_step(d, params_m[m], TVc_m[m], Pd[m], beam, w.x, hB[m], cB[m], hA[m], cA[m],
                           w.logits + (size_t)m * rows * V, w, rows, st));
    }
    RFN_TRY(mean_logits8(pl, n_models, mean, (size_t)rows * V, st));                            // eval_utils.py:282-287
    RFN_TRY(vocab_stats_topk(mean, V, rows, V, beam, w.rowmax, w.logsum, w.top_val, w.top_idx, st));
  }
  return beam_finalize(w.bs, seq, seq_logprobs, done_seq, done_logps, done_p, n_done, st);
}

int rfn_ensemble_decode_greedy(const rfn_dims* dims, int n_models, const float* const* const* params_m,
                               const float* const* TVc_m, const float* const* h0_m, const float* const* c0_m, int rows,
                               int64_t* seq, float* seq_logprobs, int32_t* d_T, void* workspace, size_t workspace_bytes,
                               rfn_stream_t stream) {
  // eval_utils.py:729-975 (what eval_ensemble.sh runs, --beam_size 1): all images of the batch advance together; per step
  // every model embeds the UNMASKED argmax (:880-883 before :893), the logits are averaged (:282-287) and the argmax of the
  // log-softmax is taken; the loop ends when no row is unfinished (:889-891, read back by the caller through d_T)
  RFN_TRY(check_dims(dims));
  RFN_CHECK_ARG(params_m && TVc_m && h0_m && c0_m && seq && seq_logprobs && d_T && workspace, "rfn_ensemble_decode_greedy: null pointer");
  RFN_CHECK_ARG(n_models >= 1 && n_models <= 8 && rows >= 1, "rfn_ensemble_decode_greedy: n_models=%d rows=%d", n_models, rows);
  cudaStream_t st = (cudaStream_t)stream;
  const rfn_dims& d = *dims;
  const PIdx ix(d);
  const int R = d.rnn_size, V = d.vocab_plus1, E = d.input_encoding_size, L = d.seq_length, A = d.att_hid_size;
  const int S1 = d.num_review_steps;
  Bump b(workspace);
  DecWork w{};
  carve_dec(d, rows, rows, 0, n_models + 1, b, w);
  float *hA[8], *hB[8], *cS[8], *Pd[8];
  for (int m = 0; m < n_models; ++m) {
    hA[m] = b.take<float>((size_t)rows * R); hB[m] = b.take<float>((size_t)rows * R);
    cS[m] = b.take<float>((size_t)rows * R);
    Pd[m] = b.take<float>((size_t)rows * S1 * A);
  }
  if (b.off > workspace_bytes) return ws_fail("rfn_ensemble_decode_greedy", workspace_bytes, b.off);
  float* mean = w.logits + (size_t)n_models * rows * V;
  for (int m = 0; m < n_models; ++m) {
    RFN_TRY(decoder_prepare(d, params_m[m], TVc_m[m], rows, Pd[m], st));
    RFN_CUDA(cudaMemcpyAsync(hA[m], h0_m[m], (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    RFN_CUDA(cudaMemcpyAsync(cS[m], c0_m[m], (size_t)rows * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  RFN_CUDA(cudaMemsetAsync(w.tok, 0, (size_t)rows * sizeof(int32_t), st));
  RFN_CUDA(cudaMemsetAsync(w.any, 0, (size_t)(L + 2) * sizeof(int32_t), st));
  RFN_CUDA(cudaMemsetAsync(w.unfinished, 0, (size_t)rows, st));
  PtrList8 pl{};
  for (int m = 0; m < n_models; ++m) pl.p[m] = w.logits + (size_t)m * rows * V;
  for (int t = 0; t <= L; ++t) {
    if (t >= 1)
      RFN_TRY(sample_select(mean, V, V, w.rowmax, w.logsum, w.top_val, w.top_idx, nullptr, L, 1.f, t, L, w.tok, w.unfinished, w.any,
                            seq, seq_logprobs, rows, st));
    for (int m = 0; m < n_models; ++m) {
      float* hin = (t & 1) ? hB[m] : hA[m];
      float* hout = (t & 1) ? hA[m] : hB[m];
      RFN_TRY(embed_gather_i32(w.tok, params_m[m][ix.embed()], w.x, rows, E, V, st));
      RFN_TRY(decoder_step(d, params_m[m], TVc_m[m], Pd[m], 1, w.x, hin, cS[m], hout, cS[m], w.logits + (size_t)m * rows * V, w, rows, st));
    }
    RFN_TRY(mean_logits8(pl, n_models, mean, (size_t)rows * V, st));
    RFN_TRY(vocab_stats_topk(mean, V, rows, V, 1, w.rowmax, w.logsum, w.top_val, w.top_idx, st));
  }
  return sample_finalize(w.any, L, d_T, st);
}

}  // extern "C"
