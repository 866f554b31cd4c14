// Persistent cooperative decoder: all L (+1) timesteps of LSTMSoftAttentionCore + logit + log-softmax statistics + token
// selection (greedy argmax or the batched beam merge) in ONE launch, for small row counts (rows <= PD_MAX_ROWS).
//
// The decoder is the only LSTM of the path whose weights are shared by all timesteps (misc/LSTMSoftAttentionCore.py:13-58;
// stages 1-2 own distinct weights per step), and at batch 16 / beam 3 x 5 images its step is a dependent chain
// of ~8 tiny kernels (bench.py config1_latency: 8 us per kernel even when replayed from a CUDA graph).  Here one CTA per SM
// stays resident for the whole loop:
//   * the gate weights [i2h | h2h | z2h] of the CTA's hidden units (4 gates x UPS units x (E + 2R) floats, 96 KB for
//     R = E = 512 on 148 SMs) and its rows of h_2_att_h are loaded into shared memory ONCE and reused by every step;
//   * per step: (C) the CTA's gate columns for all rows + the LSTM cell of its hidden units, (D) the CTA's slice of the
//     vocabulary: logits, per-slice max / sum-exp / top-k, and the NEXT step's query projection g = h_2_att_h(h') for the
//     CTA's attention columns from the same staged h', (E) per row (greedy) or per image (beam): merge of the slices and the
//     token selection with the reference's bookkeeping, followed by the next step's attention (scores, softmax over the S1
//     thought vectors, context z) -- three grid-wide barriers per step (four in beam mode) instead of eight kernel
//     boundaries; the logit weights (19.4 MB) stream from L2;
//   * the beam search's state re-ordering is an index indirection (src_row) instead of a gather pass.
// Multinomial sampling and teacher forcing keep the per-step launch path (rfn_path.cu).
#include <cooperative_groups.h>

#include <algorithm>

#include <atomic>

#include "rfn_decoder_persist.cuh"

namespace cg = cooperative_groups;

namespace rfn {

constexpr int PD_THREADS = 512;
constexpr int PD_NW = PD_THREADS / 32;
constexpr int PD_RG = 16;          // rows staged in shared memory at a time
constexpr int PD_VC = 3;           // vocabulary columns a warp processes per pass
constexpr int PD_SL = 8;           // vocabulary slices per lane held in registers during the merge (nslice <= 256)

__device__ __forceinline__ float pd_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float pd_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ bool pd_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }
// 16-byte global -> shared copy that does not occupy a register or stall the thread (cp.async, L2 only): the staging loops
// issue all their copies back to back and wait once, instead of one dependent load -> store round trip per iteration
__device__ __forceinline__ void pd_cp16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pd_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__global__ void __launch_bounds__(PD_THREADS, 1) decoder_persist_kernel(const PDArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float sm[];
  const int R = a.R, A = a.A, E = a.E, V = a.V, S1 = a.S1, L = a.L, rows = a.rows;
  const int KG = E + 2 * R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, NB = gridDim.x;
  // ---- shared-memory carve ----
  float* s_wg = sm;                                   // [4*UPS][KG] gate weights of this CTA's hidden units (resident)
  float* s_wa = s_wg + (size_t)4 * a.UPS * KG;        // [OPA][R]    h_2_att_h rows of this CTA (resident)
  float* s_act = s_wa + (size_t)a.OPA * R;            // [PD_RG][KG] staged activations [x | h | z]
  float* s_G = s_act + (size_t)PD_RG * KG;            // [4*UPS][PD_RG] gate pre-activations
  float* s_logit = s_G + (size_t)4 * a.UPS * PD_RG;   // [PD_RG][VPS]
  float* s_bg = s_logit + (size_t)PD_RG * a.VPS;      // [4*UPS]
  float* s_ba = s_bg + 4 * a.UPS;                     // [OPA]
  float* s_e = s_ba + a.OPA;                          // [max(S1, 32)]
  float* s_topv = s_e + max(S1, 32);                  // [RFN_MAX_BEAM][RFN_MAX_BEAM]
  int32_t* s_topi = reinterpret_cast<int32_t*>(s_topv + RFN_MAX_BEAM * RFN_MAX_BEAM);
  int32_t* s_src = s_topi + RFN_MAX_BEAM * RFN_MAX_BEAM;   // [PD_RG] state source rows of the staged group
  int32_t* s_tok = s_src + PD_RG;                          // [PD_RG] clamped input tokens of the staged group

  // ---- resident weights: loaded once for all timesteps ----
  const int u0 = b * a.UPS;
  const int nu = max(0, min(a.UPS, R - u0));          // hidden units owned
  const int oa0 = b * a.OPA;
  const int noa = max(0, min(a.OPA, A - oa0));        // attention columns owned
  const int v0 = b * a.VPS;
  const int nv = max(0, min(a.VPS, V - v0));          // vocabulary columns owned
  for (int o = 0; o < 4 * nu; ++o) {                  // local output o = unit * 4 + gate; gate rows are [i | f | o | g] blocks of R
    const int grow = (o & 3) * R + u0 + (o >> 2);
    float* dst = s_wg + (size_t)o * KG;
    for (int k = tid; k < E; k += PD_THREADS) dst[k] = __ldg(a.i2h_w + (size_t)grow * E + k);
    for (int k = tid; k < R; k += PD_THREADS) dst[E + k] = __ldg(a.h2h_w + (size_t)grow * R + k);
    for (int k = tid; k < R; k += PD_THREADS) dst[E + R + k] = __ldg(a.z2h_w + (size_t)grow * R + k);
    if (tid == 0) s_bg[o] = __ldg(a.i2h_b + grow) + __ldg(a.h2h_b + grow) + __ldg(a.z2h_b + grow);
  }
  for (int o = 0; o < noa; ++o) {
    for (int k = tid; k < R; k += PD_THREADS) s_wa[(size_t)o * R + k] = __ldg(a.hatt_w + (size_t)(oa0 + o) * R + k);
    if (tid == 0) s_ba[o] = __ldg(a.hatt_b + oa0 + o);
  }
  const float out_b = __ldg(a.out_b);
  for (int r = b * PD_THREADS + tid; r < rows; r += NB * PD_THREADS) a.src[r] = r;   // states start in slot order
  __syncthreads();

  auto stamp = [&](int t, int phase) {
    if (a.dbg && b == 0 && tid == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      a.dbg[t * 8 + phase] = (long long)ns;
    }
  };
  // g[row, oa0 + o] = h_2_att_h(h[row]) for the nr rows whose h sits at s_act[row][hoff ..]  (misc/LSTMSoftAttentionCore.py:69)
  auto project_g = [&](int rg0, int nr, int hoff) {
    for (int p = warp; p < noa * nr; p += PD_NW) {
      const int o = p / nr, row = p % nr;
      const float* w = s_wa + (size_t)o * R;
      const float* x = s_act + (size_t)row * KG + hoff;
      float acc = 0.f;
      for (int k = lane * 4; k < R; k += 128) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + k);
        const float4 x4 = *reinterpret_cast<const float4*>(x + k);
        acc = fmaf(w4.x, x4.x, acc); acc = fmaf(w4.y, x4.y, acc); acc = fmaf(w4.z, x4.z, acc); acc = fmaf(w4.w, x4.w, acc);
      }
      acc = pd_warp_sum(acc);
      if (lane == 0) a.g[(size_t)(rg0 + row) * A + oa0 + o] = acc + s_ba[o];
    }
  };
  // attention of row r over the S1 combined thought vectors, query projection g[gr]  (:60-79); the whole CTA works on it
  auto attention_row = [&](int r, int gr) {
    const int ra = r / a.div;
    float* s_g = s_act;                                 // A floats
    for (int k = tid; k < A; k += PD_THREADS) s_g[k] = a.g[(size_t)gr * A + k];
    __syncthreads();
    for (int n = warp; n < S1; n += PD_NW) {
      const float* pn = a.Pdec + ((size_t)ra * S1 + n) * A;
      float acc = 0.f;
      for (int k = lane * 4; k < A; k += 128) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pn + k));
        const float4 gg = *reinterpret_cast<const float4*>(s_g + k);
        const float4 ww = __ldg(reinterpret_cast<const float4*>(a.out_w + k));
        acc = fmaf(ww.x, tanhf(p.x + gg.x), acc);
        acc = fmaf(ww.y, tanhf(p.y + gg.y), acc);
        acc = fmaf(ww.z, tanhf(p.z + gg.z), acc);
        acc = fmaf(ww.w, tanhf(p.w + gg.w), acc);
      }
      acc = pd_warp_sum(acc);
      if (lane == 0) s_e[n] = acc + out_b;
    }
    __syncthreads();
    if (tid == 0) {                                     // softmax over S1 (unmasked)
      float m = -INFINITY;
      for (int n = 0; n < S1; ++n) m = fmaxf(m, s_e[n]);
      float sum = 0.f;
      for (int n = 0; n < S1; ++n) { const float ex = expf(s_e[n] - m); s_e[n] = ex; sum += ex; }
      for (int n = 0; n < S1; ++n) s_e[n] = s_e[n] / sum;
    }
    __syncthreads();
    for (int d = tid; d < R; d += PD_THREADS) {
      float acc = 0.f;
      for (int n = 0; n < S1; ++n) acc = fmaf(s_e[n], __ldg(a.TVc + ((size_t)ra * S1 + n) * R + d), acc);
      a.z[(size_t)r * R + d] = acc;
    }
    __syncthreads();
  };

  // ================= prologue: g and the attention context of step 0 from the initial state =================
  grid.sync();                                          // src written
  if (noa > 0) {
    for (int rg0 = 0; rg0 < rows; rg0 += PD_RG) {
      const int nr = min(PD_RG, rows - rg0);
      for (int i = tid; i < nr * (R / 4); i += PD_THREADS) {
        const int row = i / (R / 4), q = i % (R / 4);
        pd_cp16(s_act + (size_t)row * KG + q * 4, a.hbuf[0] + (size_t)(rg0 + row) * R + q * 4);
      }
      pd_cp_wait();
      __syncthreads();
      project_g(rg0, nr, 0);
      __syncthreads();
    }
  }
  grid.sync();
  for (int r = b; r < rows; r += NB) attention_row(r, r);
  grid.sync();

  for (int t = 0; t < a.steps; ++t) {
    stamp(t, 0);
    const float* h_in = a.hbuf[t & 1];
    const float* c_in = a.cbuf[t & 1];
    float* h_out = a.hbuf[(t + 1) & 1];
    float* c_out = a.cbuf[(t + 1) & 1];

    // ================= (C) gates of this CTA's hidden units for all rows + LSTM cell =================
    if (nu > 0) {
      for (int rg0 = 0; rg0 < rows; rg0 += PD_RG) {
        const int nr = min(PD_RG, rows - rg0);
        if (tid < nr) {
          s_src[tid] = a.src[rg0 + tid];
          int tk = a.tok[rg0 + tid];
          s_tok[tid] = tk < 0 ? 0 : (tk >= V ? V - 1 : tk);
        }
        __syncthreads();
        for (int i = tid; i < nr * (E / 4); i += PD_THREADS) {       // x = embed[token] (the UNMASKED token, :637)
          const int row = i / (E / 4), q = i % (E / 4);
          pd_cp16(s_act + (size_t)row * KG + q * 4, a.embed + (size_t)s_tok[row] * E + q * 4);
        }
        for (int i = tid; i < nr * (R / 4); i += PD_THREADS) {
          const int row = i / (R / 4), q = i % (R / 4);
          pd_cp16(s_act + (size_t)row * KG + E + q * 4, h_in + (size_t)s_src[row] * R + q * 4);
          pd_cp16(s_act + (size_t)row * KG + E + R + q * 4, a.z + (size_t)(rg0 + row) * R + q * 4);
        }
        pd_cp_wait();
        __syncthreads();
        for (int o = warp; o < 4 * nu; o += PD_NW) {                 // one resident weight row against all staged rows
          float acc[PD_RG];
#pragma unroll
          for (int row = 0; row < PD_RG; ++row) acc[row] = 0.f;
          const float* wr = s_wg + (size_t)o * KG;
          for (int k = lane * 4; k < KG; k += 128) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
            for (int row = 0; row < PD_RG; ++row) {
              if (row < nr) {
                const float4 x4 = *reinterpret_cast<const float4*>(s_act + (size_t)row * KG + k);
                acc[row] = fmaf(w4.x, x4.x, acc[row]); acc[row] = fmaf(w4.y, x4.y, acc[row]);
                acc[row] = fmaf(w4.z, x4.z, acc[row]); acc[row] = fmaf(w4.w, x4.w, acc[row]);
              }
            }
          }
#pragma unroll
          for (int row = 0; row < PD_RG; ++row) {
            const float sgate = pd_warp_sum(acc[row]);
            if (lane == 0 && row < nr) s_G[o * PD_RG + row] = sgate + s_bg[o];
          }
        }
        __syncthreads();
        for (int i = tid; i < nu * nr; i += PD_THREADS) {            // misc/LSTMSoftAttentionCore.py:83-99
          const int ul = i / nr, row = i % nr;
          const float ig = pd_sigmoid(s_G[(ul * 4 + 0) * PD_RG + row]);
          const float fg = pd_sigmoid(s_G[(ul * 4 + 1) * PD_RG + row]);
          const float og = pd_sigmoid(s_G[(ul * 4 + 2) * PD_RG + row]);
          const float gg = tanhf(s_G[(ul * 4 + 3) * PD_RG + row]);
          const float c2 = fg * c_in[(size_t)s_src[row] * R + u0 + ul] + ig * gg;
          c_out[(size_t)(rg0 + row) * R + u0 + ul] = c2;
          h_out[(size_t)(rg0 + row) * R + u0 + ul] = og * tanhf(c2);
        }
        __syncthreads();
      }
    }
    stamp(t, 1);
    grid.sync();
    stamp(t, 2);

    // ===== (D) this CTA's vocabulary slice: logits, per-slice log-softmax statistics and top-k; and the NEXT step's
    //       query projection g = h_2_att_h(h') for its attention columns, from the same staged h' =====
    if (nv > 0 || noa > 0) {
      for (int rg0 = 0; rg0 < rows; rg0 += PD_RG) {
        const int nr = min(PD_RG, rows - rg0);
        for (int i = tid; i < nr * (R / 4); i += PD_THREADS) {
          const int row = i / (R / 4), q = i % (R / 4);
          pd_cp16(s_act + (size_t)row * KG + q * 4, h_out + (size_t)(rg0 + row) * R + q * 4);
        }
        pd_cp_wait();
        __syncthreads();
        // W rows stream from L2, each reused for the nr staged rows.  Latency-bound unless many loads are in flight: a warp
        // owns a contiguous run of columns, takes PD_VC of them per pass and issues all their loads of a 256-float k-chunk
        // (PD_VC x 2 independent 128-bit loads per lane) before the FMAs
        const int cpw = (nv + PD_NW - 1) / PD_NW;                    // columns per warp
        const int wc0 = warp * cpw, wc1 = min(nv, wc0 + cpw);
        for (int vc0 = wc0; vc0 < wc1; vc0 += PD_VC) {
          float acc[PD_VC][PD_RG];
#pragma unroll
          for (int c = 0; c < PD_VC; ++c)
#pragma unroll
            for (int row = 0; row < PD_RG; ++row) acc[c][row] = 0.f;
          for (int k0 = 0; k0 < R; k0 += 256) {
            float4 w4[PD_VC][2];
#pragma unroll
            for (int c = 0; c < PD_VC; ++c) {
              const int vcc = min(vc0 + c, wc1 - 1);                 // clamped: surplus columns are computed and dropped
              const float* wr = a.logit_w + (size_t)(v0 + vcc) * R;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int k = k0 + q * 128 + lane * 4;
                w4[c][q] = (k < R) ? __ldg(reinterpret_cast<const float4*>(wr + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int k = k0 + q * 128 + lane * 4;
              if (k < R) {
#pragma unroll
                for (int row = 0; row < PD_RG; ++row) {
                  if (row < nr) {
                    const float4 x4 = *reinterpret_cast<const float4*>(s_act + (size_t)row * KG + k);
#pragma unroll
                    for (int c = 0; c < PD_VC; ++c) {
                      acc[c][row] = fmaf(w4[c][q].x, x4.x, acc[c][row]); acc[c][row] = fmaf(w4[c][q].y, x4.y, acc[c][row]);
                      acc[c][row] = fmaf(w4[c][q].z, x4.z, acc[c][row]); acc[c][row] = fmaf(w4[c][q].w, x4.w, acc[c][row]);
                    }
                  }
                }
              }
            }
          }
#pragma unroll
          for (int c = 0; c < PD_VC; ++c) {
            const int vc = vc0 + c;
            const float bias = (vc < wc1) ? __ldg(a.logit_b + v0 + vc) : 0.f;
#pragma unroll
            for (int row = 0; row < PD_RG; ++row) {
              const float sl = pd_warp_sum(acc[c][row]);
              if (lane == 0 && row < nr && vc < wc1) s_logit[row * a.VPS + vc] = sl + bias;
            }
          }
        }
        if (t + 1 < a.steps) project_g(rg0, nr, 0);                  // g of the next step, indexed by THIS step's slot order
        __syncthreads();
        for (int row = warp; row < nr && nv > 0; row += PD_NW) {
          const float* x = s_logit + row * a.VPS;
          float m = -INFINITY;
          for (int i = lane; i < nv; i += 32) m = fmaxf(m, x[i]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float se = 0.f;
          for (int i = lane; i < nv; i += 32) se += expf(x[i] - m);
          se = pd_warp_sum(se);
          const size_t pr = (size_t)b * rows + rg0 + row;
          if (lane == 0) { a.part_max[pr] = m; a.part_sum[pr] = se; }
          float pv = INFINITY;
          int pi = -1;
          for (int rnd = 0; rnd < a.ktop; ++rnd) {                   // k rounds of warp arg-best, ties -> lower index
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = lane; i < nv; i += 32) {
              const float v = x[i];
              const int n = v0 + i;
              const bool after = (v < pv) || (v == pv && n > pi);
              if (after && pd_better(v, n, bv, bi)) { bv = v; bi = n; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (pd_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { a.part_val[pr * a.ktop + rnd] = bv; a.part_idx[pr * a.ktop + rnd] = bi; }
            pv = bv; pi = bi;
          }
          if (a.logits)
            for (int i = lane; i < nv; i += 32) a.logits[(size_t)(rg0 + row) * V + v0 + i] = x[i];
        }
        __syncthreads();
      }
    }
    stamp(t, 3);
    grid.sync();
    stamp(t, 4);

    // ===== (E) merge of the slices + token selection; greedy: followed at once by the next step's attention of the row =====
    const int tt = t + 1;                                            // the selection that feeds decoder step tt
    const int groups = a.beam > 0 ? rows / a.beam : rows;            // beam: one CTA per image, greedy: per row
    const int per = a.beam > 0 ? a.beam : 1;
    for (int gi = b; gi < groups; gi += NB) {
      for (int q = warp; q < per; q += PD_NW) {
        const int r = gi * per + q;
        // (all of a lane's slice statistics are loaded before they are used: one L2 round trip instead of three)
        float pm[PD_SL], ps[PD_SL];
#pragma unroll
        for (int i = 0; i < PD_SL; ++i) {
          const int s = lane + 32 * i;
          pm[i] = (s < a.nslice) ? a.part_max[(size_t)s * rows + r] : -INFINITY;
          ps[i] = (s < a.nslice) ? a.part_sum[(size_t)s * rows + r] : 0.f;
        }
        float cv[PD_SL];
        int ci[PD_SL];
        if (a.ktop == 1) {
#pragma unroll
          for (int i = 0; i < PD_SL; ++i) {
            const int s = lane + 32 * i;
            cv[i] = (s < a.nslice) ? a.part_val[(size_t)s * rows + r] : -INFINITY;
            ci[i] = (s < a.nslice) ? a.part_idx[(size_t)s * rows + r] : 0x7fffffff;
          }
        }
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < PD_SL; ++i) m = fmaxf(m, pm[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < PD_SL; ++i)
          if (pm[i] > -INFINITY) sum += ps[i] * expf(pm[i] - m);
        sum = pd_warp_sum(sum);
        const float ls = logf(sum);
        if (lane == 0) { a.rowmax[r] = m; a.logsum[r] = ls; }
        float pv = INFINITY;
        int pi = -1;
        const int ncand = a.nslice * a.ktop;
        for (int rnd = 0; rnd < a.ktop; ++rnd) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          if (a.ktop == 1) {
#pragma unroll
            for (int i = 0; i < PD_SL; ++i)
              if (pd_better(cv[i], ci[i], bv, bi)) { bv = cv[i]; bi = ci[i]; }
          } else {
            for (int c = lane; c < ncand; c += 32) {
              const int s = c / a.ktop, j = c % a.ktop;
              const float v = a.part_val[((size_t)s * rows + r) * a.ktop + j];
              const int i = a.part_idx[((size_t)s * rows + r) * a.ktop + j];
              const bool after = (v < pv) || (v == pv && i > pi);
              if (after && pd_better(v, i, bv, bi)) { bv = v; bi = i; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (pd_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
          }
          if (lane == 0) { s_topv[q * a.ktop + rnd] = (bv - m) - ls; s_topi[q * a.ktop + rnd] = bi; }   // log-prob, torch's association
          pv = bv; pi = bi;
        }
      }
      __syncthreads();
      if (tid == 0 && tt <= L) {
        if (a.beam > 0) {
          // the reference's merge (:465-514) reads top_val[(base + q) * beam + c]: hand it the shared copy, rebased
          const int base = gi * a.beam;
          beam_merge_image(a.bs, gi, tt, s_topv - (size_t)base * a.beam, s_topi - (size_t)base * a.beam, a.src, a.tok);
        } else {
          const int r = gi;
          const int it = s_topi[0];
          const float slp = s_topv[0];
          const bool un = (tt == 1 ? true : a.unfinished[r] != 0) && (it > 0);   // :641-644
          a.unfinished[r] = un ? 1 : 0;
          if (un) atomicOr(&a.any_unfinished[tt], 1);
          a.tok[r] = it;                                                        // embed() sees the UNMASKED token (:637)
          a.seq[(size_t)r * L + (tt - 1)] = un ? (int64_t)it : 0;               // :647
          a.seq_lp[(size_t)r * L + (tt - 1)] = slp;                             // :649
        }
      }
      __syncthreads();
      if (a.beam == 0 && t + 1 < a.steps) attention_row(gi, gi);     // greedy: the row's state stays in its slot
    }
    stamp(t, 5);
    grid.sync();
    stamp(t, 6);
    if (a.beam > 0 && t + 1 < a.steps) {
      // beam: the merge re-ordered the rows; the next step's attention of slot r uses the projection of its source slot
      for (int r = b; r < rows; r += NB) attention_row(r, a.src[r]);
      grid.sync();
    }
    stamp(t, 7);

    // ================= (F) optional: the full log-softmax row of this step (sample()'s logprobs_all) =================
    if (a.lp_all) {
      for (int r = b; r < rows; r += NB) {
        const float sh = a.rowmax[r], ls = a.logsum[r];
        float* dst = a.lp_all + ((size_t)r * (L + 1) + t) * V;
        for (int v = tid; v < V; v += PD_THREADS) dst[v] = (a.logits[(size_t)r * V + v] - sh) - ls;
      }
      // (the next write to a.logits is two grid barriers away)
    }
  }
}

static std::atomic<int> g_pd_enabled{1};
static std::atomic<long long*> g_pd_dbg{nullptr};

struct PDLayout {
  int UPS, OPA, VPS, nslice;
  size_t smem;
};
static PDLayout pd_layout(const rfn_dims& d, int n_sm) {
  PDLayout l{};
  const int R = d.rnn_size, A = d.att_hid_size, E = d.input_encoding_size, V = d.vocab_plus1, S1 = d.num_review_steps;
  l.UPS = (R + n_sm - 1) / n_sm;
  l.OPA = (A + n_sm - 1) / n_sm;
  l.VPS = (V + n_sm - 1) / n_sm;
  l.nslice = (V + l.VPS - 1) / l.VPS;
  const size_t KG = (size_t)E + 2 * R;
  size_t fl = (size_t)4 * l.UPS * KG + (size_t)l.OPA * R + (size_t)PD_RG * KG + (size_t)4 * l.UPS * PD_RG + (size_t)PD_RG * l.VPS +
              4 * l.UPS + l.OPA + std::max(S1, 32) + 2 * RFN_MAX_BEAM * RFN_MAX_BEAM + 2 * PD_RG;
  l.smem = fl * sizeof(float) + 64;
  return l;
}

// scratch the persistent decoder needs beyond DecWork's fields (floats / ints carved by the caller)
size_t pd_part_floats(const rfn_dims& d, int rows, int ktop) {
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const PDLayout l = pd_layout(d, n_sm);
  return (size_t)l.nslice * rows * (2 + 2 * (size_t)std::max(1, ktop));
}

// slices `base` (pd_part_floats values) into the per-slice statistics arrays
void pd_bind_parts(const rfn_dims& d, PDArgs& a, float* base) {
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const PDLayout l = pd_layout(d, n_sm);
  const size_t n = (size_t)l.nslice * a.rows, k = (size_t)std::max(1, a.ktop);
  a.part_max = base;
  a.part_sum = base + n;
  a.part_val = base + 2 * n;
  a.part_idx = reinterpret_cast<int32_t*>(base + 2 * n + n * k);
}

bool pd_supported(const rfn_dims& d, int rows) {
  if (!g_pd_enabled.load() || rows < 1 || rows > PD_MAX_ROWS) return false;
  if (d.decoder_maxout) return false;   // the resident gate rows are laid out for the 4R cell: maxout keeps the per-step kernels
  if (d.rnn_size % 4 || d.att_hid_size % 4 || d.input_encoding_size % 4) return false;
  int dev = 0, n_sm = 0, coop = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop || n_sm < 1) return false;
  return pd_layout(d, n_sm).smem <= (size_t)max_smem;
}

int pd_launch(const rfn_dims& d, PDArgs& a, cudaStream_t st) {
  ProfScope prof__(TAG_MISC, st);
  int dev = 0, n_sm = 0;
  RFN_CUDA(cudaGetDevice(&dev));
  RFN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const PDLayout l = pd_layout(d, n_sm);
  a.UPS = l.UPS; a.OPA = l.OPA; a.VPS = l.VPS; a.nslice = l.nslice;
  a.dbg = g_pd_dbg.load();
  static bool configured = false;
  if (!configured) {
    RFN_CUDA(cudaFuncSetAttribute(decoder_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)n_sm, 1, 1);      // one CTA per SM, all co-resident (cooperative launch)
  cfg.blockDim = dim3(PD_THREADS, 1, 1);
  cfg.dynamicSmemBytes = l.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RFN_CUDA(cudaLaunchKernelEx(&cfg, decoder_persist_kernel, a));
  RFN_LAUNCH_CHECK();
  count_engine(ENG_PERSIST_DECODER);
  return RFN_OK;
}

}  // namespace rfn

extern "C" int rfn_set_persistent_decoder(int on) {
  rfn::g_pd_enabled.store(on ? 1 : 0);
  return RFN_OK;
}
extern "C" int rfn_get_persistent_decoder(void) { return rfn::g_pd_enabled.load(); }
// debugging aid: device buffer of 8 globaltimer stamps (ns) per decoder step written by CTA 0 (NULL = off):
// [0] step start, [1] (C) done, [2] after barrier, [3] (D) done, [4] after barrier, [5] (E) done, [6] after barrier, [7] beam attention done
extern "C" int rfn_debug_set_pd_timeline(long long* d_buf) {
  rfn::g_pd_dbg.store(d_buf);
  return RFN_OK;
}
