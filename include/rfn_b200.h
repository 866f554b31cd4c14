/*
 * rfn_b200.h -- C ABI of the B200-native RFNet recurrent fusion + decode path.
 *
 * The reference (cswhjiang/Recurrent_Fusion_Network) has no FFI: its boundary is the Python
 * nn.Module surface of misc/RecurrentFusionModel.py reached through models.py:14-38.  This
 * library sits directly under that surface; recurrent_fusion_network_b200/model.py binds it with
 * ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* / every `const float*` tensor argument is a DEVICE pointer owned
 *     by the caller.  The path-level and operator-level entry points never allocate device memory
 *     (all scratch comes from the caller's workspace).  Two exceptions, both outside the default
 *     path: rfn_linear_f32_engine with engines 3 / 4 / 5 (an engine-test entry) takes its operand
 *     scratch from the stream-ordered pool (cudaMallocAsync / cudaFreeAsync on `stream`);
 *   - process-wide state the library DOES keep (all of it configuration or lazily created handles,
 *     none of it data): the engine selection (rfn_set_gemm_mode, rfn_set_tc_cluster, rfn_set_h3_cluster,
 *     rfn_set_att_bf16_variant, rfn_set_pdl, rfn_set_splitk, rfn_set_concurrency), per-device side streams + events for the encoder fork / join, the
 *     kernels' one-time cudaFuncSetAttribute flags, the launch / engine counters and the optional
 *     profile records; the error string is thread-local.  Calls on different streams from different
 *     host threads are safe as long as they do not share a workspace; the engine selection is global;
 *   - `params` is a HOST array of device pointers to the model's fp32 tensors in the reference's
 *     state_dict registration order (misc/RecurrentFusionModel.py:153-184; 773 entries for the
 *     five-encoder model), Linear weights (out,in) row-major exactly as torch stores them;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and returns 0 on success or a
 *     negative rfn_status; rfn_last_error() gives the message.  No exceptions cross the ABI;
 *   - row-major contiguous tensors unless a leading dimension is passed.
 */
#ifndef RFN_B200_H
#define RFN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rfn_stream_t; /* cudaStream_t */

enum rfn_status {
  RFN_OK = 0,
  RFN_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  RFN_ERR_CUDA = -2,        /* a CUDA call / launch failed */
  RFN_ERR_WORKSPACE = -3,   /* workspace too small */
  RFN_ERR_UNSUPPORTED = -4  /* configuration outside what the kernels support */
};

#define RFN_MAX_ENCODERS 8
#define RFN_MAX_BEAM 16   /* the reference's default beam_size is 10 (misc/RecurrentFusionModel.py:353) */

/* The `opt` fields RecurrentFusionModel reads (misc/RecurrentFusionModel.py:120-151). */
typedef struct rfn_dims {
  int32_t J;                              /* number of CNN encoders (feat_array_info entries) */
  int32_t att_num[RFN_MAX_ENCODERS];      /* N_j */
  int32_t att_feat_size[RFN_MAX_ENCODERS];/* D_j */
  int32_t fc_feat_size[RFN_MAX_ENCODERS]; /* F_j */
  int32_t rnn_size;                       /* R */
  int32_t att_hid_size;                   /* A */
  int32_t input_encoding_size;            /* E */
  int32_t vocab_plus1;                    /* V1 = vocab_size + 1 */
  int32_t top_words_count;                /* K */
  int32_t num_review_steps_0;             /* S0: stage-1 fusion steps */
  int32_t num_review_steps;               /* S1: stage-2 review steps */
  int32_t seq_length;                     /* L */
  int32_t review_maxout;                  /* opt.review_maxout: stage-2 cells are 5R wide, in_transform = max of the last two R-blocks
                                           * instead of tanh (misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:25,60) */
  int32_t decoder_maxout;                 /* opt.maxout: the same for the decoder cell (misc/LSTMSoftAttentionCore.py:25,89) */
} rfn_dims;

/* ---- library -------------------------------------------------------------------------- */
const char* rfn_last_error(void);
int rfn_version(void);
/* 0 if the current device is sm_100 (B200); RFN_ERR_UNSUPPORTED otherwise. */
int rfn_check_device(void);
/* number of entries `params` must hold for these dims (773 for the full model) */
int rfn_num_params(const rfn_dims* dims);
/* Weight cache of the split-fp16 / bf16 engine (engine modes 4 / 5): the split of every weight matrix the path's large
 * GEMMs consume (rfn_split_rows_f32 applied per GEMM, joint row scale over the sources of one GEMM), built once per
 * weight version into a caller-owned buffer of rfn_wcache_bytes(dims, bf16) bytes (~ the size of the fp32 model).
 * Every path-level entry point below takes `params` with rfn_num_param_slots(dims) = rfn_num_params(dims) + 1 entries:
 * the LAST entry is the device pointer of that buffer, or NULL to split the weights on the fly in the workspace (about
 * 2 x the model size of extra HBM traffic and ~300 extra launches per call).  A cache built with bf16 = 0 serves mode 4,
 * bf16 = 1 serves mode 5; it must be rebuilt after the weights change (optimizer.step(), load_state_dict). */
int rfn_num_param_slots(const rfn_dims* dims);
size_t rfn_wcache_bytes(const rfn_dims* dims, int bf16);
int rfn_wcache_build(const rfn_dims* dims, const float* const* params, int bf16, void* wcache, size_t wcache_bytes,
                     rfn_stream_t stream);
/* number of kernels this library has launched in the calling process (bench accounting) */
uint64_t rfn_launch_count(void);
/* launches per GEMM kernel family since the process started (rfn_engine_num() families, named by rfn_engine_name()):
 * the tests and smoke() use it to prove WHICH kernel instantiation a configuration actually ran */
int rfn_engine_num(void);
const char* rfn_engine_name(int engine);
int rfn_engine_launch_counts(uint64_t* out, int n);
/* selects the contraction engine for GEMMs with >= 128 rows: 0 = fp32 SIMT FMA; 1 = tcgen05 3xTF32 with chunked
 * round-to-nearest accumulation (fp32-equivalent); 2 = tcgen05 single-pass TF32 (reduced precision); 3 (the default) =
 * mode 1, except that the two fused-epilogue GEMMs of the path (attention projection, logits) compute the cross terms
 * x_lo.w and x.w_lo of the split product as BF16 MMAs (8 instead of 12 tensor-core units per k-block; measured error
 * 1.5e-6 rms against 1.3e-6 for mode 1, same captions as mode 1 on 4,998 of 5,000 bench images).  Problems with fewer
 * rows always use the SIMT kernel on the decode path. */
int rfn_set_gemm_mode(int mode);
int rfn_get_gemm_mode(void);
/* 1: every GEMM of the following calls may take the split-K route (partial tiles summed with atomic adds: the summation
 * order, hence the last bits of the result, vary from run to run).  Default 0: only the training operators ask for it
 * (RFN_GEMM_SPLITK); training.rl_forward_loss turns it on around its two no-tape decodes of a few hundred rows, whose
 * tokens are samples / a baseline, not outputs that must reproduce bit for bit. */
int rfn_set_splitk(int on);
int rfn_get_splitk(void);
/* Tensor-engine GEMMs with >= 256 rows and columns: 0 = one CTA per 128 x 256 tile; 1 = 2-CTA clusters
 * (tcgen05 cta_group::2, 256 x 256 tiles, operands split across the pair), one tile per cluster;
 * 2 = persistent 2-CTA clusters looping over tiles with the epilogue overlapped (3xTF32 mode). */
int rfn_set_tc_cluster(int on);
int rfn_get_tc_cluster(void);
/* Split-fp16 / bf16 engine (modes 4 / 5): CTAs per cluster.  2 (default) = one tcgen05 cta_group::2 pair per 256 x 256 tile;
 * 4 = two pairs on vertically adjacent tiles that share the W tile through TMA multicast (each CTA fetches half of its W rows):
 * 25 % less L2 -> shared-memory traffic per MMA, for GEMMs with more than 256 rows. */
int rfn_set_h3_cluster(int ctas);
int rfn_get_h3_cluster(void);
/* Engine mode 5: load scheme of the attention context sum over the bf16 feature copy (z = sum_n a_n A_n,
 * misc/AttentionModelCore.py:45-47).  0 = 8-byte loads, 1 = 16-byte loads, 2 (default) = 16-byte loads in batches of eight
 * locations with D split across CTAs instead of a pass over a few leftover columns.  Same results in all three. */
int rfn_set_att_bf16_variant(int variant);
int rfn_get_att_bf16_variant(void);
/* Debugging aid: device buffer receiving 8 clock64() stamps per CTA of the 2-CTA GEMM kernel
 * (start, init done, first MMA, last MMA, last drain, epilogue done, exit); NULL switches it off. */
int rfn_debug_set_timeline(long long* d_buf, int epilogue_kind /* -1 all, 0 store, 1 score, 2 vocab */);

/* 1: the frequent kernels of the decode path are launched with programmatic dependent launch (their launch latency and
 * prologue overlap the tail of the preceding kernel of the stream; each waits with griddepcontrol.wait before touching
 * global memory); 0 (default): plain stream order.  Results are bit-identical; measured gain on B200: none within noise
 * (20.34 vs 20.44 ms per 625-image step), the path is not bound by launch gaps once it is replayed from a CUDA graph. */
int rfn_set_pdl(int on);
int rfn_get_pdl(void);

/* 1 (default): decodes with at most 64 decoder rows (greedy batches, beam x few images) run the whole timestep loop of the
 * decoder LSTM as ONE cooperative launch (one CTA per SM, the gate weights of each CTA's hidden units resident in shared
 * memory for all timesteps, five grid barriers per step); 0: always the per-step launch sequence. */
int rfn_set_persistent_decoder(int on);
int rfn_get_persistent_decoder(void);
/* Debugging aid: device buffer receiving 8 globaltimer stamps (ns) per decoder step from CTA 0 of the persistent decoder
 * (step start, phase A done, after its barrier, B done, C done, D done, after its barrier, E done); NULL switches it off. */
int rfn_debug_set_pd_timeline(long long* d_buf);

/* 1 (default): the J independent encoder cells of a fusion step run on internal side streams forked
 * from / joined into the caller's stream; 0: everything is serialised on the caller's stream. */
int rfn_set_concurrency(int on);

/* Optional per-kernel-class timing: when enabled every launch is bracketed by CUDA events on its
 * stream; rfn_profile_read() synchronises them and returns summed milliseconds and launch counts
 * per class (rfn_profile_num_tags() classes, named by rfn_profile_tag_name()), then clears. */
int rfn_profile_enable(int on);
int rfn_profile_num_tags(void);
const char* rfn_profile_tag_name(int tag);
int rfn_profile_read(float* ms, uint64_t* launches, int n);

/* ---- operator level (the per-timestep cores' building blocks) --------------------------- */

/* y[M,N] = (accumulate ? y : 0) + sum_i x_i[M,K_i] . W_i[N,K_i]^T + sum_i bias_i[N]
 * replaces nn.Linear at every call site of the path, e.g. H2h(H) + z2h(z)
 * (misc/RecurrentFusionModel.py:53), i2h + h2h + z2h (misc/LSTMSoftAttentionCore.py:81).
 * n_src in 1..3; bias_i may be NULL; K_i % 4 == 0 and 16-byte aligned rows required.
 * accumulate is a flag word: bit 0 = add into y; RFN_GEMM_SPLITK = the tensor engine may split a long contraction
 * over clusters and add the partial tiles atomically (summation order then varies from run to run: used for the
 * weight gradients dU = dP^T . A only, never on the decode path). */
#define RFN_GEMM_SPLITK 2
int rfn_linear_f32(int n_src, const float* const* x, const int* ldx, const float* const* W,
                   const int* K, const float* const* bias, float* y, int ldy, int M, int N,
                   int accumulate, rfn_stream_t stream);

/* Same contraction on an explicitly chosen engine (0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05
 * single-pass TF32) regardless of the row-count heuristic; used by the engine parity tests. */
int rfn_linear_f32_engine(int engine, int n_src, const float* const* x, const int* ldx,
                          const float* const* W, const int* K, const float* const* bias, float* y,
                          int ldy, int M, int N, int accumulate, rfn_stream_t stream);

/* Split-fp16 tensor engine (engine mode 4, the default; mode 5 = its single-pass bf16 sibling), operator level.
 * rfn_split_rows_f32 turns fp32 rows (n_src matrices that share the row index, e.g. the concatenated inputs H | z of
 * H2h(H) + z2h(z), misc/RecurrentFusionModel.py:53, or the rows of their weight matrices) into a power-of-two row scale
 * and two fp16 pieces per element, x = (x0 + x1) / s with 22 mantissa bits (bf16 != 0: one bf16 piece, no scale);
 * `out` needs rfn_split_bytes(rows, n_src, K, bf16) bytes.  rfn_linear_split then computes the same contraction as
 * rfn_linear_f32 from two such buffers (x: M rows, W: N rows) as x0.w0 + x1.w0 + x0.w1 on the tcgen05 tensor cores
 * (three kind::f16 MMAs per product, fp32 accumulation drained into round-to-nearest registers); M, N >= 256. */
size_t rfn_split_bytes(int rows, int n_src, const int* K, int bf16);
int rfn_split_rows_f32(int n_src, const float* const* x, const int* ldx, const int* K, int rows, int bf16,
                       void* out, size_t out_bytes, rfn_stream_t stream);
int rfn_linear_split(int bf16, int n_src, const void* x_split, const void* w_split, const int* K,
                     const float* const* bias, float* y, int ldy, int M, int N, int accumulate,
                     rfn_stream_t stream);

/* Additive soft attention given P = att_2_att_h(A) (misc/AttentionModelCore.py:36-47):
 *   e[r,n] = w . tanh(P[r/div, n, :] + g[r, :]) + wb ; a = softmax_n(e) ; z[r,:] = sum_n a[r,n] A[r/div,n,:]
 * A (rowsA,N,D), P (rowsA,N,Ah), g (rows,Ah) = h_2_att_h(h), w (Ah), d_wb device scalar,
 * z (rows, ldz), alpha (rows,N) optional.  `div` lets `div` consecutive query rows share one
 * feature row (beam rows of one image); rowsA = ceil(rows/div). */
int rfn_attention_step_f32(const float* A, const float* P, const float* g, const float* w,
                           const float* d_wb, float* z, int ldz, float* alpha, int rows, int N,
                           int D, int Ah, int div, rfn_stream_t stream);

/* Full AttentionModelCore.forward(pre_h, att_seq) (misc/AttentionModelCore.py:31-48) from the six
 * module tensors.  workspace: rows*N*Ah + rows*Ah floats. */
int rfn_attention_core_f32(const float* h, const float* A, const float* U_w, const float* U_b,
                           const float* Wh_w, const float* Wh_b, const float* v_w,
                           const float* d_v_b, float* z, float* alpha, int rows, int N, int D,
                           int R, int Ah, void* workspace, size_t workspace_bytes,
                           rfn_stream_t stream);

/* LSTM update, gate layout [i|f|o|g] along 4R (misc/RecurrentFusionModel.py:55-73):
 * c' = sig(f) c + sig(i) tanh(g); h' = sig(o) tanh(c').  h_out2 (optional) receives a second
 * copy of h' with leading dimension ldh2 (the thought-vector slot). */
int rfn_lstm_cell_f32(const float* G, const float* c_prev, float* h_out, float* c_out,
                      float* h_out2, int ldh2, int rows, int R, rfn_stream_t stream);

/* The same with nn.Dropout applied to h' (training, misc/LSTMSoftAttentionCore.py:98-100: the dropped h' is both the output
 * and the carried state): h' = scale * mask * sig(o) tanh(c'), mask (rows,R) of 0/1 (NULL: no dropout). */
int rfn_lstm_cell_drop_f32(const float* G, const float* c_prev, const float* mask, float scale, float* h_out,
                           float* c_out, float* h_out2, int ldh2, float* h_out3, int ldh3, int rows, int R,
                           rfn_stream_t stream);
/* The maxout variant of the cell (opt.maxout / opt.review_maxout = 1): G is (rows, 5R) = [i|f|o|g1|g2] and the input
 * transform is max(g1, g2) with no tanh (misc/LSTMSoftAttentionCore.py:89-91); maxout = 0 is rfn_lstm_cell_drop_f32.  Backward:
 * the gradient of the transform goes to the larger of g1 / g2 (to both halves when they are equal, as torch.max does). */
int rfn_lstm_cell_ex_f32(const float* G, const float* c_prev, const float* mask, float scale, int maxout, float* h_out,
                         float* c_out, float* h_out2, int ldh2, float* h_out3, int ldh3, int rows, int R,
                         rfn_stream_t stream);
int rfn_lstm_cell_bwd_ex_f32(const float* G, const float* c_prev, int n_dh, const float* const* dh, const int* ld_dh,
                             const float* mask, float scale, int maxout, const float* dc_next, float* dG, float* dc_prev,
                             int rows, int R, rfn_stream_t stream);
/* out[r,:] = alpha * sum_i x_i[r,:] over n <= 8 strided (rows, R) sources, and
 * out[r,:] = (((0 + in_0[r,:]) + in_1[r,:]) + ...) / n with in_j = in + j*stride (the stage-1 -> stage-2 bridge,
 * misc/RecurrentFusionModel.py:307-309, in Python's summation order). */
int rfn_sum_strided_f32(int n, const float* const* x, const int* ldx, float alpha, float* out, int ldo, int rows, int R,
                        rfn_stream_t stream);
int rfn_mean_tensors_f32(const float* in, size_t stride, int n, float* out, int ld_out, int rows, int R, int ld_in,
                         rfn_stream_t stream);

/* lp[r,:] = log_softmax(logits[r,:]) (F.log_softmax, misc/RecurrentFusionModel.py:278). */
int rfn_log_softmax_f32(const float* logits, int ld_in, float* lp, int ld_out, int rows, int V,
                        rfn_stream_t stream);

/* ---- path level --------------------------------------------------------------------------- */

/* Bytes of scratch the path-level calls need for `rows` feature rows and `dec_rows` decoder rows
 * (dec_rows = rows for greedy/sample/teacher forcing, rows*beam for beam search). */
size_t rfn_workspace_bytes(const rfn_dims* dims, int rows, int dec_rows);

/* scratch for rfn_ensemble_decode_beam */
size_t rfn_ensemble_workspace_bytes(const rfn_dims* dims, int n_models, int images, int beam);

/* get_init_state + get_thought_vectors (misc/RecurrentFusionModel.py:333-343, :283-331):
 * fc[j] (rows,F_j), att[j] (rows,N_j,D_j) ->
 *   TVc (rows,S1,R), state h/c (rows,R) = final stage-2 state,
 *   TV (J,rows,S0,R) optional, reason_pred (J+1,rows,K) optional.
 * If fc is NULL the initial states are taken from init_h[j] / init_c[j] (rows,R) instead of
 * fc2h_j(fc_j) -- the state_list a caller got from get_init_state. */
int rfn_thought_vectors(const rfn_dims* dims, const float* const* params, const float* const* fc,
                        const float* const* init_h, const float* const* init_c,
                        const float* const* att, int rows, float* TVc, float* h_out, float* c_out,
                        float* TV, float* reason_pred, void* workspace, size_t workspace_bytes,
                        rfn_stream_t stream);

/* one_time_step (misc/RecurrentFusionModel.py:345-350; misc/LSTMSoftAttentionCore.py:60-102):
 * xt (rows,E), TVc (rows/div,S1,R), state in/out (rows,R) -> logits (rows,V1) (not log-probs). */
int rfn_one_time_step(const rfn_dims* dims, const float* const* params, const float* xt,
                      const float* TVc, int div, const float* h_in, const float* c_in,
                      float* h_out, float* c_out, float* logits, int rows, void* workspace,
                      size_t workspace_bytes, rfn_stream_t stream);

/* Teacher-forced decoder (forward's loop, misc/RecurrentFusionModel.py:259-279) for T steps:
 * seq (rows, ld_seq) int64, feeds columns 0..T-1 -> logprobs (rows,T,V1).  The caller derives T
 * from the first all-zero column (:274-275). */
int rfn_decode_teacher_forced(const rfn_dims* dims, const float* const* params, const float* TVc,
                              const float* h0, const float* c0, const int64_t* seq, int ld_seq,
                              int T, int rows, float* logprobs, void* workspace,
                              size_t workspace_bytes, rfn_stream_t stream);

/* Greedy (uniforms == NULL) or multinomial sample() loop (misc/RecurrentFusionModel.py:616-658).
 * Runs all L+1 steps on the device; outputs are (rows,L) / (rows,L+1,V1) buffers and
 * d_T (device int32) receives the number of token columns the reference would have produced
 * (its early `break`, :645); the caller slices [:T] / [:T+1].  lp_all may be NULL.
 * uniforms (rows,L) in [0,1) drive inverse-CDF sampling of exp(lp/temperature) in index order
 * (the reference draws on the CPU RNG, :624-631, which cannot be replayed -- SURVEY D8). */
int rfn_decode_sample(const rfn_dims* dims, const float* const* params, const float* TVc,
                      const float* h0, const float* c0, int rows, const float* uniforms,
                      float temperature, int64_t* seq, float* seq_logprobs, float* lp_all,
                      int32_t* d_T, void* workspace, size_t workspace_bytes, rfn_stream_t stream);

/* Batched sample_beam (misc/RecurrentFusionModel.py:352-543) over `images` images at once,
 * beam rows expanded on the device (stage 1-2 outputs are shared by an image's beams).
 *   seq (images,L) int64, seq_logprobs (images,L): best finished beam per image (:529-531);
 *   done_seq (images,cap,L) int32, done_logps (images,cap,L), done_p (images,cap) and
 *   n_done (images) list every finished beam sorted by -p (top_seq / top_prob), cap = beam*L. */
int rfn_decode_beam(const rfn_dims* dims, const float* const* params, const float* TVc,
                    const float* h0, const float* c0, int images, int beam, int64_t* seq,
                    float* seq_logprobs, int32_t* done_seq, float* done_logps, float* done_p,
                    int32_t* n_done, void* workspace, size_t workspace_bytes,
                    rfn_stream_t stream);

/* Ensemble beam search (eval_utils.py:268-290 logit mean -> log_softmax; :482-658 beam loop) over
 * n_models weight sets sharing one beam; params_m[m], TVc_m[m], h0_m[m], c0_m[m] per model. */
int rfn_ensemble_decode_beam(const rfn_dims* dims, int n_models, const float* const* const* params_m,
                             const float* const* TVc_m, const float* const* h0_m,
                             const float* const* c0_m, int images, int beam, int64_t* seq,
                             float* seq_logprobs, int32_t* done_seq, float* done_logps,
                             float* done_p, int32_t* n_done, void* workspace,
                             size_t workspace_bytes, rfn_stream_t stream);

/* Ensemble greedy decode (eval_utils.py:729-975, the loop eval_ensemble.sh runs with --beam_size 1): `rows` images advance
 * together, per step argmax of log_softmax(mean_m logit_m); every model embeds the UNMASKED argmax, finished rows write 0.
 * seq (rows,L) int64, seq_logprobs (rows,L); *d_T = number of valid columns (the reference's early break, :889-891).
 * Workspace: rfn_ensemble_workspace_bytes(dims, n_models, rows, 1). */
int rfn_ensemble_decode_greedy(const rfn_dims* dims, int n_models, const float* const* const* params_m,
                               const float* const* TVc_m, const float* const* h0_m,
                               const float* const* c0_m, int rows, int64_t* seq, float* seq_logprobs,
                               int32_t* d_T, void* workspace, size_t workspace_bytes,
                               rfn_stream_t stream);

/* ReviewNetEnsembleCriterion's sequence term (misc/utils.py:161-184), fused over the vocab:
 * out[0] = -(1/rows) sum_{b,t} mask[b,t] ((1-eps) lp[b,t,y] + eps/V1 sum_v lp[b,t,v]).
 * target (rows, ld_t) int64, mask (rows, ld_t) f32; only the first T columns are read. */
int rfn_xe_loss_f32(const float* logprobs, const int64_t* target, const float* mask, int ld_t,
                    int rows, int T, int V, float eps, float* out, rfn_stream_t stream);

/* The same with lp[b,t,:] at logprobs + b*ld_b + t*ld_s (floats): the time-major (T, rows, V) log-prob table of the fused
 * training tape (tape.py) is read in place, ld_b = V, ld_s = rows*V. */
int rfn_xe_loss_strided_f32(const float* logprobs, size_t ld_b, size_t ld_s, const int64_t* target, const float* mask,
                            int ld_t, int rows, int T, int V, float eps, float* out, rfn_stream_t stream);

/* ReviewNetRewardCriterion's sequence + entropy terms, non-PPO (misc/utils.py:50-72):
 * out[0] = -(1/rows) sum m l R + (entropy_reg/rows) sum_{b,t} m0 sum_v p log p. */
int rfn_rl_loss_f32(const float* sample_logprobs, const int64_t* seq, const float* reward,
                    const float* logprobs_all, int ld_lp_rows, int rows, int T, int V,
                    float entropy_reg, float* out, rfn_stream_t stream);

int rfn_rl_loss_strided_f32(const float* sample_logprobs, const int64_t* seq, const float* reward,
                            const float* logprobs_all, size_t ld_b, size_t ld_s, int rows, int T, int V,
                            float entropy_reg, float* out, rfn_stream_t stream);

/* nn.MultiLabelMarginLoss (mean reduction) scaled by `weight`, the discriminative term of both
 * criteria (misc/utils.py:76-82, :186-190): out[0] (+)= weight * mean_rows(margin loss).
 * pred (rows,K) f32, target (rows,K) int64, class ids first, -1 terminated. */
int rfn_multilabel_margin_f32(const float* pred, const int64_t* target, int rows, int K, float weight,
                              int accumulate, float* out, rfn_stream_t stream);

/* model_ensemble_feat_array_one_step's tail (eval_utils.py:282-288):
 * lp = log_softmax((((0 + l_0) + l_1) + ...) / n) over n <= 8 logit tensors (rows,V). */
int rfn_mean_log_softmax_f32(int n, const float* const* logits, int rows, int V, float* mean_scratch,
                             float* lp, rfn_stream_t stream);

/* ---- backward operators (XE / RL training: SURVEY.md 8a rows a7, a8, a12) ------------------------
 * Called by recurrent_fusion_network_b200/autograd.py, whose torch.autograd.Function wrappers
 * stand where the reference relies on autograd through nn.Linear / tanh / softmax / bmm. */

/* C[M,N] (+)= op(A)[M,K] . op(B)[K,N]; a_kmajor: A stored (M,K) row-major else (K,M); b_kmajor: B stored
 * (N,K) row-major else (K,N).  dX = dY . W is (1,0); dW += dY^T . X is (0,0). */
int rfn_gemm_general_f32(int a_kmajor, int b_kmajor, const float* A, int lda, const float* B, int ldb,
                         float* C, int ldc, int M, int N, int K, int accumulate, rfn_stream_t stream);
/* Same contraction on an explicitly chosen engine (0 = fp32 SIMT; 1 / 2 = tcgen05 3xTF32 / TF32, layout (1,0) only: the B
 * operand is read MN-major, no transposed copy of the weights); used by the engine parity tests. */
int rfn_gemm_general_f32_engine(int engine, int a_kmajor, int b_kmajor, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, int M, int N, int K, int accumulate, rfn_stream_t stream);
/* dst[c, r] = src[r, c] for r < rows, c < cols; dst rows are ld_dst long and columns rows..ld_dst-1 are zero-filled
 * (ld_dst = rows rounded up to a multiple of 4).  Brings dP and A of dU = dP^T . A into the K-major layout
 * of the tensor engine: dU = rfn_linear_f32(x = dP^T, W = A^T). */
int rfn_transpose_f32(const float* src, int ld_src, int rows, int cols, float* dst, int ld_dst, rfn_stream_t stream);
/* db[n] (+)= sum_m dY[m,n]  (bias gradient) */
int rfn_colsum_f32(const float* dY, int ld, int M, int N, float* db, int accumulate, rfn_stream_t stream);
/* backward of rfn_attention_step_f32: from dz and the saved alpha -> dP (rows,N,Ah), dg (rows,Ah),
 * dw (Ah, +=), dwb (1, +=), and dA (rowsA,N,D, +=) when the attended set needs a gradient (thought vectors). */
int rfn_attention_step_bwd_f32(const float* A, const float* P, const float* g, const float* w,
                               const float* alpha, const float* dz, int lddz, float* dP, float* dg,
                               float* dw, float* dwb, float* dA, int rows, int N, int D, int Ah, int div,
                               rfn_stream_t stream);
/* backward of rfn_lstm_cell_f32: (dh, dc_next may be NULL) -> dG (rows,4R), dc_prev (rows,R) */
int rfn_lstm_cell_bwd_f32(const float* G, const float* c_prev, const float* dh, const float* dc_next,
                          float* dG, float* dc_prev, int rows, int R, rfn_stream_t stream);
/* The same with dh = sum of n_dh (<= 8) strided sources dh[i] (rows, R, leading dimension ld_dh[i]) and an optional dropout
 * keep-mask on h (dh_raw = scale * mask * dh): the hand-scheduled backward of tape.py feeds the gradient pieces of h (the
 * thought-vector slot, the slice of dH from the next fusion step, the next step's query gradient) without add kernels. */
int rfn_lstm_cell_bwd_multi_f32(const float* G, const float* c_prev, int n_dh, const float* const* dh, const int* ld_dh,
                                const float* mask, float scale, const float* dc_next, float* dG, float* dc_prev,
                                int rows, int R, rfn_stream_t stream);
/* dX[M,K] (+)= sum_i dY_i[M,N_i] . W_i[N_i,K], 1 <= n_src <= 3: the input gradient of y = sum_i x_i W_i^T (nn.Linear weights
 * (out, in) are read as they lie, MN-major on the tensor engine), and with several dY_i the sum over the consumers of one
 * input -- dH of a fusion step, misc/RecurrentFusionModel.py:53,102-107 -- in one launch. */
int rfn_linear_bwd_x_f32(int n_src, const float* const* dY, const int* lddy, const float* const* W, const int* ldw,
                         const int* Nc, float* dX, int lddx, int M, int K, int accumulate, rfn_stream_t stream);
/* dx = dlp - exp(lp) * sum_v dlp */
int rfn_log_softmax_bwd_f32(const float* lp, size_t ld_lp, const float* dlp, size_t ld_d, float* dx,
                            size_t ld_x, int rows, int V, rfn_stream_t stream);
/* x[r,:] = embed[tok[r*ld_tok], :]  (nn.Embedding, misc/RecurrentFusionModel.py:276) and its backward */
int rfn_embed_f32(const int64_t* tok, int ld_tok, const float* embed, float* x, int rows, int E, int V1,
                  rfn_stream_t stream);
int rfn_embed_bwd_f32(const int64_t* tok, int ld_tok, const float* dx, float* dE, int rows, int E, int V1,
                      rfn_stream_t stream);
/* out[r,k] = max_s in[r,s,k] (torch.max(reason_mat, 1), misc/RecurrentFusionModel.py:303) and its backward */
int rfn_max_over_steps_f32(const float* in, float* out, int rows, int S, int K, rfn_stream_t stream);
int rfn_max_over_steps_bwd_f32(const float* in, const float* dout, float* din, int rows, int S, int K,
                               rfn_stream_t stream);
/* out = alpha * x + beta * y  (y may be NULL) */
int rfn_axpby_f32(float alpha, const float* x, float beta, const float* y, float* out, size_t n,
                  rfn_stream_t stream);
/* out = scale * x * m  (nn.Dropout with an explicit keep-mask m, scale = 1/(1-p)) */
int rfn_mul_scale_f32(float scale, const float* x, const float* m, float* out, size_t n, rfn_stream_t stream);
/* one token per log-prob row: arg-max (uniforms NULL; ties -> lower index, torch.max) or the inverse CDF of
 * exp(lp/temperature) against uniforms[r] (misc/RecurrentFusionModel.py:619-635); lp_out = lp[r, tok]. */
int rfn_select_token_f32(const float* lp, size_t ld, int rows, int V, const float* uniforms, float temperature,
                         int64_t* tok, float* lp_out, rfn_stream_t stream);
/* out[r] = x[r, idx[r]] (logprobs.gather, :632) and its backward (dx zero except column idx[r]) */
int rfn_gather_cols_f32(const float* x, size_t ld, const int64_t* idx, float* out, int rows, rfn_stream_t stream);
int rfn_scatter_cols_f32(const float* dout, const int64_t* idx, float* dx, size_t ld, int rows, int V,
                         rfn_stream_t stream);
/* out[r,:] = x[r / g, :] (each unique image row expanded to the g = seq_per_img replicas the loader feeds,
 * dataloader.py:251-252) and its backward out[r,:] = sum_{i<g} x[r*g + i, :] */
int rfn_expand_rows_f32(const float* x, int g, float* out, int rows_out, int R, rfn_stream_t stream);
int rfn_group_sum_f32(const float* x, int g, float* out, int rows_out, int R, rfn_stream_t stream);
/* gradients of the criteria w.r.t. their log-prob inputs (gout: device scalar, upstream gradient) */
int rfn_xe_loss_bwd_f32(const int64_t* target, const float* mask, int ld_t, int rows, int T, int V, float eps,
                        const float* gout, float* dlp, rfn_stream_t stream);
int rfn_rl_loss_bwd_f32(const int64_t* seq, const float* reward, const float* lp_all, int ld_lp_rows, int rows,
                        int T, int T1, int V, float entropy_reg, const float* gout, float* dslp,
                        float* dlp_all, rfn_stream_t stream);
int rfn_multilabel_margin_bwd_f32(const float* pred, const int64_t* target, int rows, int K, float weight,
                                  const float* gout, float* dx, rfn_stream_t stream);
/* strided variants (dlp[b,t,:] at dlp + b*ld_b + t*ld_s; likewise lp_all / dlp_all), see rfn_xe_loss_strided_f32 */
int rfn_xe_loss_bwd_strided_f32(const int64_t* target, const float* mask, int ld_t, int rows, int T, int V, float eps,
                                const float* gout, float* dlp, size_t ld_b, size_t ld_s, rfn_stream_t stream);
int rfn_rl_loss_bwd_strided_f32(const int64_t* seq, const float* reward, const float* lp_all, size_t ld_b, size_t ld_s,
                                int rows, int T, int T1, int V, float entropy_reg, const float* gout, float* dslp,
                                float* dlp_all, rfn_stream_t stream);

/* Fused clip_gradient (element-wise clamp to +-grad_clip, misc/utils.py:292-296; <= 0 disables) + Adam step with
 * L2 weight decay (torch.optim.Adam semantics, train.py:56,160-163) over n_tensors parameter tensors; `step` is the
 * 1-based update count (bias correction).  grad_scale multiplies the gradient first (1 / world_size turns the sum of a
 * data-parallel all-reduce into the mean without another pass; 1 otherwise).  HOST arrays of device pointers / element
 * counts.  d_hyper (nullable):
 * device float[2] = {step, lr} read by the kernel instead of the two host arguments, so that a captured CUDA graph
 * of the training step follows the update count and the learning-rate schedule. */
int rfn_adam_step_f32(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                      const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay,
                      float grad_clip, float grad_scale, int step, const float* d_hyper, rfn_stream_t stream);

/* ---- CIDEr-D reward scorer (cider/pyciderevalcap/ciderD/ciderD_scorer.py:114-199 as driven by
 * get_rewards.py:39-112).  Captions are int32 token rows (the tokens up to and including the first 0 form the
 * caption, get_rewards.py:20-27; at most 32 tokens).  hyp (n_hyp, ld_h); hyp_img[n_hyp] = image of each hypothesis;
 * refs (n_img, R, Lr) with n_refs[n_img] valid references per image.  Document frequencies: open-addressing table
 * df_keys (cap, 4) int32 (n-gram tokens, unused = -1, empty slot = INT32_MIN in column 0), df_vals (cap) double, cap
 * a power of two (0 = empty table), slot = rfn_ciderd_hash(key) & (cap-1) with linear probing.  ref_len = log(#docs).
 * sims_scratch: n_hyp*R*4 doubles.  scores (n_hyp) double = CIDEr-D x 10 per hypothesis, fp64 throughout. */
int rfn_ciderd_scores_f64(const int32_t* hyp, int ld_h, int n_hyp, const int32_t* hyp_img, const int32_t* refs,
                          const int32_t* n_refs, int R, int Lr, const int32_t* df_keys, const double* df_vals,
                          int df_cap, double ref_len, double sigma, double* sims_scratch, double* scores,
                          rfn_stream_t stream);
/* reward[b, t] = weight * (scores[b] - scores[rows + b]) (use_baseline) or weight * scores[b], broadcast over T
 * (get_rewards.py:96-110) */
int rfn_ciderd_reward_f32(const double* scores, int rows, int T, double weight, int use_baseline, float* reward,
                          rfn_stream_t stream);
uint64_t rfn_ciderd_hash(const int32_t* key4);

#ifdef __cplusplus
}
#endif
#endif /* RFN_B200_H */
