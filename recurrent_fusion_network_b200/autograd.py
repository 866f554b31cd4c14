"""torch.autograd.Function wrappers around the forward and backward kernels of librfn_b200.so.

The reference trains through autograd over nn.Linear / tanh / softmax / bmm / log_softmax
(train.py:154-163, train_rl.py:160-191).  Here autograd is only the tape: every forward and every
backward below is one or a few calls into the C ABI (include/rfn_b200.h); no arithmetic is done with torch
ops.  Gradients are checked against autograd through the oracle in tests/test_gpu_training.py."""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._capi import check, lib, ptr, ptr_array, stream


def _c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _dense3(t):
    """A (rows, T, V) fp32 table as the criteria kernels address it: unit stride over V and either the contiguous layout or
    its time-major transpose (what tape.DecoderFn produces); anything else is copied."""
    if t.dtype == torch.float32 and t.dim() == 3 and t.stride(2) == 1:
        rows, T, V = t.shape
        if (t.stride(0), t.stride(1)) in ((T * V, V), (V, rows * V)):
            return t
    return t.float().contiguous()


def _gemm_general(a_k, b_k, A, lda, B, ldb, Cm, ldc, M, N, K, accumulate=0):
    check(lib().rfn_gemm_general_f32(int(a_k), int(b_k), ptr(A), lda, ptr(B), ldb, ptr(Cm), ldc, M, N, K, int(accumulate),
                                     stream()), "rfn_gemm_general_f32")


def _linear_fwd(xs, Ws, bs, rows, out_features):
    y = torch.empty(rows, out_features, dtype=torch.float32, device=xs[0].device)
    n, done = len(xs), 0
    while done < n:
        grp = list(range(done, min(n, done + 3)))
        ld = (C.c_int * len(grp))(*[xs[i].stride(0) for i in grp])
        ks = (C.c_int * len(grp))(*[xs[i].shape[1] for i in grp])
        check(lib().rfn_linear_f32(len(grp), ptr_array([xs[i] for i in grp]), ld, ptr_array([Ws[i] for i in grp]), ks,
                                   ptr_array([bs[i] for i in grp]), ptr(y), y.stride(0), rows, out_features,
                                   (1 if done > 0 else 0) | 2, stream()), "rfn_linear_f32")   # 2 = RFN_GEMM_SPLITK
        done += len(grp)
    return y


class LinearFn(Function):
    """y = sum_i x_i W_i^T + b_i  (nn.Linear sums such as H2h(H) + z2h(z), misc/RecurrentFusionModel.py:53)."""

    @staticmethod
    def forward(ctx, n, *args):
        xs = [_c(a) for a in args[:n]]
        Ws = list(args[n:2 * n])
        bs = list(args[2 * n:3 * n])
        rows, out_features = xs[0].shape[0], Ws[0].shape[0]
        y = _linear_fwd(xs, Ws, bs, rows, out_features)
        ctx.n = n
        ctx.has_bias = [b is not None for b in bs]
        ctx.save_for_backward(*xs, *Ws)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        n = ctx.n
        saved = ctx.saved_tensors
        xs, Ws = saved[:n], saved[n:]
        dy = _c(dy)
        M, N = dy.shape
        gx, gw, gb = [None] * n, [None] * n, [None] * n
        for i in range(n):
            K = xs[i].shape[1]
            if ctx.needs_input_grad[1 + i]:           # dX = dY . W
                gx[i] = torch.empty(M, K, dtype=torch.float32, device=dy.device)
                _gemm_general(1, 0, dy, N, Ws[i], K, gx[i], K, M, K, N)
            if ctx.needs_input_grad[1 + n + i]:       # dW = dY^T . X
                gw[i] = torch.empty(N, K, dtype=torch.float32, device=dy.device)
                _gemm_general(0, 0, dy, N, xs[i], xs[i].stride(0), gw[i], K, N, K, M)
            if ctx.has_bias[i] and ctx.needs_input_grad[1 + 2 * n + i]:
                gb[i] = torch.empty(N, dtype=torch.float32, device=dy.device)
                check(lib().rfn_colsum_f32(ptr(dy), N, M, N, ptr(gb[i]), 0, stream()), "rfn_colsum_f32")
        return (None, *gx, *gw, *gb)


def linear(srcs):
    """srcs: list of (x, nn.Linear)."""
    n = len(srcs)
    return LinearFn.apply(n, *[x for x, _ in srcs], *[m.weight for _, m in srcs], *[m.bias for _, m in srcs])


_TRANSPOSED = {}   # data_ptr -> (version, shape, A^T): features transposed once per training step (cleared by _stages)


def clear_transposed_cache():
    _TRANSPOSED.clear()


def _transpose(x2d):
    rows, cols = x2d.shape
    ld = (rows + 3) // 4 * 4
    out = torch.empty(cols, ld, dtype=torch.float32, device=x2d.device)
    check(lib().rfn_transpose_f32(ptr(x2d), x2d.stride(0), rows, cols, ptr(out), ld, stream()), "rfn_transpose_f32")
    return out


def _transposed_features(A2):
    key = A2.data_ptr()
    hit = _TRANSPOSED.get(key)
    if hit is None or hit[0] != A2._version or hit[1] != tuple(A2.shape):
        hit = (A2._version, tuple(A2.shape), _transpose(A2))
        _TRANSPOSED[key] = hit
    return hit[2]


def _dw_long(dY, X2, out_features, in_features):
    """dW = dY^T . X for a long contraction (rows x attention locations) on the tensor engine: both operands are
    brought into its K-major layout (X^T cached across the steps that share the features) and the contraction is
    split over clusters (RFN_GEMM_SPLITK)."""
    dYt = _transpose(dY)
    Xt = _transposed_features(X2)
    K = dYt.shape[1]
    dW = torch.empty(out_features, in_features, dtype=torch.float32, device=dY.device)
    ld = (C.c_int * 1)(K)
    ks = (C.c_int * 1)(K)
    check(lib().rfn_linear_f32(1, ptr_array([dYt]), ld, ptr_array([Xt]), ks, ptr_array([None]), ptr(dW), in_features,
                               out_features, in_features, 2, stream()), "rfn_linear_f32")
    return dW


class AttentionFn(Function):
    """AttentionModelCore.forward (misc/AttentionModelCore.py:31-48) with its full backward."""

    @staticmethod
    def forward(ctx, h, A, U_w, U_b, Wh_w, Wh_b, v_w, v_b):
        h, A = _c(h), _c(A)
        rows, N, D = A.shape
        R, Ah = h.shape[1], U_w.shape[0]
        dev = h.device
        g = _linear_fwd([h], [Wh_w], [Wh_b], rows, Ah)
        P = _linear_fwd([A.view(rows * N, D)], [U_w], [U_b], rows * N, Ah)
        z = torch.empty(rows, D, dtype=torch.float32, device=dev)
        alpha = torch.empty(rows, N, dtype=torch.float32, device=dev)
        check(lib().rfn_attention_step_f32(ptr(A), ptr(P), ptr(g), ptr(v_w), ptr(v_b), ptr(z), D, ptr(alpha), rows, N, D, Ah,
                                           1, stream()), "rfn_attention_step_f32")
        ctx.save_for_backward(h, A, P, g, alpha, U_w, Wh_w, v_w)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        h, A, P, g, alpha, U_w, Wh_w, v_w = ctx.saved_tensors
        dz = _c(dz)
        rows, N, D = A.shape
        R, Ah = h.shape[1], U_w.shape[0]
        dev = h.device
        dP = torch.empty(rows * N, Ah, dtype=torch.float32, device=dev)
        dg = torch.empty(rows, Ah, dtype=torch.float32, device=dev)
        dwv = torch.zeros(Ah + 1, dtype=torch.float32, device=dev)     # one fill for both accumulators
        dw, dwb = dwv[:Ah].view(1, Ah), dwv[Ah:]
        need_dA = ctx.needs_input_grad[1]
        dA = torch.zeros(rows, N, D, dtype=torch.float32, device=dev) if need_dA else None
        check(lib().rfn_attention_step_bwd_f32(ptr(A), ptr(P), ptr(g), ptr(v_w), ptr(alpha), ptr(dz), D, ptr(dP), ptr(dg),
                                               ptr(dw), ptr(dwb), ptr(dA), rows, N, D, Ah, 1, stream()),
              "rfn_attention_step_bwd_f32")
        A2 = A.view(rows * N, D)
        if rows * N >= 1024 and Ah >= 256 and D >= 256 and D % 4 == 0 and not need_dA and lib().rfn_get_gemm_mode() >= 1:
            dU_w = _dw_long(dP, A2, Ah, D)                                     # dU = dP^T . A on the tensor engine
        else:
            dU_w = torch.empty(Ah, D, dtype=torch.float32, device=dev)
            _gemm_general(0, 0, dP, Ah, A2, D, dU_w, D, Ah, D, rows * N)      # dU = dP^T . A
        dU_b = torch.empty(Ah, dtype=torch.float32, device=dev)
        check(lib().rfn_colsum_f32(ptr(dP), Ah, rows * N, Ah, ptr(dU_b), 0, stream()), "rfn_colsum_f32")
        if need_dA:                                                            # dA += dP . U
            _gemm_general(1, 0, dP, Ah, U_w, D, dA.view(rows * N, D), D, rows * N, D, Ah, accumulate=1)
        dWh_w = torch.empty(Ah, R, dtype=torch.float32, device=dev)
        _gemm_general(0, 0, dg, Ah, h, R, dWh_w, R, Ah, R, rows)
        dWh_b = torch.empty(Ah, dtype=torch.float32, device=dev)
        check(lib().rfn_colsum_f32(ptr(dg), Ah, rows, Ah, ptr(dWh_b), 0, stream()), "rfn_colsum_f32")
        dh = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty(rows, R, dtype=torch.float32, device=dev)
            _gemm_general(1, 0, dg, Ah, Wh_w, R, dh, R, rows, R, Ah)
        return dh, dA, dU_w, dU_b, dWh_w, dWh_b, dw, dwb


def attention(att_mod, pre_h, att_seq):
    return AttentionFn.apply(pre_h, att_seq, att_mod.att_2_att_h.weight, att_mod.att_2_att_h.bias,
                             att_mod.h_2_att_h.weight, att_mod.h_2_att_h.bias, att_mod.att_h_2_out.weight,
                             att_mod.att_h_2_out.bias)


class CellFn(Function):
    """LSTM update, gate order [i|f|o|g] (misc/RecurrentFusionModel.py:55-73)."""

    @staticmethod
    def forward(ctx, G, c_prev, maxout=0):
        G, c_prev = _c(G), _c(c_prev)
        rows, R = c_prev.shape
        h = torch.empty_like(c_prev)
        c = torch.empty_like(c_prev)
        check(lib().rfn_lstm_cell_ex_f32(ptr(G), ptr(c_prev), None, 1.0, int(maxout), ptr(h), ptr(c), None, 0, None, 0, rows, R,
                                         stream()), "rfn_lstm_cell_ex_f32")
        ctx.save_for_backward(G, c_prev)
        ctx.maxout = int(maxout)
        return h, c

    @staticmethod
    @once_differentiable
    def backward(ctx, dh, dc):
        G, c_prev = ctx.saved_tensors
        rows, R = c_prev.shape
        dG = torch.empty_like(G)
        dcp = torch.empty_like(c_prev)
        srcs = [_c(dh)] if dh is not None else []
        ld = (C.c_int * 1)(R)
        check(lib().rfn_lstm_cell_bwd_ex_f32(ptr(G), ptr(c_prev), len(srcs), ptr_array(srcs) if srcs else None, ld, None, 1.0,
                                             ctx.maxout, ptr(_c(dc)) if dc is not None else None, ptr(dG), ptr(dcp), rows, R,
                                             stream()), "rfn_lstm_cell_bwd_ex_f32")
        return dG, dcp, None


class LogSoftmaxFn(Function):
    @staticmethod
    def forward(ctx, logits):
        logits = _c(logits)
        rows, V = logits.shape
        lp = torch.empty_like(logits)
        check(lib().rfn_log_softmax_f32(ptr(logits), V, ptr(lp), V, rows, V, stream()), "rfn_log_softmax_f32")
        ctx.save_for_backward(lp)
        return lp

    @staticmethod
    @once_differentiable
    def backward(ctx, dlp):
        (lp,) = ctx.saved_tensors
        dlp = _c(dlp)
        rows, V = lp.shape
        dx = torch.empty_like(lp)
        check(lib().rfn_log_softmax_bwd_f32(ptr(lp), V, ptr(dlp), V, ptr(dx), V, rows, V, stream()), "rfn_log_softmax_bwd_f32")
        return dx


class EmbedFn(Function):
    @staticmethod
    def forward(ctx, tok, weight):
        tok = tok.to(torch.int64).contiguous()
        rows = tok.shape[0]
        V1, E = weight.shape
        x = torch.empty(rows, E, dtype=torch.float32, device=weight.device)
        check(lib().rfn_embed_f32(ptr(tok), 1, ptr(weight), ptr(x), rows, E, V1, stream()), "rfn_embed_f32")
        ctx.save_for_backward(tok)
        ctx.shape = (V1, E)
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, dx):
        (tok,) = ctx.saved_tensors
        V1, E = ctx.shape
        dE = torch.zeros(V1, E, dtype=torch.float32, device=dx.device)
        check(lib().rfn_embed_bwd_f32(ptr(tok), 1, ptr(_c(dx)), ptr(dE), tok.shape[0], E, V1, stream()), "rfn_embed_bwd_f32")
        return None, dE


class MaxOverStepsFn(Function):
    """torch.max(reason_mat, 1)[0] over the review steps (misc/RecurrentFusionModel.py:229, :253)."""

    @staticmethod
    def forward(ctx, x):  # (rows, S, K)
        x = _c(x)
        rows, S, K = x.shape
        out = torch.empty(rows, K, dtype=torch.float32, device=x.device)
        check(lib().rfn_max_over_steps_f32(ptr(x), ptr(out), rows, S, K, stream()), "rfn_max_over_steps_f32")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        rows, S, K = x.shape
        din = torch.empty_like(x)
        check(lib().rfn_max_over_steps_bwd_f32(ptr(x), ptr(_c(dout)), ptr(din), rows, S, K, stream()),
              "rfn_max_over_steps_bwd_f32")
        return din


def _axpby(alpha, x, beta, y):
    out = torch.empty_like(x)
    check(lib().rfn_axpby_f32(float(alpha), ptr(x), float(beta), ptr(y) if y is not None else None, ptr(out), x.numel(),
                              stream()), "rfn_axpby_f32")
    return out


class MeanFn(Function):
    """(x_0 + x_1 + ...) / J  -- the stage-1 -> stage-2 bridge (misc/RecurrentFusionModel.py:233-235)."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [_c(x) for x in xs]
        acc = xs[0]
        for x in xs[1:]:
            acc = _axpby(1.0, acc, 1.0, x)
        ctx.n = len(xs)
        return _axpby(1.0 / len(xs), acc, 0.0, None)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        g = _axpby(1.0 / ctx.n, _c(dout), 0.0, None)
        return tuple(g for _ in range(ctx.n))


class DropoutFn(Function):
    """nn.Dropout with an explicit keep-mask (torch draws the mask; the arithmetic is ours)."""

    @staticmethod
    def forward(ctx, x, mask, scale):
        x, mask = _c(x), _c(mask)
        out = torch.empty_like(x)
        check(lib().rfn_mul_scale_f32(float(scale), ptr(x), ptr(mask), ptr(out), x.numel(), stream()), "rfn_mul_scale_f32")
        ctx.save_for_backward(mask)
        ctx.scale = scale
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        dout = _c(dout)
        dx = torch.empty_like(dout)
        check(lib().rfn_mul_scale_f32(float(ctx.scale), ptr(dout), ptr(mask), ptr(dx), dout.numel(), stream()),
              "rfn_mul_scale_f32")
        return dx, None, None


#: test hook: a list of keep-masks consumed front to back instead of drawing them (tests/test_gpu_tape.py feeds the per-op
#: tape and the fused tape the same masks)
MASK_QUEUE = None


def dropout(x, p, training, mask=None):
    if not training or p <= 0:
        return x
    if mask is None and MASK_QUEUE:
        mask = MASK_QUEUE.pop(0)
    if mask is None:
        mask = (torch.rand_like(x) >= p).float()
    return DropoutFn.apply(x, mask, 1.0 / (1.0 - p))


class ExpandRowsFn(Function):
    """Each unique image row -> g identical rows (the seq_per_img replicas, dataloader.py:251-252);
    backward sums the replicas' gradients, so stages 1-2 run once per image."""

    @staticmethod
    def forward(ctx, x, g):
        x = _c(x)
        n = x.shape[0]
        R = x[0].numel()
        out = torch.empty((n * g,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        check(lib().rfn_expand_rows_f32(ptr(x), g, ptr(out), n * g, R, stream()), "rfn_expand_rows_f32")
        ctx.g = g
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        dout = _c(dout)
        g = ctx.g
        n = dout.shape[0] // g
        R = dout[0].numel()
        dx = torch.empty((n,) + tuple(dout.shape[1:]), dtype=torch.float32, device=dout.device)
        check(lib().rfn_group_sum_f32(ptr(dout), g, ptr(dx), n, R, stream()), "rfn_group_sum_f32")
        return dx, None


class GatherColsFn(Function):
    """logprobs.gather(1, it) (misc/RecurrentFusionModel.py:632)."""

    @staticmethod
    def forward(ctx, lp, idx):
        lp = _c(lp)
        idx = idx.to(torch.int64).contiguous()
        rows, V = lp.shape
        out = torch.empty(rows, dtype=torch.float32, device=lp.device)
        check(lib().rfn_gather_cols_f32(ptr(lp), V, ptr(idx), ptr(out), rows, stream()), "rfn_gather_cols_f32")
        ctx.save_for_backward(idx)
        ctx.V = V
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        rows = idx.shape[0]
        dx = torch.empty(rows, ctx.V, dtype=torch.float32, device=dout.device)
        check(lib().rfn_scatter_cols_f32(ptr(_c(dout)), ptr(idx), ptr(dx), ctx.V, rows, ctx.V, stream()), "rfn_scatter_cols_f32")
        return dx, None


def select_token(lp, uniforms=None, temperature=1.0):
    """-> (token (rows,) int64, lp[token] (rows,)) without gradient."""
    lp = _c(lp.detach())
    rows, V = lp.shape
    tok = torch.empty(rows, dtype=torch.int64, device=lp.device)
    val = torch.empty(rows, dtype=torch.float32, device=lp.device)
    check(lib().rfn_select_token_f32(ptr(lp), V, rows, V, ptr(_c(uniforms)) if uniforms is not None else None,
                                     float(temperature), ptr(tok), ptr(val), stream()), "rfn_select_token_f32")
    return tok, val


# ---- criteria -------------------------------------------------------------------------------------------
class XeLossFn(Function):
    """Sequence term of ReviewNetEnsembleCriterion (misc/utils.py:161-184)."""

    @staticmethod
    def forward(ctx, lp, target, mask, eps):
        lp = _dense3(lp)
        rows, T, V = lp.shape
        target = target[:, :T].to(torch.int64).contiguous()
        mask = _c(mask[:, :T])
        out = torch.zeros(1, dtype=torch.float32, device=lp.device)
        check(lib().rfn_xe_loss_strided_f32(ptr(lp), lp.stride(0), lp.stride(1), ptr(target), ptr(mask), target.stride(0), rows, T,
                                            V, float(eps), ptr(out), stream()), "rfn_xe_loss_strided_f32")
        ctx.save_for_backward(target, mask)
        ctx.dims = (rows, T, V, float(eps))
        ctx.strides = (lp.stride(0), lp.stride(1))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        target, mask = ctx.saved_tensors
        rows, T, V, eps = ctx.dims
        ld_b, ld_s = ctx.strides      # the gradient takes the layout of the log-prob table (time-major from tape.DecoderFn)
        dlp = torch.empty_strided((rows, T, V), (ld_b, ld_s, 1), dtype=torch.float32, device=gout.device)
        check(lib().rfn_xe_loss_bwd_strided_f32(ptr(target), ptr(mask), target.stride(0), rows, T, V, eps, ptr(_c(gout)), ptr(dlp),
                                                ld_b, ld_s, stream()), "rfn_xe_loss_bwd_strided_f32")
        return dlp, None, None, None


class RlLossFn(Function):
    """Sequence + entropy terms of ReviewNetRewardCriterion, non-PPO (misc/utils.py:50-72)."""

    @staticmethod
    def forward(ctx, slp, seq, reward, lp_all, entropy_reg):
        slp, reward, lp_all = _c(slp), _c(reward), _dense3(lp_all)
        seq = seq.to(torch.int64).contiguous()
        rows, T = slp.shape
        T1, V = lp_all.shape[1], lp_all.shape[2]
        out = torch.zeros(1, dtype=torch.float32, device=slp.device)
        check(lib().rfn_rl_loss_strided_f32(ptr(slp), ptr(seq), ptr(reward), ptr(lp_all), lp_all.stride(0), lp_all.stride(1), rows,
                                            T, V, float(entropy_reg), ptr(out), stream()), "rfn_rl_loss_strided_f32")
        ctx.save_for_backward(seq, reward, lp_all)
        ctx.dims = (rows, T, T1, V, float(entropy_reg))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        seq, reward, lp_all = ctx.saved_tensors
        rows, T, T1, V, ent = ctx.dims
        dslp = torch.empty(rows, T, dtype=torch.float32, device=gout.device)
        dlp = torch.empty_strided((rows, T1, V), (lp_all.stride(0), lp_all.stride(1), 1), dtype=torch.float32, device=gout.device)
        check(lib().rfn_rl_loss_bwd_strided_f32(ptr(seq), ptr(reward), ptr(lp_all), lp_all.stride(0), lp_all.stride(1), rows, T, T1,
                                                V, ent, ptr(_c(gout)), ptr(dslp), ptr(dlp), stream()), "rfn_rl_loss_bwd_strided_f32")
        return dslp, None, None, dlp, None


class MarginFn(Function):
    """weight * nn.MultiLabelMarginLoss()(pred, target) (misc/utils.py:76-82, :186-190)."""

    @staticmethod
    def forward(ctx, pred, target, weight):
        ctx.in_shape = pred.shape
        pred = _c(pred)
        if pred.dim() == 1:
            pred = pred.unsqueeze(0)
        target = target.to(torch.int64).contiguous()
        rows, K = pred.shape
        out = torch.zeros(1, dtype=torch.float32, device=pred.device)
        check(lib().rfn_multilabel_margin_f32(ptr(pred), ptr(target), rows, K, float(weight), 0, ptr(out), stream()),
              "rfn_multilabel_margin_f32")
        ctx.save_for_backward(pred, target)
        ctx.weight = float(weight)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        pred, target = ctx.saved_tensors
        rows, K = pred.shape
        dx = torch.empty_like(pred)
        check(lib().rfn_multilabel_margin_bwd_f32(ptr(pred), ptr(target), rows, K, ctx.weight, ptr(_c(gout)), ptr(dx),
                                                  stream()), "rfn_multilabel_margin_bwd_f32")
        return dx.view(ctx.in_shape), None, None


class AddScalarsFn(Function):
    """Sum of 1-element loss terms (device side, no host sync)."""

    @staticmethod
    def forward(ctx, *terms):
        acc = _c(terms[0])
        for t in terms[1:]:
            acc = _axpby(1.0, acc, 1.0, _c(t))
        ctx.n = len(terms)
        return acc

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        return tuple(gout for _ in range(ctx.n))
