"""Hand-scheduled training tape: stage 1, stage 2 and the teacher-forced decoder as ONE autograd.Function each.

autograd.py mirrors the reference op by op (train.py:154-163 runs autograd over nn.Linear / tanh / softmax / bmm): ~3,900
dependent kernels per XE step, a third of them autograd's own glue (gradient accumulation adds, zero fills, memsets in
front of every split-K GEMM), all on one stream.  Here the same arithmetic is issued by explicit forward and backward
loops over the SAME kernels of librfn_b200.so:

  * the J encoder cells of a fusion / review step run on J side streams (Jacobi update: they are independent,
    misc/RecurrentFusionModel.py:102-112), in the backward pass too;
  * weight gradients (dW = dY^T X, bias column sums) leave the dependent chain: they are issued on separate
    weight-gradient streams as soon as their dY exists;
  * the gradient pieces of a hidden state (thought-vector slot, slice of dH, query gradient) are summed inside the cell
    backward kernel (rfn_lstm_cell_bwd_multi_f32) instead of by add kernels; dH of a fusion step is one multi-source GEMM;
  * the decoder's weights are shared by all T steps: their gradients are T-batched GEMMs (contraction T x rows) after the
    time loop, the loop-invariant att_2_att_h(TV_comb) and the x_t . i2h^T term are hoisted out of the loop, the logits of
    all steps are one (T x rows, V) GEMM;
  * split-K GEMM outputs live in one pre-zeroed arena per stage (one fill instead of a memset per GEMM).

The numbers are those of the op-by-op tape up to summation order (checked against oracle autograd, all 773 tensors,
tests/test_gpu_training.py).  Used by training.py when stage-1 / stage-2 dropout is off (the shipped scripts) and
model.fused_tape is true; decoder dropout (drop_prob_lm) is supported with an explicit keep-mask.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ._capi import check, lib, ptr, ptr_array

_f32 = torch.float32


# ---- streams -------------------------------------------------------------------------------------------------------------
class _Streams:
    """Per-device pool: `enc[j]` carries encoder j's dependent chain, `wg[j]` its weight gradients."""
    _pools = {}

    def __init__(self, device, n):
        self.enc = [torch.cuda.Stream(device=device) for _ in range(n)]
        self.wg = [torch.cuda.Stream(device=device) for _ in range(n)]

    @classmethod
    def get(cls, device, n):
        key = (device.index if device.index is not None else torch.cuda.current_device(), n)
        p = cls._pools.get(key)
        if p is None:
            p = cls._pools[key] = _Streams(device, n)
        return p


#: temporaries read by kernels on side streams: the caching allocator hands a freed block to the next allocation of the
#: allocating (main) stream at once, so they stay referenced until the streams have been joined back
_KEEP = []


#: data-parallel hook (dist.OverlappedGradSync.install sets it): an object with early(pairs, streams) and join().  Stage 1 owns
#: 95 % of the parameters (distinct weights per fusion step) and its backward is the LAST thing a step computes: handing each
#: fusion step's finished weight gradients to the all-reduce right away is what lets the collective overlap the remaining steps.
GRAD_SINK = None

#: phase markers for profiles/experiments/r2_phase_timeline.py: when a list, mark(name) records a timing event on the current
#: stream (external events are legal inside a CUDA-graph capture and are re-recorded by every replay)
MARKS = None


def mark(name):
    if MARKS is not None:
        ev = torch.cuda.Event(enable_timing=True, external=True)
        ev.record()
        MARKS.append((name, ev))


def _after(dst, src):
    """dst waits for everything issued to src so far."""
    ev = torch.cuda.Event()
    ev.record(src)
    dst.wait_event(ev)


def _sid(st):
    return st.cuda_stream


class _Arena:
    """Views into one zero-filled buffer (split-K GEMMs and atomically accumulated gradients add into them)."""

    def __init__(self, device):
        self.device = device
        self.req = []

    def take(self, *shape):
        n = 1
        for s in shape:
            n *= int(s)
        n4 = (n + 63) // 64 * 64     # 256-byte granules: every view stays 16-byte aligned for the tensor engine
        self.req.append((shape, n, n4))
        return len(self.req) - 1

    def build(self):
        total = sum(r[2] for r in self.req)
        buf = torch.zeros(max(total, 1), dtype=_f32, device=self.device)
        out, off = [], 0
        for shape, n, n4 in self.req:
            out.append(buf[off:off + n].view(*shape))
            off += n4
        return out


# ---- raw operator calls (explicit stream, no autograd) ---------------------------------------------------------------------
def _lin(xs, Ws, bs, y, M, N, acc, st):
    """y[M,N] (+)= sum_i x_i W_i^T + b_i; x_i / y may be row-strided views."""
    n = len(xs)
    done = 0
    while done < n:
        grp = list(range(done, min(n, done + 3)))
        ld = (C.c_int * len(grp))(*[xs[i].stride(0) for i in grp])
        ks = (C.c_int * len(grp))(*[xs[i].shape[1] for i in grp])
        flags = (1 if (acc or done > 0) else 0) | 2
        check(lib().rfn_linear_f32(len(grp), ptr_array([xs[i] for i in grp]), ld, ptr_array([Ws[i] for i in grp]), ks,
                                   ptr_array([bs[i] for i in grp]), ptr(y), y.stride(0), M, N, flags, st), "rfn_linear_f32")
        done += len(grp)


def _bwd_x(dYs, Ws, dX, M, K, acc, st):
    """dX[M,K] (+)= sum_i dY_i . W_i   (W_i: nn.Linear weights (N_i, K))."""
    n = len(dYs)
    done = 0
    while done < n:
        grp = list(range(done, min(n, done + 3)))
        ldy = (C.c_int * len(grp))(*[dYs[i].stride(0) for i in grp])
        ldw = (C.c_int * len(grp))(*[Ws[i].stride(0) for i in grp])
        nc = (C.c_int * len(grp))(*[Ws[i].shape[0] for i in grp])
        check(lib().rfn_linear_bwd_x_f32(len(grp), ptr_array([dYs[i] for i in grp]), ldy, ptr_array([Ws[i] for i in grp]), ldw, nc,
                                         ptr(dX), dX.stride(0), M, K, 1 if (acc or done > 0) else 0, st), "rfn_linear_bwd_x_f32")
        done += len(grp)


def _dw(dY, X, st, long_ok=True, xt=None):
    """dW[N,K] = dY[M,N]^T . X[M,K]; xt: X^T (K, ld) if the caller already holds it (shared by several gradients)."""
    M, N = dY.shape
    K = X.shape[1]
    dW = torch.empty(N, K, dtype=_f32, device=dY.device)
    # tensor engine through K-major copies: long contractions, and short ones (an 80-row batch) when the output is large --
    # the FFMA kernel needs 40 us for a 2048 x 4608 x 80 gate-weight gradient, transposes + tensor GEMM a third of that
    big = (M >= 1024 and N >= 256 and K >= 256) or (M >= 32 and N >= 256 and K >= 256 and N * K >= (1 << 21))
    if long_ok and big and K % 4 == 0 and lib().rfn_get_gemm_mode() >= 1:
        dYt = _transpose(dY, st)
        Xt = xt if xt is not None else _transpose(X, st)
        Kc = dYt.shape[1]
        ld = (C.c_int * 1)(Kc)
        ks = (C.c_int * 1)(Kc)
        check(lib().rfn_linear_f32(1, ptr_array([dYt]), ld, ptr_array([Xt]), ks, ptr_array([None]), ptr(dW), K, N, K, 2, st),
              "rfn_linear_f32")
        return dW
    check(lib().rfn_gemm_general_f32(0, 0, ptr(dY), dY.stride(0), ptr(X), X.stride(0), ptr(dW), K, N, K, M, 0, st),
          "rfn_gemm_general_f32")
    return dW


def _transpose(x2d, st):
    rows, cols = x2d.shape
    ld = (rows + 3) // 4 * 4
    out = torch.empty(cols, ld, dtype=_f32, device=x2d.device)
    check(lib().rfn_transpose_f32(ptr(x2d), x2d.stride(0), rows, cols, ptr(out), ld, st), "rfn_transpose_f32")
    _KEEP.append(out)
    return out


def _dw_pre(dY, Xt, K, st):
    """dW = dY^T . X with X^T (K, ld) already transposed (the image features, shared by the S0 fusion steps)."""
    M, N = dY.shape
    dYt = _transpose(dY, st)
    Kc = dYt.shape[1]
    dW = torch.empty(N, K, dtype=_f32, device=dY.device)
    ld = (C.c_int * 1)(Kc)
    ks = (C.c_int * 1)(Kc)
    check(lib().rfn_linear_f32(1, ptr_array([dYt]), ld, ptr_array([Xt]), ks, ptr_array([None]), ptr(dW), K, N, K, 2, st),
          "rfn_linear_f32")
    return dW


def _split(x2d, st, bf16):
    """fp32 rows -> the split engine's operand (power-of-two row scale + two fp16 pieces, or one bf16 piece; rfn_h3.cuh)."""
    rows, K = x2d.shape
    ks = (C.c_int * 1)(K)
    nbytes = lib().rfn_split_bytes(rows, 1, ks, bf16)
    out = torch.empty(int(nbytes), dtype=torch.uint8, device=x2d.device)
    ld = (C.c_int * 1)(x2d.stride(0))
    check(lib().rfn_split_rows_f32(1, ptr_array([x2d]), ld, ks, rows, bf16, ptr(out), out.numel(), st), "rfn_split_rows_f32")
    _KEEP.append(out)
    return out


def _lin_split(xs, W, b, y, M, N, bf16, st):
    """y = x W^T + b with x pre-split (`xs` from _split) and W split here: the stage-1 att_2_att_h projection, 89 % of the
    step's FLOPs, on the split-fp16 tcgen05 engine (engine modes 4 / 5) instead of 3xTF32."""
    K = W.shape[1]
    ws = _split(W, st, bf16)
    ks = (C.c_int * 1)(K)
    check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(ws), ks, ptr_array([b]), ptr(y), y.stride(0), M, N, 0, st), "rfn_linear_split")


def _colsum(dY, st):
    M, N = dY.shape
    db = torch.empty(N, dtype=_f32, device=dY.device)
    check(lib().rfn_colsum_f32(ptr(dY), dY.stride(0), M, N, ptr(db), 0, st), "rfn_colsum_f32")
    return db


def _att_fwd(A, P, g, v_w, v_b, z, alpha, rows, N, D, Ah, st):
    check(lib().rfn_attention_step_f32(ptr(A), ptr(P), ptr(g), ptr(v_w), ptr(v_b), ptr(z), z.stride(0), ptr(alpha), rows, N, D, Ah,
                                       1, st), "rfn_attention_step_f32")


def _att_bwd(A, P, g, v_w, alpha, dz, dP, dg, dw, dwb, dA, rows, N, D, Ah, st):
    check(lib().rfn_attention_step_bwd_f32(ptr(A), ptr(P), ptr(g), ptr(v_w), ptr(alpha), ptr(dz), dz.stride(0), ptr(dP), ptr(dg),
                                           ptr(dw), ptr(dwb), ptr(dA), rows, N, D, Ah, 1, st), "rfn_attention_step_bwd_f32")


def _cell(G, c_prev, mask, scale, h, c, h2, h3, rows, R, st, maxout=0):
    check(lib().rfn_lstm_cell_ex_f32(ptr(G), ptr(c_prev), ptr(mask), float(scale), int(maxout), ptr(h), ptr(c), ptr(h2),
                                     h2.stride(0) if h2 is not None else 0, ptr(h3), h3.stride(0) if h3 is not None else 0,
                                     rows, R, st), "rfn_lstm_cell_ex_f32")


def _sum(srcs, alpha, out, rows, R, st):
    """out = alpha * sum(srcs) over any number of row-strided (rows, R) views."""
    srcs = list(srcs)
    first = True
    while srcs:
        grp, srcs = srcs[:7 if not first else 8], srcs[7 if not first else 8:]
        if not first:
            grp = [out] + grp
        ld = (C.c_int * len(grp))(*[t.stride(0) for t in grp])
        a = alpha if not srcs else 1.0
        check(lib().rfn_sum_strided_f32(len(grp), ptr_array(grp), ld, float(a), ptr(out), out.stride(0), rows, R, st),
              "rfn_sum_strided_f32")
        first = False
    return out


def _cell_bwd(G, c_prev, srcs, mask, scale, dc_next, dG, dc_prev, rows, R, st, scratch=None, maxout=0):
    srcs = [s for s in srcs if s is not None]
    if len(srcs) > 8:      # more encoders than the kernel has source slots: pre-sum the tail
        _sum(srcs[7:], 1.0, scratch, rows, R, st)
        srcs = srcs[:7] + [scratch]
    ld = (C.c_int * max(1, len(srcs)))(*[t.stride(0) for t in srcs])
    check(lib().rfn_lstm_cell_bwd_ex_f32(ptr(G), ptr(c_prev), len(srcs), ptr_array(srcs) if srcs else None, ld, ptr(mask),
                                         float(scale), int(maxout), ptr(dc_next), ptr(dG), ptr(dc_prev), rows, R, st),
          "rfn_lstm_cell_bwd_ex_f32")


def _cont(t):
    return t if (t.dtype == _f32 and t.is_contiguous()) else t.float().contiguous()


# ---- stage 1: S0 fusion steps over J encoders (misc/RecurrentFusionModel.py:18-114, 283-305) ---------------------------------
class Stage1Fn(Function):
    """(h0_j = c0_j, att_j, parameters of the S0 x J cells) -> TV_j (rows, S0, R) for every encoder, mean_j h_j^S0, mean_j c_j^S0.
    Parameter order per (s, j): att_2_att_h.w .b, h_2_att_h.w .b, att_h_2_out.w .b, H2h.w .b, z2h.w .b."""

    @staticmethod
    def forward(ctx, J, S0, *t):
        h0 = [_cont(x) for x in t[:J]]
        att = [_cont(x) for x in t[J:2 * J]]
        prm = t[2 * J:]
        P = [[prm[(s * J + j) * 10:(s * J + j) * 10 + 10] for j in range(J)] for s in range(S0)]
        dev = h0[0].device
        rows, R = h0[0].shape
        A = P[0][0][0].shape[0]
        N = [a.shape[1] for a in att]
        D = [a.shape[2] for a in att]
        main = torch.cuda.current_stream()
        pool = _Streams.get(dev, J)
        E = torch.empty
        Hs = E(S0 + 1, rows, J * R, dtype=_f32, device=dev)
        Cs = E(J, S0, rows, R, dtype=_f32, device=dev)
        TV = [E(rows, S0, R, dtype=_f32, device=dev) for _ in range(J)]
        Pb = [E(S0, rows * N[j], A, dtype=_f32, device=dev) for j in range(J)]
        al = [E(S0, rows, N[j], dtype=_f32, device=dev) for j in range(J)]
        z = [E(S0, rows, D[j], dtype=_f32, device=dev) for j in range(J)]
        ar = _Arena(dev)
        ig = [ar.take(S0, rows, A) for _ in range(J)]
        iG = [ar.take(S0, rows, 4 * R) for _ in range(J)]
        bufs = ar.build()
        g = [bufs[i] for i in ig]
        G = [bufs[i] for i in iG]
        torch.cat(h0, 1, out=Hs[0])
        mark("s1_fwd_begin")
        mode = lib().rfn_get_gemm_mode()
        bf16 = 1 if mode == 5 else 0
        asplit = [None] * J
        for s in range(S0):
            for j in range(J):
                e = pool.enc[j]
                _after(e, main)
                st = _sid(e)
                if s == 0 and mode >= 4 and rows * N[j] >= 256 and A >= 256 and D[j] % 8 == 0:
                    asplit[j] = _split(att[j].view(rows * N[j], D[j]), st, bf16)     # once for the S0 steps
                U_w, U_b, Wh_w, Wh_b, v_w, v_b, H2h_w, H2h_b, z2h_w, z2h_b = P[s][j]
                h_in = Hs[s][:, j * R:(j + 1) * R]
                _lin([h_in], [Wh_w], [Wh_b], g[j][s], rows, A, True, st)
                if asplit[j] is not None:
                    _lin_split(asplit[j], U_w, U_b, Pb[j][s], rows * N[j], A, bf16, st)
                else:
                    _lin([att[j].view(rows * N[j], D[j])], [U_w], [U_b], Pb[j][s], rows * N[j], A, False, st)
                _att_fwd(att[j], Pb[j][s], g[j][s], v_w, v_b, z[j][s], al[j][s], rows, N[j], D[j], A, st)
                _lin([Hs[s], z[j][s]], [H2h_w, z2h_w], [H2h_b, z2h_b], G[j][s], rows, 4 * R, True, st)
                c_prev = h0[j] if s == 0 else Cs[j][s - 1]
                _cell(G[j][s], c_prev, None, 1.0, None, Cs[j][s], Hs[s + 1][:, j * R:(j + 1) * R], TV[j][:, s, :], rows, R, st)
            for j in range(J):
                _after(main, pool.enc[j])
        hbar = E(rows, R, dtype=_f32, device=dev)
        cbar = E(rows, R, dtype=_f32, device=dev)
        sm = _sid(main)
        _KEEP.clear()
        mark("s1_fwd_end")
        check(lib().rfn_mean_tensors_f32(ptr(Hs[S0]), R, J, ptr(hbar), R, rows, R, J * R, sm), "rfn_mean_tensors_f32")
        check(lib().rfn_mean_tensors_f32(ptr(Cs[0][S0 - 1]), S0 * rows * R, J, ptr(cbar), R, rows, R, R, sm), "rfn_mean_tensors_f32")
        ctx.J, ctx.S0 = J, S0
        ctx.save_for_backward(*h0, *att, *prm, Hs, Cs, *Pb, *al, *z, *g, *G)
        return (*TV, hbar, cbar)

    @staticmethod
    @once_differentiable
    def backward(ctx, *gr):
        J, S0 = ctx.J, ctx.S0
        sv = ctx.saved_tensors
        h0, att = sv[:J], sv[J:2 * J]
        np_ = S0 * J * 10
        prm = sv[2 * J:2 * J + np_]
        rest = sv[2 * J + np_:]
        Hs, Cs = rest[0], rest[1]
        Pb, al, z, g, G = (rest[2 + k * J:2 + (k + 1) * J] for k in range(5))
        P = [[prm[(s * J + j) * 10:(s * J + j) * 10 + 10] for j in range(J)] for s in range(S0)]
        dTV = [(_cont(x) if x is not None else None) for x in gr[:J]]
        dhbar, dcbar = gr[J], gr[J + 1]
        dev = Hs.device
        rows, R = h0[0].shape
        A = P[0][0][0].shape[0]
        N = [a.shape[1] for a in att]
        D = [a.shape[2] for a in att]
        main = torch.cuda.current_stream()
        pool = _Streams.get(dev, J)
        sm = _sid(main)
        E = torch.empty
        # the bridge: h = mean_j h_j, c = mean_j c_j
        dhm = dcm = None
        if dhbar is not None:
            dhm = E(rows, R, dtype=_f32, device=dev)
            _sum([_cont(dhbar)], 1.0 / J, dhm, rows, R, sm)
        if dcbar is not None:
            dcm = E(rows, R, dtype=_f32, device=dev)
            _sum([_cont(dcbar)], 1.0 / J, dcm, rows, R, sm)
        dG = [E(S0, rows, 4 * R, dtype=_f32, device=dev) for _ in range(J)]
        dP = [E(S0, rows * N[j], A, dtype=_f32, device=dev) for j in range(J)]
        dg = [E(S0, rows, A, dtype=_f32, device=dev) for j in range(J)]
        dC = [E(2, rows, R, dtype=_f32, device=dev) for _ in range(J)]
        scratch = E(J, rows, R, dtype=_f32, device=dev)
        ar = _Arena(dev)
        idz = [ar.take(S0, rows, D[j]) for j in range(J)]
        idHp = [ar.take(S0, rows, J * R) for _ in range(J)]     # dG_j . H2h_j: encoder j's share of dH per step (split-K)
        idhq = [ar.take(S0, rows, R) for _ in range(J)]         # dg_j . h_2_att_h_j: query gradient per step
        idw = [ar.take(S0, A + 1) for _ in range(J)]            # att_h_2_out weight | bias gradients (atomics)
        idUb = [ar.take(S0, A) for _ in range(J)]
        bufs = ar.build()
        dHp = [bufs[i] for i in idHp]
        dhq = [bufs[i] for i in idhq]
        dwv = [bufs[i] for i in idw]
        dUb = [bufs[i] for i in idUb]
        dz = [bufs[i] for i in idz]
        # the image features in the tensor engine's K-major layout, once for the S0 steps (dU = dP^T . A)
        use_tc = lib().rfn_get_gemm_mode() >= 1
        At = [None] * J
        for j in range(J):
            if use_tc and rows * N[j] >= 1024 and A >= 256 and D[j] >= 256 and D[j] % 4 == 0:
                _after(pool.wg[j], main)
                At[j] = _transpose(att[j].view(rows * N[j], D[j]), _sid(pool.wg[j]))
        mark("s1_bwd_begin")
        # dU_{s,j} = dP_{s,j}^T . A_j shares A_j over the S0 steps: the steps' dP^T are stacked into one (S0*A, rows*N) operand and
        # all S0 gradients of an encoder come from ONE GEMM with S0*A rows after the loop -- enough output tiles to fill the GPU
        # without split-K, so it runs on the split-fp16 engine in modes 4 / 5 (the per-step form needs the 3xTF32 split-K kernel)
        batch_dU = [At[j] is not None and (rows * N[j]) % 4 == 0 and S0 * A >= 256 for j in range(J)]
        dPt = [E(S0 * A, At[j].shape[1], dtype=_f32, device=dev) if batch_dU[j] else None for j in range(J)]
        grads = [[None] * J for _ in range(S0)]
        big_dw = use_tc and rows >= 32 and 4 * R * J * R >= (1 << 21) and R % 4 == 0
        for s in range(S0 - 1, -1, -1):
            Ht = _transpose(Hs[s], sm) if big_dw else None      # H^T of this step, shared by the J encoders' dH2h
            for j in range(J):
                e, w = pool.enc[j], pool.wg[j]
                _after(e, main)
                st, sw = _sid(e), _sid(w)
                U_w, U_b, Wh_w, Wh_b, v_w, v_b, H2h_w, H2h_b, z2h_w, z2h_b = P[s][j]
                srcs = [dTV[j][:, s, :] if dTV[j] is not None else None]
                if s == S0 - 1:
                    srcs.append(dhm)
                else:
                    srcs.append(dhq[j][s + 1])
                    srcs += [dHp[k][s + 1][:, j * R:(j + 1) * R] for k in range(J)]
                c_prev = h0[j] if s == 0 else Cs[j][s - 1]
                dc_next = dcm if s == S0 - 1 else dC[j][(s + 1) & 1]
                _cell_bwd(G[j][s], c_prev, srcs, None, 1.0, dc_next, dG[j][s], dC[j][s & 1], rows, R, st, scratch[j])
                _after(w, e)
                h_in = Hs[s][:, j * R:(j + 1) * R]
                # weight gradients of the gate GEMM, off the dependent chain
                dH2h_w = _dw(dG[j][s], Hs[s], sw, xt=Ht)
                dz2h_w = _dw(dG[j][s], z[j][s], sw)
                dH2h_b = _colsum(dG[j][s], sw)
                dz2h_b = _colsum(dG[j][s], sw)
                # dependent chain: dz -> attention backward -> query gradient; this encoder's share of dH
                _bwd_x([dG[j][s]], [z2h_w], dz[j][s], rows, D[j], True, st)
                _att_bwd(att[j], Pb[j][s], g[j][s], v_w, al[j][s], dz[j][s], dP[j][s], dg[j][s], dwv[j][s][:A], dwv[j][s][A:],
                         None, rows, N[j], D[j], A, st)
                _bwd_x([dg[j][s]], [Wh_w], dhq[j][s], rows, R, True, st)
                _bwd_x([dG[j][s]], [H2h_w], dHp[j][s], rows, J * R, True, st)
                _after(w, e)
                if batch_dU[j]:
                    dU_w = None
                    check(lib().rfn_transpose_f32(ptr(dP[j][s]), A, rows * N[j], A, ptr(dPt[j][s * A:(s + 1) * A]), dPt[j].shape[1], sw),
                          "rfn_transpose_f32")
                elif At[j] is not None:
                    dU_w = _dw_pre(dP[j][s], At[j], D[j], sw)
                else:
                    dU_w = _dw(dP[j][s], att[j].view(rows * N[j], D[j]), sw)
                check(lib().rfn_colsum_f32(ptr(dP[j][s]), A, rows * N[j], A, ptr(dUb[j][s]), 1, sw), "rfn_colsum_f32")
                dWh_w = _dw(dg[j][s], h_in, sw)
                dWh_b = _colsum(dg[j][s], sw)
                grads[s][j] = [dU_w, dUb[j][s], dWh_w, dWh_b, dwv[j][s][:A].view(1, A), dwv[j][s][A:], dH2h_w, dH2h_b, dz2h_w,
                               dz2h_b]
            for j in range(J):
                _after(main, pool.enc[j])
            if GRAD_SINK is not None:      # this step's 10 J weight gradients are final once the wgrad streams get there
                GRAD_SINK.early([(P[s][j][k], grads[s][j][k]) for j in range(J) for k in range(10) if grads[s][j][k] is not None],
                                pool.wg + pool.enc)
        # h0_j is both the initial hidden and the initial cell state (misc/RecurrentFusionModel.py:333-343)
        dh0 = []
        for j in range(J):
            out = E(rows, R, dtype=_f32, device=dev)
            _sum([dhq[j][0], dC[j][0]] + [dHp[k][0][:, j * R:(j + 1) * R] for k in range(J)], 1.0, out, rows, R, sm)
            dh0.append(out)
        mark("s1_bwd_chain_end")
        mode = lib().rfn_get_gemm_mode()
        for j in range(J):
            if not batch_dU[j]:
                continue
            sw = _sid(pool.wg[j])
            Kc = rows * N[j]
            dU_all = E(S0 * A, D[j], dtype=_f32, device=dev)
            if mode >= 4 and D[j] >= 256:
                bf16 = 1 if mode == 5 else 0
                xs = _split(dPt[j][:, :Kc], sw, bf16)
                ws = _split(At[j][:, :Kc], sw, bf16)
                ks = (C.c_int * 1)(Kc)
                check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(ws), ks, ptr_array([None]), ptr(dU_all), D[j], S0 * A, D[j], 0, sw),
                      "rfn_linear_split")
            else:
                ld = (C.c_int * 1)(dPt[j].shape[1])
                ks = (C.c_int * 1)(dPt[j].shape[1])
                check(lib().rfn_linear_f32(1, ptr_array([dPt[j]]), ld, ptr_array([At[j]]), ks, ptr_array([None]), ptr(dU_all), D[j],
                                           S0 * A, D[j], 0, sw), "rfn_linear_f32")
            for s in range(S0):
                grads[s][j][0] = dU_all[s * A:(s + 1) * A]
        if GRAD_SINK is not None and any(batch_dU):
            GRAD_SINK.early([(P[s][j][0], grads[s][j][0]) for s in range(S0) for j in range(J) if batch_dU[j]], pool.wg)
        for j in range(J):
            _after(main, pool.wg[j])
        if GRAD_SINK is not None:
            GRAD_SINK.join()                # the early all-reduces wrote the gradients in place: order them before their consumers
        mark("s1_bwd_wgrad_end")
        _KEEP.clear()
        flat = []
        for s in range(S0):
            for j in range(J):
                flat += list(grads[s][j])
        return (None, None, *dh0, *([None] * J), *flat)


# ---- stage 2: S1 review steps over the J thought-vector sets (misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:41-73) ----
class Stage2Fn(Function):
    """(TV_j, h, c, parameters of the S1 review cells) -> TVc (rows, S1, R), h^S1, c^S1.
    Parameter order per step: h2h.w .b, z_2_h[j].w .b (j = 0..J-1), then per j: att_2_att_h.w .b, h_2_att_h.w .b, att_h_2_out.w .b."""

    @staticmethod
    def forward(ctx, J, S1, *t):
        TV = [_cont(x) for x in t[:J]]
        hbar, cbar = _cont(t[J]), _cont(t[J + 1])
        prm = t[J + 2:]
        per = 2 + 2 * J + 6 * J
        dev = hbar.device
        rows, R = hbar.shape
        S0 = TV[0].shape[1]
        A = prm[2 + 2 * J].shape[0]
        main = torch.cuda.current_stream()
        pool = _Streams.get(dev, J)
        sm = _sid(main)
        E = torch.empty
        TVc = E(rows, S1, R, dtype=_f32, device=dev)
        Cs = E(max(S1 - 1, 1), rows, R, dtype=_f32, device=dev)
        hfin = E(rows, R, dtype=_f32, device=dev)
        cfin = E(rows, R, dtype=_f32, device=dev)
        Pb = [E(S1, rows * S0, A, dtype=_f32, device=dev) for _ in range(J)]
        al = [E(S1, rows, S0, dtype=_f32, device=dev) for _ in range(J)]
        z = [E(S1, rows, R, dtype=_f32, device=dev) for _ in range(J)]
        ar = _Arena(dev)
        ig = [ar.take(S1, rows, A) for _ in range(J)]
        GW = prm[0].shape[0]                       # 4R, or 5R with review_maxout
        mo = 1 if GW == 5 * R else 0
        iG = ar.take(S1, rows, GW)
        bufs = ar.build()
        g = [bufs[i] for i in ig]
        G = bufs[iG]
        mark("s2_fwd_begin")
        for s in range(S1):
            p = prm[s * per:(s + 1) * per]
            hin = hbar if s == 0 else TVc[:, s - 1, :]
            for j in range(J):
                e = pool.enc[j]
                _after(e, main)
                st = _sid(e)
                U_w, U_b, Wh_w, Wh_b, v_w, v_b = p[2 + 2 * J + 6 * j:2 + 2 * J + 6 * j + 6]
                _lin([hin], [Wh_w], [Wh_b], g[j][s], rows, A, True, st)
                _lin([TV[j].view(rows * S0, R)], [U_w], [U_b], Pb[j][s], rows * S0, A, False, st)
                _att_fwd(TV[j], Pb[j][s], g[j][s], v_w, v_b, z[j][s], al[j][s], rows, S0, R, A, st)
            for j in range(J):
                _after(main, pool.enc[j])
            _lin([hin] + [z[j][s] for j in range(J)], [p[0]] + [p[2 + 2 * j] for j in range(J)],
                 [p[1]] + [p[3 + 2 * j] for j in range(J)], G[s], rows, GW, True, sm)
            last = s == S1 - 1
            c_prev = cbar if s == 0 else Cs[s - 1]
            _cell(G[s], c_prev, None, 1.0, hfin if last else None, cfin if last else Cs[s], TVc[:, s, :], None, rows, R, sm, mo)
        mark("s2_fwd_end")
        ctx.J, ctx.S1 = J, S1
        ctx.save_for_backward(*TV, hbar, cbar, *prm, TVc, Cs, *Pb, *al, *z, *g, G)
        return TVc, hfin, cfin

    @staticmethod
    @once_differentiable
    def backward(ctx, dTVc, dhfin, dcfin):
        J, S1 = ctx.J, ctx.S1
        sv = ctx.saved_tensors
        TV, hbar, cbar = sv[:J], sv[J], sv[J + 1]
        per = 2 + 2 * J + 6 * J
        prm = sv[J + 2:J + 2 + S1 * per]
        rest = sv[J + 2 + S1 * per:]
        TVc, Cs = rest[0], rest[1]
        Pb, al, z, g = (rest[2 + k * J:2 + (k + 1) * J] for k in range(4))
        G = rest[2 + 4 * J]
        dev = hbar.device
        rows, R = hbar.shape
        S0 = TV[0].shape[1]
        A = prm[2 + 2 * J].shape[0]
        main = torch.cuda.current_stream()
        pool = _Streams.get(dev, J)
        sm = _sid(main)
        E = torch.empty
        dTVc = _cont(dTVc) if dTVc is not None else None
        GW = prm[0].shape[0]
        mo = 1 if GW == 5 * R else 0
        dG = E(S1, rows, GW, dtype=_f32, device=dev)
        dC = E(2, rows, R, dtype=_f32, device=dev)
        dP = [E(S1, rows * S0, A, dtype=_f32, device=dev) for _ in range(J)]
        dg = [E(S1, rows, A, dtype=_f32, device=dev) for _ in range(J)]
        scratch = E(rows, R, dtype=_f32, device=dev)
        ar = _Arena(dev)
        idz = [ar.take(S1, rows, R) for _ in range(J)]
        idTV = [ar.take(rows, S0, R) for _ in range(J)]
        idhx = ar.take(S1, rows, R)
        idhq = [ar.take(S1, rows, R) for _ in range(J)]
        idw = [ar.take(S1, A + 1) for _ in range(J)]
        idUb = [ar.take(S1, A) for _ in range(J)]
        bufs = ar.build()
        dTV = [bufs[i] for i in idTV]
        dz = [bufs[i] for i in idz]
        dhx = bufs[idhx]
        dhq = [bufs[i] for i in idhq]
        dwv = [bufs[i] for i in idw]
        dUb = [bufs[i] for i in idUb]
        mark("s2_bwd_begin")
        grads = [None] * S1
        for s in range(S1 - 1, -1, -1):
            p = prm[s * per:(s + 1) * per]
            hin = hbar if s == 0 else TVc[:, s - 1, :]
            srcs = [dTVc[:, s, :] if dTVc is not None else None]
            if s == S1 - 1:
                srcs.append(_cont(dhfin) if dhfin is not None else None)
            else:
                srcs.append(dhx[s + 1])
                srcs += [dhq[j][s + 1] for j in range(J)]
            c_prev = cbar if s == 0 else Cs[s - 1]
            dc_next = (_cont(dcfin) if dcfin is not None else None) if s == S1 - 1 else dC[(s + 1) & 1]
            _cell_bwd(G[s], c_prev, srcs, None, 1.0, dc_next, dG[s], dC[s & 1], rows, R, sm, scratch, mo)
            gs = [None] * per
            for j in range(J):
                e, w = pool.enc[j], pool.wg[j]
                _after(e, main)
                st, sw = _sid(e), _sid(w)
                U_w, U_b, Wh_w, Wh_b, v_w, v_b = p[2 + 2 * J + 6 * j:2 + 2 * J + 6 * j + 6]
                zw = p[2 + 2 * j]
                _bwd_x([dG[s]], [zw], dz[j][s], rows, R, True, st)
                _att_bwd(TV[j], Pb[j][s], g[j][s], v_w, al[j][s], dz[j][s], dP[j][s], dg[j][s], dwv[j][s][:A], dwv[j][s][A:],
                         dTV[j], rows, S0, R, A, st)
                _bwd_x([dP[j][s]], [U_w], dTV[j].view(rows * S0, R), rows * S0, R, True, st)
                _bwd_x([dg[j][s]], [Wh_w], dhq[j][s], rows, R, True, st)
                _after(w, e)
                gs[2 + 2 * j] = _dw(dG[s], z[j][s], sw)
                gs[3 + 2 * j] = _colsum(dG[s], sw)
                o = 2 + 2 * J + 6 * j
                gs[o] = _dw(dP[j][s], TV[j].view(rows * S0, R), sw)
                check(lib().rfn_colsum_f32(ptr(dP[j][s]), A, rows * S0, A, ptr(dUb[j][s]), 1, sw), "rfn_colsum_f32")
                gs[o + 1] = dUb[j][s]
                gs[o + 2] = _dw(dg[j][s], hin, sw)
                gs[o + 3] = _colsum(dg[j][s], sw)
                gs[o + 4] = dwv[j][s][:A].view(1, A)
                gs[o + 5] = dwv[j][s][A:]
            _bwd_x([dG[s]], [p[0]], dhx[s], rows, R, True, sm)
            w0 = pool.wg[0]
            gs[0] = _dw(dG[s], hin, _sid(w0))
            gs[1] = _colsum(dG[s], _sid(w0))
            for j in range(J):
                _after(main, pool.enc[j])
            grads[s] = gs
        dhbar = E(rows, R, dtype=_f32, device=dev)
        _sum([dhx[0]] + [dhq[j][0] for j in range(J)], 1.0, dhbar, rows, R, sm)
        dcbar = dC[0]
        mark("s2_bwd_chain_end")
        for j in range(J):
            _after(main, pool.wg[j])
        mark("s2_bwd_wgrad_end")
        _KEEP.clear()
        flat = []
        for s in range(S1):
            flat += grads[s]
        return (None, None, *dTV, dhbar, dcbar, *flat)


# ---- decoder: T teacher-forced steps of LSTMSoftAttentionCore + logit + log_softmax -------------------------------------------
class DecoderFn(Function):
    """(X (T*rows, E) time-major embedded tokens, TVc, h0, c0, keep-mask (T, rows, R) or None, decoder + logit parameters)
    -> log-probs (T, rows, V) TIME-MAJOR (the caller transposes the view).
    Parameter order: i2h.w .b, h2h.w .b, z2h.w .b, att_2_att_h.w .b, h_2_att_h.w .b, att_h_2_out.w .b, logit.w .b
    (misc/LSTMSoftAttentionCore.py:60-102; logit + log_softmax misc/RecurrentFusionModel.py:278)."""

    @staticmethod
    def forward(ctx, T, scale, X, TVc, h0, c0, mask, *prm):
        X, TVc, h0, c0 = _cont(X), _cont(TVc), _cont(h0), _cont(c0)
        Wi, bi, Whh, bh, Wz, bz, U_w, U_b, Wh_w, Wh_b, v_w, v_b, Wl, bl = prm
        dev = X.device
        rows, R = h0.shape
        S1 = TVc.shape[1]
        A, E_, V = U_w.shape[0], X.shape[1], Wl.shape[0]
        main = torch.cuda.current_stream()
        sm = _sid(main)
        pool = _Streams.get(dev, 1)
        side = pool.wg[0]
        E = torch.empty
        Pdec = E(rows * S1, A, dtype=_f32, device=dev)
        Hx = E(T + 1, rows, R, dtype=_f32, device=dev)
        Cx = E(T, rows, R, dtype=_f32, device=dev)
        Z = E(T, rows, R, dtype=_f32, device=dev)
        al = E(T, rows, S1, dtype=_f32, device=dev)
        GW = Wi.shape[0]                           # 4R, or 5R with opt.maxout
        mo = 1 if GW == 5 * R else 0
        G = E(T, rows, GW, dtype=_f32, device=dev)
        g = torch.zeros(T, rows, A, dtype=_f32, device=dev)
        Hx[0].copy_(h0)
        mark("dec_fwd_begin")
        # hoisted: the x_t . i2h^T + b term of every step (one (T*rows)-row GEMM, side stream), att_2_att_h(TV_comb)
        _after(side, main)
        _lin([X], [Wi], [bi], G.view(T * rows, GW), T * rows, GW, False, _sid(side))
        _lin([TVc.view(rows * S1, R)], [U_w], [U_b], Pdec, rows * S1, A, False, sm)
        _after(main, side)
        for t in range(T):
            _lin([Hx[t]], [Wh_w], [Wh_b], g[t], rows, A, True, sm)
            _att_fwd(TVc, Pdec, g[t], v_w, v_b, Z[t], al[t], rows, S1, R, A, sm)
            _lin([Hx[t], Z[t]], [Whh, Wz], [bh, bz], G[t], rows, GW, True, sm)
            _cell(G[t], c0 if t == 0 else Cx[t - 1], mask[t] if mask is not None else None, scale, Hx[t + 1], Cx[t], None, None,
                  rows, R, sm, mo)
        mark("dec_fwd_loop_end")
        logits = E(T * rows, V, dtype=_f32, device=dev)
        _lin([Hx[1:].view(T * rows, R)], [Wl], [bl], logits, T * rows, V, False, sm)
        lp = E(T, rows, V, dtype=_f32, device=dev)
        check(lib().rfn_log_softmax_f32(ptr(logits), V, ptr(lp), V, T * rows, V, sm), "rfn_log_softmax_f32")
        mark("dec_fwd_end")
        ctx.T, ctx.scale, ctx.has_mask = T, scale, mask is not None
        ctx.save_for_backward(X, TVc, c0, *prm, Pdec, Hx, Cx, Z, al, G, g, lp, *([mask] if mask is not None else []))
        return lp

    @staticmethod
    @once_differentiable
    def backward(ctx, dlp):
        T, scale = ctx.T, ctx.scale
        sv = ctx.saved_tensors
        X, TVc, c0 = sv[:3]
        prm = sv[3:17]
        Pdec, Hx, Cx, Z, al, G, g, lp = sv[17:25]
        mask = sv[25] if ctx.has_mask else None
        Wi, bi, Whh, bh, Wz, bz, U_w, U_b, Wh_w, Wh_b, v_w, v_b, Wl, bl = prm
        dev = X.device
        rows, R = c0.shape
        S1 = TVc.shape[1]
        A, E_, V = U_w.shape[0], X.shape[1], Wl.shape[0]
        main = torch.cuda.current_stream()
        sm = _sid(main)
        pool = _Streams.get(dev, 1)
        w = pool.wg[0]
        sw = _sid(w)
        E = torch.empty
        dlp = _cont(dlp)
        mark("dec_bwd_begin")
        dlogits = E(T * rows, V, dtype=_f32, device=dev)
        check(lib().rfn_log_softmax_bwd_f32(ptr(lp), V, ptr(dlp), V, ptr(dlogits), V, T * rows, V, sm), "rfn_log_softmax_bwd_f32")
        ar = _Arena(dev)
        idH = ar.take(T, rows, R)
        idhz = ar.take(T, rows, 2 * R)
        idhq = ar.take(T, rows, R)
        idTVc = ar.take(rows, S1, R)
        idw = ar.take(A + 1)
        dHall, dhz, dhq, dTVc, dwv = ar.build()
        _bwd_x([dlogits], [Wl], dHall.view(T * rows, R), T * rows, R, True, sm)
        _after(w, main)
        dWl = _dw(dlogits, Hx[1:].view(T * rows, R), sw)
        dbl = _colsum(dlogits, sw)
        Wcat = torch.cat([Whh, Wz], 1)                 # (4R, 2R): dG . [h2h | z2h] in one launch per step
        GW = Wi.shape[0]
        mo = 1 if GW == 5 * R else 0
        dG = E(T, rows, GW, dtype=_f32, device=dev)
        dC = E(2, rows, R, dtype=_f32, device=dev)
        dP = E(T, rows * S1, A, dtype=_f32, device=dev)
        dg = E(T, rows, A, dtype=_f32, device=dev)
        mark("dec_bwd_loop_begin")
        for t in range(T - 1, -1, -1):
            srcs = [dHall[t]]
            if t < T - 1:
                srcs += [dhz[t + 1][:, :R], dhq[t + 1]]
            _cell_bwd(G[t], c0 if t == 0 else Cx[t - 1], srcs, mask[t] if mask is not None else None, scale,
                      None if t == T - 1 else dC[(t + 1) & 1], dG[t], dC[t & 1], rows, R, sm, None, mo)
            _bwd_x([dG[t]], [Wcat], dhz[t], rows, 2 * R, True, sm)
            _att_bwd(TVc, Pdec, g[t], v_w, al[t], dhz[t][:, R:], dP[t], dg[t], dwv[:A], dwv[A:], dTVc, rows, S1, R, A, sm)
            _bwd_x([dg[t]], [Wh_w], dhq[t], rows, R, True, sm)
        dh0 = E(rows, R, dtype=_f32, device=dev)
        _sum([dhz[0][:, :R], dhq[0]], 1.0, dh0, rows, R, sm)
        dc0 = dC[0]
        mark("dec_bwd_loop_end")
        # everything below is off the dependent chain: T-batched weight gradients, the embedding-side dX, the hoisted projection
        _after(w, main)
        dG2 = dG.view(T * rows, GW)
        dX = E(T * rows, E_, dtype=_f32, device=dev)
        _bwd_x([dG2], [Wi], dX, T * rows, E_, False, sw)
        dPs = E(rows * S1, A, dtype=_f32, device=dev)      # sum over the steps of dP_t (P is loop invariant)
        check(lib().rfn_colsum_f32(ptr(dP), rows * S1 * A, T, rows * S1 * A, ptr(dPs), 0, sw), "rfn_colsum_f32")
        _bwd_x([dPs], [U_w], dTVc.view(rows * S1, R), rows * S1, R, True, sw)
        dU_w = _dw(dPs, TVc.view(rows * S1, R), sw)
        dU_b = _colsum(dPs, sw)
        Hprev = Hx[:T].view(T * rows, R)
        dWi = _dw(dG2, X, sw)
        dWhh = _dw(dG2, Hprev, sw)
        dWz = _dw(dG2, Z.view(T * rows, R), sw)
        dbi, dbh, dbz = _colsum(dG2, sw), _colsum(dG2, sw), _colsum(dG2, sw)
        dg2 = dg.view(T * rows, A)
        dWh_w = _dw(dg2, Hprev, sw)
        dWh_b = _colsum(dg2, sw)
        _after(main, w)
        mark("dec_bwd_end")
        _KEEP.clear()
        return (None, None, dX, dTVc, dh0, dc0, None, dWi, dbi, dWhh, dbh, dWz, dbz, dU_w, dU_b, dWh_w, dWh_b,
                dwv[:A].view(1, A), dwv[A:], dWl, dbl)


# ---- model-level entry points ---------------------------------------------------------------------------------------------------
def stage1_params(model):
    out = []
    for s in range(model.num_review_steps_0):
        core = model.review_steps_individual[s]
        for j in range(model.num_feat_array):
            c = core.lstm[j]
            a = c.att_model
            out += [a.att_2_att_h.weight, a.att_2_att_h.bias, a.h_2_att_h.weight, a.h_2_att_h.bias, a.att_h_2_out.weight,
                    a.att_h_2_out.bias, c.H2h.weight, c.H2h.bias, c.z2h.weight, c.z2h.bias]
    return out


def stage2_params(model):
    out = []
    J = model.num_feat_array
    for s in range(model.num_review_steps):
        c = model.review_steps[s]
        out += [c.h2h.weight, c.h2h.bias]
        for j in range(J):
            out += [c.z_2_h[j].weight, c.z_2_h[j].bias]
        for j in range(J):
            a = c.att_model[j]
            out += [a.att_2_att_h.weight, a.att_2_att_h.bias, a.h_2_att_h.weight, a.h_2_att_h.bias, a.att_h_2_out.weight,
                    a.att_h_2_out.bias]
    return out


def decoder_params(model):
    d = model.decoder
    return [d.i2h.weight, d.i2h.bias, d.h2h.weight, d.h2h.bias, d.z2h.weight, d.z2h.bias, d.att_2_att_h.weight, d.att_2_att_h.bias,
            d.h_2_att_h.weight, d.h_2_att_h.bias, d.att_h_2_out.weight, d.att_h_2_out.bias, model.logit.weight, model.logit.bias]


def usable(model):
    """The fused tape covers the shipped training configuration: no stage-1 / stage-2 dropout (opts.py defaults)."""
    return bool(getattr(model, "fused_tape", True)) and not model._dropout_active(model.drop_prob_fusion) and \
        not model._dropout_active(model.drop_prob_reason)


def thought_vectors(model, fc, att):
    """get_init_state + stages 1-2 (misc/RecurrentFusionModel.py:333-343, 283-331) -> TVc, reason_pred list, (h, c)."""
    from . import autograd as AG
    J = model.num_feat_array
    S0, S1 = model.num_review_steps_0, model.num_review_steps
    h0 = [AG.linear([(fc[j], model.fc2h[j])]) for j in range(J)]       # h_j^0 = c_j^0 = fc2h_j(fc_j)
    out = Stage1Fn.apply(J, S0, *h0, *att, *stage1_params(model))
    TV, hbar, cbar = list(out[:J]), out[J], out[J + 1]
    rows = hbar.shape[0]
    reason_pred = []
    for j in range(J):       # reason_linear_individual over all S0 steps at once, then the max over steps (:291,:303)
        rm = AG.linear([(TV[j].view(rows * S0, -1), model.reason_linear_individual[j])]).view(rows, S0, -1)
        reason_pred.append(AG.MaxOverStepsFn.apply(rm))
    TVc, h, c = Stage2Fn.apply(J, S1, *TV, hbar, cbar, *stage2_params(model))
    rc = AG.linear([(TVc.view(rows * S1, -1), model.reason_linear)]).view(rows, S1, -1)
    reason_pred.append(AG.MaxOverStepsFn.apply(rc))
    return TVc, reason_pred, (h.unsqueeze(0), c.unsqueeze(0))


def decode_teacher_forced(model, tokens, TVc, state):
    """Log-probs of T teacher-forced decoder steps, (rows, T, V) as a transposed view of the time-major table.
    tokens (rows, T) int64: the input token of every step."""
    from . import autograd as AG
    rows, T = tokens.shape
    X = AG.EmbedFn.apply(tokens.t().reshape(-1), model.embed.weight)
    h0, c0 = state[0].squeeze(0), state[1].squeeze(0)
    mask, scale = None, 1.0
    p = float(model.drop_prob_lm)
    if model._dropout_active(p):
        if AG.MASK_QUEUE:                                      # test hook: the masks the per-op tape would consume
            mask = torch.stack([AG.MASK_QUEUE.pop(0) for _ in range(T)], 0).contiguous()
        else:
            mask = (torch.rand(T, rows, model.rnn_size, device=X.device) >= p).float()
        scale = 1.0 / (1.0 - p)
    lp = DecoderFn.apply(T, scale, X, TVc, h0, c0, mask, *decoder_params(model))
    return lp.transpose(0, 1)
