"""Drop-in mirror of the reference's RFNet module surface on top of librfn_b200.so.

Same class names, constructor arguments, sub-module / parameter names (so reference checkpoints
load with ``load_state_dict`` unchanged -- 773 tensors for the five-encoder model) and the same
call signatures and return structures as

  misc/AttentionModelCore.py                       AttentionModelCore
  misc/RecurrentFusionModel.py:18-74               LSTMFusionNoInputCore
  misc/RecurrentFusionModel.py:77-114              FeatArrayFusionNoInputCore
  misc/LSTMSoftMultiAttentionFeatArrayNoInputCore  LSTMSoftMultiAttentionFeatArrayNoInputCore
  misc/LSTMSoftAttentionCore.py                    LSTMSoftAttentionCore
  misc/LSTMSoftAttentionNoInputCore.py             LSTMSoftAttentionNoInputCore
  misc/RecurrentFusionModel.py:117-658             RecurrentFusionModel

Every forward here hands raw device pointers to the C ABI (include/rfn_b200.h); nothing is
computed with torch ops and there is no CPU path.  nn.Linear / nn.Embedding are parameter
containers only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _capi
from ._capi import check, lib, ptr, ptr_array, stream

_INIT = 0.1


def _uniform_(lin: nn.Linear, bias: bool = True):
    lin.weight.data.uniform_(-_INIT, _INIT)
    if bias:
        lin.bias.data.uniform_(-_INIT, _INIT)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise _capi.RfnError("recurrent_fusion_network_b200 runs on CUDA tensors only (no CPU fallback): "
                             "move the model and its inputs to the GPU")


def _taping(*tensors):
    """True when autograd must record: the call then goes through autograd.py's Function wrappers."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _dropout(x, p, training):
    from . import autograd as AG
    return AG.dropout(x, p, training)


class _Workspace:
    """Grow-only device scratch handed to the C ABI (the library never allocates)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


_WS = _Workspace()   # operator-level wrappers only (AttentionModelCore used on its own); every model owns its workspace


# --------------------------------------------------------------------------------------------------
# operator-level wrappers
# --------------------------------------------------------------------------------------------------
def linear(srcs, rows: int, out_features: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = sum_i x_i W_i^T + b_i through rfn_linear_f32.  srcs: list of (x, nn.Linear-like)."""
    xs = [_f32c(x) for x, _ in srcs]
    _require_cuda(xs[0])
    if out is None and _taping(*xs, *[m.weight for _, m in srcs]):
        from . import autograd as AG
        return AG.linear(list(zip(xs, [m for _, m in srcs])))
    n = len(srcs)
    y = out if out is not None else torch.empty(rows, out_features, dtype=torch.float32, device=xs[0].device)
    done = 0
    while done < n:
        grp = list(range(done, min(n, done + 3)))
        xa = ptr_array([xs[i] for i in grp])
        wa = ptr_array([srcs[i][1].weight for i in grp])
        ba = ptr_array([srcs[i][1].bias for i in grp])
        ld = (C.c_int * len(grp))(*[xs[i].shape[1] for i in grp])
        ks = (C.c_int * len(grp))(*[xs[i].shape[1] for i in grp])
        check(lib().rfn_linear_f32(len(grp), xa, ld, wa, ks, ba, ptr(y), y.stride(0), rows, out_features,
                                   1 if done > 0 else 0, stream()), "rfn_linear_f32")
        done += len(grp)
    return y


def lstm_cell(G: torch.Tensor, c_prev: torch.Tensor, maxout: int = 0):
    """maxout: G is (rows, 5R) and the input transform is the max of its last two R-blocks (no tanh)."""
    if _taping(G, c_prev):
        from . import autograd as AG
        return AG.CellFn.apply(G, c_prev, int(bool(maxout)))
    rows, R = c_prev.shape
    h = torch.empty_like(c_prev)
    c = torch.empty_like(c_prev)
    check(lib().rfn_lstm_cell_ex_f32(ptr(G), ptr(c_prev), None, 1.0, int(bool(maxout)), ptr(h), ptr(c), None, 0, None, 0, rows, R,
                                     stream()), "rfn_lstm_cell_ex_f32")
    return h, c


def _last_layer(t: torch.Tensor) -> torch.Tensor:
    """state[-1] of a (num_layers, rows, R) state tensor.  num_layers is 1 on this path: a squeeze is a view whose backward
    is a view too, where select's backward zero-fills a full tensor and copies (one fill + one copy kernel per use)."""
    return t.squeeze(0) if t.shape[0] == 1 else t[-1]


def _attention(att_mod, pre_h: torch.Tensor, att_seq: torch.Tensor) -> torch.Tensor:
    """AttentionModelCore.forward through rfn_attention_core_f32."""
    pre_h, att_seq = _f32c(pre_h), _f32c(att_seq)
    _require_cuda(pre_h)
    if _taping(pre_h, att_seq, att_mod.att_2_att_h.weight):
        from . import autograd as AG
        return AG.attention(att_mod, pre_h, att_seq)
    rows, N, D = att_seq.shape
    R = pre_h.shape[1]
    Ah = att_mod.att_2_att_h.weight.shape[0]
    z = torch.empty(rows, D, dtype=torch.float32, device=pre_h.device)
    nbytes = (rows * N * Ah + rows * Ah) * 4 + 256
    ws = _WS.get(nbytes, pre_h.device)
    check(lib().rfn_attention_core_f32(
        ptr(pre_h), ptr(att_seq), ptr(att_mod.att_2_att_h.weight), ptr(att_mod.att_2_att_h.bias),
        ptr(att_mod.h_2_att_h.weight), ptr(att_mod.h_2_att_h.bias), ptr(att_mod.att_h_2_out.weight),
        ptr(att_mod.att_h_2_out.bias), ptr(z), None, rows, N, D, R, Ah, ptr(ws), ws.numel(), stream()),
        "rfn_attention_core_f32")
    return z


# --------------------------------------------------------------------------------------------------
# per-timestep cores (operator-level mirrors)
# --------------------------------------------------------------------------------------------------
class AttentionModelCore(nn.Module):
    """misc/AttentionModelCore.py:8-48"""

    def __init__(self, rnn_size, att_feat_size, att_num, att_hid_size):
        super().__init__()
        self.rnn_size, self.att_hid_size = rnn_size, att_hid_size
        self.att_feat_size, self.att_num = att_feat_size, att_num
        self.att_2_att_h = nn.Linear(att_feat_size, att_hid_size)
        self.h_2_att_h = nn.Linear(rnn_size, att_hid_size)
        self.att_h_2_out = nn.Linear(att_hid_size, 1)
        for m in (self.att_2_att_h, self.h_2_att_h, self.att_h_2_out):
            _uniform_(m)

    def forward(self, pre_h, att_seq):
        return _attention(self, pre_h, att_seq)


class LSTMFusionNoInputCore(nn.Module):
    """Stage-1 cell of one encoder (misc/RecurrentFusionModel.py:18-74)."""

    def __init__(self, H_size, rnn_size, att_feat_size, att_num, att_hid_size, drop_prob_fusion, maxout=0):
        super().__init__()
        # maxout = 1 is unreachable through RecurrentFusionModel (FeatArrayFusionNoInputCore never forwards fusion_maxout,
        # misc/RecurrentFusionModel.py:93-97); the cell itself supports it as the reference class does (:33-38, :61-65)
        self.drop_prob_fusion, self.att_hid_size, self.maxout = drop_prob_fusion, att_hid_size, maxout
        self.H_size, self.rnn_size, self.att_feat_size, self.att_num = H_size, rnn_size, att_feat_size, att_num
        self.att_model = AttentionModelCore(rnn_size, att_feat_size, att_num, att_hid_size)
        gw = (5 if maxout else 4) * rnn_size
        self.H2h = nn.Linear(H_size, gw)
        self.z2h = nn.Linear(att_feat_size, gw)
        self.dropout = nn.Dropout(drop_prob_fusion)
        self.H2h.weight.data.uniform_(-_INIT, _INIT)   # biases keep nn.Linear's default (:43-45)
        self.z2h.weight.data.uniform_(-_INIT, _INIT)

    def forward(self, H, att_feat, state):
        pre_h, pre_c = _last_layer(state[0]), _last_layer(state[1])
        z = self.att_model(pre_h, att_feat)
        G = linear([(H, self.H2h), (z, self.z2h)], pre_h.shape[0], self.H2h.out_features)
        next_h, next_c = lstm_cell(G, _f32c(pre_c), self.maxout)
        next_h = _dropout(next_h, self.drop_prob_fusion, self.training)
        return next_h, (next_h.unsqueeze(0), next_c.unsqueeze(0))


class FeatArrayFusionNoInputCore(nn.Module):
    """One stage-1 fusion step over all encoders (misc/RecurrentFusionModel.py:77-114)."""

    def __init__(self, num_feat_array, rnn_size, att_feat_size, att_num, att_hid_size, drop_prob_fusion, maxout=0):
        super().__init__()
        self.rnn_size, self.drop_prob_fusion = rnn_size, drop_prob_fusion
        self.att_feat_size, self.att_num, self.att_hid_size = att_feat_size, att_num, att_hid_size
        self.maxout, self.num_feat_array = maxout, num_feat_array
        self.H_size = num_feat_array * rnn_size
        self.Z_size = sum(att_feat_size)
        self.lstm = nn.ModuleList([
            LSTMFusionNoInputCore(self.H_size, rnn_size, att_feat_size[i], att_num[i], att_hid_size, drop_prob_fusion)
            for i in range(num_feat_array)])     # fusion_maxout is never forwarded (:93-97)
        self.dropout = nn.Dropout(drop_prob_fusion)

    def forward(self, att_seq, state_list):
        H = torch.cat([_last_layer(state_list[i][0]) for i in range(self.num_feat_array)], 1)   # previous states (:102-107)
        output_list = []
        for i in range(self.num_feat_array):
            output, state_list[i] = self.lstm[i](H, att_seq[i], state_list[i])
            output_list.append(output)
        return output_list, state_list


class LSTMSoftMultiAttentionFeatArrayNoInputCore(nn.Module):
    """Stage-2 review cell (misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:9-73)."""

    def __init__(self, rnn_size, att_feat_size, att_num, att_hid_size, drop_prob_lm, maxout=0):
        super().__init__()
        assert len(att_feat_size) == len(att_num)
        self.rnn_size, self.att_feat_size, self.att_num = rnn_size, att_feat_size, att_num
        self.num_feat_array = len(att_feat_size)
        self.drop_prob_lm, self.att_hid_size, self.maxout = drop_prob_lm, att_hid_size, maxout
        gw = (5 if maxout else 4) * rnn_size                   # :25-30
        self.h2h = nn.Linear(rnn_size, gw)
        self.z_2_h = nn.ModuleList([nn.Linear(att_feat_size[i], gw) for i in range(self.num_feat_array)])
        self.att_model = nn.ModuleList([AttentionModelCore(rnn_size, att_feat_size[i], att_num[i], att_hid_size)
                                        for i in range(self.num_feat_array)])
        self.dropout = nn.Dropout(drop_prob_lm)
        _uniform_(self.h2h)                            # z_2_h keeps nn.Linear's default init (:35-38)

    def forward(self, att_seq, state):
        pre_h, pre_c = _last_layer(state[0]), _last_layer(state[1])
        zs = [self.att_model[i](pre_h, att_seq[i]) for i in range(self.num_feat_array)]
        srcs = [(pre_h, self.h2h)] + [(zs[i], self.z_2_h[i]) for i in range(self.num_feat_array)]
        G = linear(srcs, pre_h.shape[0], self.h2h.out_features)
        next_h, next_c = lstm_cell(G, _f32c(pre_c), self.maxout)
        next_h = _dropout(next_h, self.drop_prob_lm, self.training)
        return next_h, (next_h.unsqueeze(0), next_c.unsqueeze(0))


class LSTMSoftAttentionCore(nn.Module):
    """Decoder cell (misc/LSTMSoftAttentionCore.py:12-102)."""

    def __init__(self, input_encoding_size, rnn_size, att_feat_size, att_num, att_hid_size, drop_prob_lm, maxout=0):
        super().__init__()
        self.input_encoding_size, self.rnn_size, self.drop_prob_lm = input_encoding_size, rnn_size, drop_prob_lm
        self.att_feat_size, self.att_num, self.att_hid_size, self.maxout = att_feat_size, att_num, att_hid_size, maxout
        gw = (5 if maxout else 4) * rnn_size                   # :25-32
        self.i2h = nn.Linear(input_encoding_size, gw)
        self.h2h = nn.Linear(rnn_size, gw)
        self.z2h = nn.Linear(att_feat_size, gw)
        self.att_2_att_h = nn.Linear(att_feat_size, att_hid_size)
        self.h_2_att_h = nn.Linear(rnn_size, att_hid_size)
        self.att_h_2_out = nn.Linear(att_hid_size, 1)
        self.dropout = nn.Dropout(drop_prob_lm)
        for m in (self.i2h, self.h2h, self.z2h, self.att_2_att_h, self.h_2_att_h, self.att_h_2_out):
            _uniform_(m)

    def forward(self, xt, att_seq, state):
        pre_h, pre_c = _last_layer(state[0]), _last_layer(state[1])
        z = _attention(self, pre_h, att_seq)
        G = linear([(xt, self.i2h), (pre_h, self.h2h), (z, self.z2h)], pre_h.shape[0], self.i2h.out_features)
        next_h, next_c = lstm_cell(G, _f32c(pre_c), self.maxout)
        next_h = _dropout(next_h, self.drop_prob_lm, self.training)
        return next_h, (next_h.unsqueeze(0), next_c.unsqueeze(0))


class LSTMSoftAttentionNoInputCore(nn.Module):
    """ReviewNet's review cell (misc/LSTMSoftAttentionNoInputCore.py:10-97): the J=1 special case of
    LSTMFusionNoInputCore with h2h(pre_h) in place of H2h(H).  mil_feats / matching_feats
    are accepted and ignored exactly as in the reference."""

    def __init__(self, rnn_size, att_feat_size, att_num, att_hid_size, drop_prob_lm, maxout=0):
        super().__init__()
        self.rnn_size, self.drop_prob_lm = rnn_size, drop_prob_lm
        self.att_feat_size, self.att_num, self.att_hid_size, self.maxout = att_feat_size, att_num, att_hid_size, maxout
        gw = (5 if maxout else 4) * rnn_size                   # :23-28
        self.h2h = nn.Linear(rnn_size, gw)
        self.z2h = nn.Linear(att_feat_size, gw)
        self.att_2_att_h = nn.Linear(att_feat_size, att_hid_size)
        self.h_2_att_h = nn.Linear(rnn_size, att_hid_size)
        self.att_h_2_out = nn.Linear(att_hid_size, 1)
        self.dropout = nn.Dropout(drop_prob_lm)
        for m in (self.h2h, self.z2h, self.att_2_att_h, self.h_2_att_h, self.att_h_2_out):
            _uniform_(m)
        self.h2h.bias.data.fill_(-1)                   # :40-42
        self.z2h.bias.data.fill_(-1)

    def forward(self, att_seq, mil_feats, matching_feats, state):
        pre_h, pre_c = _last_layer(state[0]), _last_layer(state[1])
        z = _attention(self, pre_h, att_seq)
        G = linear([(pre_h, self.h2h), (z, self.z2h)], pre_h.shape[0], self.h2h.out_features)
        next_h, next_c = lstm_cell(G, _f32c(pre_c), self.maxout)
        next_h = _dropout(next_h, self.drop_prob_lm, self.training)
        return next_h, (next_h.unsqueeze(0), next_c.unsqueeze(0))


# --------------------------------------------------------------------------------------------------
# lazily materialised per-image views (building 5000 Python lists eagerly would dominate decode time)
# --------------------------------------------------------------------------------------------------
class _ReasonPredBatch(Sequence):
    """reason_pred_batch[k] -> list of J+1 tensors (beam, K), as sample_beam returns per image."""

    def __init__(self, reason: torch.Tensor, beam: int):
        self._r, self._beam = reason, beam   # (J+1, images, K)

    def __len__(self):
        return self._r.shape[1]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        return [self._r[j, k].unsqueeze(0).expand(self._beam, -1) for j in range(self._r.shape[0])]


class _TopList(Sequence):
    """top_seq / top_prob of sample_beam (:533-541): per image the finished beams sorted by -p, as a (n_done, L) int64
    tensor (kind 'seq') or a list of floats (kind 'prob').  Lazy: building 5000 Python objects eagerly costs ~15 ms per
    decode call, a twentieth of the whole end-to-end call."""

    def __init__(self, data, n_done, kind):
        self._d, self._n, self._kind = data, n_done, kind

    def __len__(self):
        return len(self._n)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        n = int(self._n[k])
        return self._d[k, :n].long() if self._kind == "seq" else self._d[k, :n].tolist()


class _DoneBeams(Sequence):
    """done_beams[k] -> list of {'seq','logps','p'} sorted by -p (misc/RecurrentFusionModel.py:529)."""

    def __init__(self, done_seq, done_lp, done_p, n_done):
        self._s, self._l, self._p, self._n = done_seq, done_lp, done_p, n_done   # done_lp stays on the device

    def __len__(self):
        return len(self._n)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if self._l.is_cuda:
            self._l = self._l.cpu()   # copied on first access only
        return [{"seq": self._s[k, i].long(), "logps": self._l[k, i], "p": float(self._p[k, i])} for i in range(int(self._n[k]))]


# --------------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------------
class RecurrentFusionModel(nn.Module):
    """misc/RecurrentFusionModel.py:117-658 with the per-timestep work in librfn_b200.so."""

    #: images decoded per device call (bounds the workspace: logits are rows x 9488 floats)
    chunk_images = 1024
    #: training only: set to seq_per_img (5) when the batch holds that many consecutive replicas of every image
    #: (dataloader.py:251-252) and stage-1/2 dropout is 0; stages 1-2 then run once per image
    dedup_rows = 1
    #: training only, with dedup_rows = g > 1: the feature tensors passed in hold ONE row per image (what
    #: ingest.FeatureIngest ships over PCIe) while labels hold g consecutive rows per image
    unique_feature_rows = False
    #: training only: gradient-enabled calls go through the hand-scheduled multi-stream tape (tape.py) whenever stage-1 /
    #: stage-2 dropout is off and (for forward()) scheduled sampling is off; False selects the op-by-op tape (autograd.py)
    fused_tape = True

    def __init__(self, opt):
        super().__init__()
        self.vocab_size = opt.vocab_size
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_type = opt.rnn_type
        self.rnn_size = opt.rnn_size
        self.num_layers = opt.num_layers
        self.drop_prob_lm = opt.drop_prob_lm
        self.drop_prob_reason = opt.drop_prob_reason
        self.drop_prob_fusion = opt.drop_prob_fusion
        self.seq_length = opt.seq_length
        self.num_review_steps = opt.num_review_steps
        self.num_review_steps_0 = opt.num_review_steps_0
        self.top_words_count = opt.top_words_count
        self.att_hid_size = opt.att_hid_size
        self.ss_prob = 0.0
        self.review_maxout = opt.review_maxout
        self.decoder_maxout = opt.maxout
        self.fusion_maxout = opt.fusion_maxout
        self.use_cuda = opt.use_cuda
        self.feat_array_info = opt.feat_array_info
        self.num_feat_array = len(self.feat_array_info)
        self.fc_feat_size = [f["fc_feat_size"] for f in self.feat_array_info]
        self.att_feat_size = [f["att_feat_size"] for f in self.feat_array_info]
        self.att_num = [f["att_num"] for f in self.feat_array_info]
        J, R = self.num_feat_array, self.rnn_size

        self.fc2h = nn.ModuleList([nn.Linear(self.fc_feat_size[i], R) for i in range(J)])
        self.embed = nn.Embedding(self.vocab_size + 1, self.input_encoding_size)
        self.logit = nn.Linear(R, self.vocab_size + 1)
        self.review_steps_individual = nn.ModuleList([
            FeatArrayFusionNoInputCore(J, R, self.att_feat_size, self.att_num, self.att_hid_size,
                                       self.drop_prob_fusion, self.fusion_maxout)
            for _ in range(self.num_review_steps_0)])
        self.reason_linear_individual = nn.ModuleList([nn.Linear(R, self.top_words_count) for _ in range(J)])
        self.review_steps = nn.ModuleList([
            LSTMSoftMultiAttentionFeatArrayNoInputCore(R, [R] * J, [self.num_review_steps_0] * J, self.att_hid_size,
                                                       self.drop_prob_reason, self.review_maxout)
            for _ in range(self.num_review_steps)])
        self.reason_linear = nn.Linear(R, self.top_words_count)
        self.decoder = LSTMSoftAttentionCore(self.input_encoding_size, R, R, self.num_review_steps,
                                             self.att_hid_size, self.drop_prob_lm, self.decoder_maxout)
        self.init_weights()
        self.done_beams = []
        self._dims = _capi.make_dims(list(zip(self.att_num, self.att_feat_size, self.fc_feat_size)), R,
                                     self.att_hid_size, self.input_encoding_size, self.vocab_size + 1,
                                     self.top_words_count, self.num_review_steps_0, self.num_review_steps,
                                     self.seq_length, review_maxout=self.review_maxout, decoder_maxout=self.decoder_maxout)
        self._pcache = None
        self._wsobj = _Workspace()   # this model's scratch for the C ABI (grow-only; swapped by graphs.GraphedBeamSearch)

    def init_weights(self):  # :188-196
        self.embed.weight.data.uniform_(-_INIT, _INIT)
        self.logit.weight.data.uniform_(-_INIT, _INIT)
        self.logit.bias.data.fill_(0)
        self.reason_linear.weight.data.uniform_(-_INIT, _INIT)
        for i in range(self.num_feat_array):
            self.reason_linear_individual[i].weight.data.uniform_(-_INIT, _INIT)
            self.fc2h[i].weight.data.uniform_(-_INIT, _INIT)

    # ---- plumbing ------------------------------------------------------------------------------
    #: keep a split copy of the weights for the fp16x3 / bf16 engines (modes 4 / 5) between inference calls; rebuilt
    #: whenever a parameter's storage or version changes.  Costs about the size of the fp32 model in HBM.
    weight_cache = True

    def _params(self):
        """HOST array of device pointers in state_dict order plus the trailing weight-cache slot (rfn_num_param_slots);
        cached until a parameter's storage moves, the engine mode changes, or -- for the weight cache -- a parameter is
        modified (its autograd version counter changes)."""
        ps = self._plist()
        mode = lib().rfn_get_gemm_mode()
        use_cache = bool(self.weight_cache) and mode >= 4 and not torch.is_grad_enabled()
        key = (tuple(p.data_ptr() for p in ps), mode, use_cache,
               (tuple(p._version for p in ps), _capi.WEIGHTS_EPOCH[0]) if use_cache else None)
        if self._pcache is None or self._pcache[0] != key:
            n = lib().rfn_num_params(C.byref(self._dims))
            if n != len(ps):
                raise _capi.RfnError(f"parameter count {len(ps)} != rfn_num_params {n}")
            for p in ps:
                _require_cuda(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise _capi.RfnError("parameters must be contiguous fp32")
            arr = ptr_array(ps + [None])
            wcache = None
            if use_cache:
                bf16 = 1 if mode == 5 else 0
                nbytes = lib().rfn_wcache_bytes(C.byref(self._dims), bf16)
                wcache = getattr(self, "_wcache", None)
                if wcache is None or wcache.numel() < nbytes or wcache.device != ps[0].device:
                    wcache = torch.empty(nbytes, dtype=torch.uint8, device=ps[0].device)
                check(lib().rfn_wcache_build(C.byref(self._dims), arr, bf16, ptr(wcache), wcache.numel(), stream()),
                      "rfn_wcache_build")
                arr[len(ps)] = ptr(wcache)
            self._wcache = wcache
            self._pcache = (key, arr)
        return self._pcache[1]

    def _plist(self):
        """The parameters in state_dict order.  Walking the module tree costs milliseconds for the 773-tensor model, which is
        what a batch-16 decode takes on the device: the Parameter objects survive .cuda() / load_state_dict, so the list is
        built once."""
        pl = self.__dict__.get("_plist_cache")
        if pl is None:
            pl = list(self.parameters())
            self.__dict__["_plist_cache"] = pl
        return pl

    def _apply(self, fn, *args, **kwargs):          # .cuda() / .to() / .float(): drop the derived caches
        self.__dict__.pop("_plist_cache", None)
        self._pcache = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):     # assign=True swaps the Parameter objects
        self.__dict__.pop("_plist_cache", None)
        self._pcache = None
        return super().load_state_dict(*args, **kwargs)

    def _dropout_active(self, p):
        return self.training and p > 0

    def _needs_tape(self):
        """Gradient recording (or train-mode dropout) sends the call through training.py's per-op loop."""
        if self._dropout_active(self.drop_prob_fusion) or self._dropout_active(self.drop_prob_reason) or \
                self._dropout_active(self.drop_prob_lm):
            return True
        return torch.is_grad_enabled() and any(p.requires_grad for p in self._plist())

    def _check_feats(self, fc_feats, att_feats, allow_host=False):
        J = self.num_feat_array
        if len(fc_feats) != J or len(att_feats) != J:
            raise _capi.RfnError(f"expected {J} fc / att feature tensors")
        _require_cuda(self._plist()[0])
        fc = [_f32c(t) for t in fc_feats]
        att = [_f32c(t) for t in att_feats]
        rows = fc[0].shape[0]
        for j in range(J):
            if not allow_host:
                _require_cuda(fc[j]); _require_cuda(att[j])
            if tuple(fc[j].shape) != (rows, self.fc_feat_size[j]) or \
                    tuple(att[j].shape) != (rows, self.att_num[j], self.att_feat_size[j]):
                raise _capi.RfnError(f"encoder {j}: feature shapes {tuple(fc[j].shape)} / {tuple(att[j].shape)} "
                                     f"do not match feat_array_info")
        return fc, att, rows

    def _ws(self, rows, dec_rows, device):
        n = lib().rfn_workspace_bytes(C.byref(self._dims), rows, dec_rows)
        if n == 0:
            raise _capi.RfnError("rfn_workspace_bytes: " + lib().rfn_last_error().decode())
        return self._wsobj.get(n, device)

    def _thought_vectors(self, fc, att, rows, init_state=None, want_reason=True, dec_rows=None, want_tv=False):
        """Stages 1-2 on `rows` feature rows -> TVc (rows,S1,R), reason (J+1,rows,K) or None, h, c (rows,R)."""
        dev = att[0].device
        R, S0, S1, K, J = self.rnn_size, self.num_review_steps_0, self.num_review_steps, self.top_words_count, self.num_feat_array
        TVc = torch.empty(rows, S1, R, dtype=torch.float32, device=dev)
        h = torch.empty(rows, R, dtype=torch.float32, device=dev)
        c = torch.empty(rows, R, dtype=torch.float32, device=dev)
        reason = torch.empty(J + 1, rows, K, dtype=torch.float32, device=dev) if want_reason else None
        TV = torch.empty(J, rows, S0, R, dtype=torch.float32, device=dev) if want_tv else None
        ws = self._ws(rows, dec_rows if dec_rows is not None else rows, dev)
        if init_state is None:
            fca, iha, ica = ptr_array(fc), None, None
        else:
            ih = [_f32c(s[0][-1]) for s in init_state]
            ic = [_f32c(s[1][-1]) for s in init_state]
            fca, iha, ica = None, ptr_array(ih), ptr_array(ic)
        check(lib().rfn_thought_vectors(C.byref(self._dims), self._params(), fca, iha, ica, ptr_array(att), rows,
                                        ptr(TVc), ptr(h), ptr(c), ptr(TV), ptr(reason), ptr(ws), ws.numel(), stream()),
              "rfn_thought_vectors")
        if want_tv:
            return TVc, reason, h, c, TV
        return TVc, reason, h, c

    @staticmethod
    def _reason_list(reason):
        # the reference squeezes (rows,K) tensors (:304, :328): a no-op unless rows == 1
        return [reason[j].squeeze() for j in range(reason.shape[0])]

    def _inference_guard(self, what):
        if self._dropout_active(self.drop_prob_fusion) or self._dropout_active(self.drop_prob_reason) or \
                self._dropout_active(self.drop_prob_lm):
            raise _capi.RfnError(f"{what}: dropout in training mode runs through the autograd path "
                                 "(recurrent_fusion_network_b200.training); call model.eval() for inference")

    # ---- reference surface ---------------------------------------------------------------------
    def get_init_state(self, fc_feats):  # :333-343
        state_list = []
        for i in range(self.num_feat_array):
            x = _f32c(fc_feats[i])
            init_h = linear([(x, self.fc2h[i])], x.shape[0], self.rnn_size).unsqueeze(0)
            state_list.append((init_h, init_h.clone()))
        return state_list

    def get_thought_vectors(self, fc_feats, att_feats, state_list):  # :283-331
        self._inference_guard("get_thought_vectors")
        fc, att, rows = self._check_feats(fc_feats, att_feats)
        TVc, reason, h, c = self._thought_vectors(fc, att, rows, init_state=state_list)
        return TVc, self._reason_list(reason), (h.unsqueeze(0), c.unsqueeze(0))

    def one_time_step(self, xt, fc_feats, thought_vectors_comb, state_decode):  # :345-350 -> LOGITS
        xt, TVc = _f32c(xt), _f32c(thought_vectors_comb)
        _require_cuda(xt)
        rows = xt.shape[0]
        h_in, c_in = _f32c(state_decode[0][-1]), _f32c(state_decode[1][-1])
        h = torch.empty_like(h_in)
        c = torch.empty_like(c_in)
        logits = torch.empty(rows, self.vocab_size + 1, dtype=torch.float32, device=xt.device)
        ws = self._ws(rows, rows, xt.device)
        check(lib().rfn_one_time_step(C.byref(self._dims), self._params(), ptr(xt), ptr(TVc), 1, ptr(h_in), ptr(c_in),
                                      ptr(h), ptr(c), ptr(logits), rows, ptr(ws), ws.numel(), stream()),
              "rfn_one_time_step")
        return logits, (h.unsqueeze(0), c.unsqueeze(0))

    def forward(self, fc_feats, att_feats, seq):  # :198-281
        if self._needs_tape() or self.ss_prob > 0:
            from . import training
            return training.forward_xe(self, fc_feats, att_feats, seq)
        self._inference_guard("forward")
        fc, att, rows = self._check_feats(fc_feats, att_feats)
        seq = seq.to(device=fc[0].device, dtype=torch.int64).contiguous()
        # T' = index of the first all-zero column i >= 1 (:274-275)
        colsum = (seq != 0).any(dim=0).cpu().tolist()
        T = seq.shape[1]
        for i in range(1, seq.shape[1]):
            if not colsum[i]:
                T = i
                break
        TVc, reason, h, c = self._thought_vectors(fc, att, rows)
        out = torch.empty(rows, T, self.vocab_size + 1, dtype=torch.float32, device=fc[0].device)
        ws = self._ws(rows, rows, fc[0].device)
        check(lib().rfn_decode_teacher_forced(C.byref(self._dims), self._params(), ptr(TVc), ptr(h), ptr(c), ptr(seq),
                                              seq.stride(0), T, rows, ptr(out), ptr(ws), ws.numel(), stream()),
              "rfn_decode_teacher_forced")
        return out, self._reason_list(reason)

    def sample(self, fc_feats, att_feats, opt={}):  # :545-658
        sample_max = opt.get("sample_max", 1)
        beam_size = opt.get("beam_size", 1)
        temperature = opt.get("temperature", 1.0)
        if beam_size > 1:
            return self.sample_beam(fc_feats, att_feats, opt)
        if self._needs_tape():
            from . import training
            return training.sample_with_grad(self, fc_feats, att_feats, opt)
        self._inference_guard("sample")
        seq, slp, lp_all, reason, dT = self.sample_device(fc_feats, att_feats, opt)
        want_all = lp_all is not None
        T = int(dT.item())   # the reference's early break (:645) -- one 4-byte read per call
        if T == 0:
            raise RuntimeError("sample(): every row emitted <eos> at t=1; the reference fails here too "
                               "(torch.cat of an empty list, misc/RecurrentFusionModel.py:655)")
        return seq[:, :T], slp[:, :T], (lp_all[:, :T + 1] if want_all else None), self._reason_list(reason)

    def sample_device(self, fc_feats, att_feats, opt={}):
        """The device part of sample() (greedy / multinomial), free of host synchronisation and therefore capturable in a
        CUDA graph: full-width (rows, L) tokens / log-probs, (rows, L + 1, V1) log-prob table (or None), reason_pred
        (J+1, rows, K) and a device int32 T = number of valid columns (the reference's early break, :645)."""
        sample_max = opt.get("sample_max", 1)
        temperature = opt.get("temperature", 1.0)
        with torch.no_grad():
            fc, att, rows = self._check_feats(fc_feats, att_feats)
            dev = fc[0].device
            L, V1 = self.seq_length, self.vocab_size + 1
            uniforms = None
            if not sample_max:
                uniforms = opt.get("uniforms")
                if uniforms is None:   # the reference draws on the CPU RNG (:624-631); we take torch's CUDA generator
                    uniforms = torch.rand(rows, L, device=dev, dtype=torch.float32)
                uniforms = _f32c(uniforms.to(dev))
            seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
            slp = torch.empty(rows, L, dtype=torch.float32, device=dev)
            want_all = opt.get("return_logprobs_all", True)
            lp_all = torch.empty(rows, L + 1, V1, dtype=torch.float32, device=dev) if want_all else None
            dT = torch.zeros(1, dtype=torch.int32, device=dev)
            TVc, reason, h, c = self._thought_vectors(fc, att, rows)
            ws = self._ws(rows, rows, dev)
            check(lib().rfn_decode_sample(C.byref(self._dims), self._params(), ptr(TVc), ptr(h), ptr(c), rows,
                                          ptr(uniforms), float(temperature), ptr(seq), ptr(slp), ptr(lp_all), ptr(dT),
                                          ptr(ws), ws.numel(), stream()), "rfn_decode_sample")
        return seq, slp, lp_all, reason, dT

    def _device(self):
        return self._plist()[0].device

    def _chunks(self, fc, att, rows, step):
        """Yields (k0, k1, fc_chunk, att_chunk) with the chunk resident on the device.  Host (pinned)
        inputs are streamed over PCIe on a side stream, double buffered, so the copy of chunk i+1
        overlaps the decode of chunk i (dataloader.py hands the reference host arrays, train.py:116-133)."""
        if fc[0].is_cuda:
            for k0 in range(0, rows, step):
                k1 = min(rows, k0 + step)
                yield k0, k1, [f[k0:k1] for f in fc], [a[k0:k1] for a in att]
            return
        dev = self._device()
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        # equal chunks (a short tail chunk would fall below the row counts the tensor engines take: < 128 rows run on the
        # fp32 SIMT kernels, an order of magnitude slower per row)
        n_chunks = (rows + step - 1) // step
        step = (rows + n_chunks - 1) // n_chunks
        n_buf = min(step, rows)
        key = (n_buf, dev)
        if getattr(self, "_staging_key", None) != key:
            self._staging = [([torch.empty(n_buf, *f.shape[1:], dtype=torch.float32, device=dev) for f in fc],
                              [torch.empty(n_buf, *a.shape[1:], dtype=torch.float32, device=dev) for a in att])
                             for _ in range(2)]
            self._staging_key = key
        staging = self._staging
        # (tapering the final chunks to shorten the exposed last decode was measured: 352 ms against 336 ms per 5000
        # images -- small chunks decode less efficiently than they copy)
        spans = [(k0, min(rows, k0 + step)) for k0 in range(0, rows, step)]
        ready, free = [None, None], [None, None]

        def issue(i):
            k0, k1 = spans[i]
            b = i % 2
            with torch.cuda.stream(cs):
                if free[b] is not None:
                    cs.wait_event(free[b])
                else:
                    cs.wait_stream(cur)
                for j in range(len(fc)):
                    staging[b][0][j][:k1 - k0].copy_(fc[j][k0:k1], non_blocking=True)
                    staging[b][1][j][:k1 - k0].copy_(att[j][k0:k1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                ready[b] = ev

        issue(0)
        for i, (k0, k1) in enumerate(spans):
            if i + 1 < len(spans):
                issue(i + 1)
            b = i % 2
            cur.wait_event(ready[b])
            yield k0, k1, [t[:k1 - k0] for t in staging[b][0]], [t[:k1 - k0] for t in staging[b][1]]
            ev = torch.cuda.Event()
            ev.record(cur)
            free[b] = ev

    def beam_search(self, fc_feats, att_feats, beam_size=3, want_reason=True):
        """Batched device beam search returning DEVICE tensors only:
        (seq (B,L) i64, seqLogprobs (B,L), done_seq (B,beam*L,L) i32, done_logps, done_p (B,beam*L), n_done (B),
        reason_pred (J+1,B,K) or None).  sample_beam() wraps this into the reference's return structure."""
        with torch.no_grad():
            fc, att, rows = self._check_feats(fc_feats, att_feats, allow_host=True)
            return self._beam_tensors(fc, att, rows, beam_size, want_reason)

    def _beam_tensors(self, fc, att, rows, beam_size, want_reason=True):
        dev = self._device()
        L, K, J = self.seq_length, self.top_words_count, self.num_feat_array
        cap = beam_size * L
        seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
        slp = torch.empty(rows, L, dtype=torch.float32, device=dev)
        done_seq = torch.zeros(rows, cap, L, dtype=torch.int32, device=dev)
        done_lp = torch.zeros(rows, cap, L, dtype=torch.float32, device=dev)
        done_p = torch.full((rows, cap), float("nan"), dtype=torch.float32, device=dev)
        n_done = torch.zeros(rows, dtype=torch.int32, device=dev)
        reason = torch.empty(J + 1, rows, K, dtype=torch.float32, device=dev) if want_reason else None
        step = max(1, int(self.chunk_images))
        for k0, k1, fck, attk in self._chunks(fc, att, rows, step):
            n = k1 - k0
            TVc, rsn, h, c = self._thought_vectors(fck, attk, n, want_reason=want_reason, dec_rows=n * beam_size)
            if want_reason:
                reason[:, k0:k1] = rsn
            ws = self._ws(n, n * beam_size, dev)
            check(lib().rfn_decode_beam(C.byref(self._dims), self._params(), ptr(TVc), ptr(h), ptr(c), n, beam_size,
                                        ptr(seq[k0:k1]), ptr(slp[k0:k1]), ptr(done_seq[k0:k1]), ptr(done_lp[k0:k1]),
                                        ptr(done_p[k0:k1]), ptr(n_done[k0:k1]), ptr(ws), ws.numel(), stream()),
                  "rfn_decode_beam")
        return seq, slp, done_seq, done_lp, done_p, n_done, reason

    def sample_beam(self, fc_feats, att_feats, opt={}):  # :352-543
        beam_size = opt.get("beam_size", 10)
        assert beam_size <= self.vocab_size + 1, "lets assume this for now"   # :360
        if beam_size > _capi.MAX_BEAM:
            raise _capi.RfnError(f"beam_size {beam_size} > {_capi.MAX_BEAM} is not built")
        self._inference_guard("sample_beam")
        with torch.no_grad():
            fc, att, rows = self._check_feats(fc_feats, att_feats, allow_host=True)
            seq, slp, done_seq, done_lp, done_p, n_done, reason = self._beam_tensors(fc, att, rows, beam_size)
            # the caption gather: one D2H of the finished-beam lists
            n_cpu = n_done.cpu()
            ds_cpu = done_seq.cpu()
            dp_cpu = done_p.cpu()
            nl = n_cpu.tolist()
            self.done_beams = _DoneBeams(ds_cpu, done_lp, dp_cpu, nl)
            return seq, slp, _TopList(ds_cpu, nl, "seq"), _TopList(dp_cpu, nl, "prob"), _ReasonPredBatch(reason, beam_size)
