"""Ensemble decoding (eval_utils.py:268-290, :387-719) with the models' REAL signatures (the reference's
ensemble code is stale): per step the M models' logits are averaged, then log_softmax; one
beam over tokens is shared and every model's LSTM state is forked with it.  All M models live on the
same GPU and images are sharded across GPUs, so no per-step inter-GPU copy exists."""
import ctypes as C

import torch

from . import _capi
from ._capi import check, lib, ptr, ptr_array, stream


def model_ensemble_feat_array_one_step(model_list, xt_list, state_list, thought_vector_list):
    """-> (logit_list, state_list_new, logprob)   (eval_utils.py:268-290)"""
    logit_list, state_new = [], []
    for m, model in enumerate(model_list):
        logit, st = model.one_time_step(xt_list[m], None, thought_vector_list[m], state_list[m])
        logit_list.append(logit)
        state_new.append(st)
    rows, V = logit_list[0].shape
    scratch = torch.empty_like(logit_list[0])
    lp = torch.empty_like(logit_list[0])
    check(lib().rfn_mean_log_softmax_f32(len(logit_list), ptr_array(logit_list), rows, V, ptr(scratch), ptr(lp), stream()),
          "rfn_mean_log_softmax_f32")
    return logit_list, state_new, lp


def ensemble_sample_beam(models, fc_feats, att_feats, opt={}):
    """Beam search over the logit-mean ensemble.  Returns (seq (B,L), seqLogprobs (B,L), top_seq, top_prob)."""
    beam = opt.get("beam_size", 3)
    m0 = models[0]
    M = len(models)
    if M > 8:
        raise _capi.RfnError("at most 8 ensemble members are supported")
    with torch.no_grad():
        fc, att, rows = m0._check_feats(fc_feats, att_feats)
        dev = fc[0].device
        L = m0.seq_length
        cap = beam * L
        seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
        slp = torch.empty(rows, L, dtype=torch.float32, device=dev)
        done_seq = torch.zeros(rows, cap, L, dtype=torch.int32, device=dev)
        done_lp = torch.zeros(rows, cap, L, dtype=torch.float32, device=dev)
        done_p = torch.full((rows, cap), float("nan"), dtype=torch.float32, device=dev)
        n_done = torch.zeros(rows, dtype=torch.int32, device=dev)
        step = max(1, int(m0.chunk_images) // M)
        for k0 in range(0, rows, step):
            k1 = min(rows, k0 + step)
            n = k1 - k0
            tv, hs, cs = [], [], []
            for model in models:
                TVc, _, h, c = model._thought_vectors([f[k0:k1] for f in fc], [a[k0:k1] for a in att], n, want_reason=False)
                tv.append(TVc); hs.append(h); cs.append(c)
            nbytes = lib().rfn_ensemble_workspace_bytes(C.byref(m0._dims), M, n, beam)
            ws = m0._wsobj.get(nbytes, dev)   # after the last _thought_vectors call above: the same buffer may be handed out
            pm = (C.POINTER(C.c_void_p) * M)(*[C.cast(model._params(), C.POINTER(C.c_void_p)) for model in models])
            check(lib().rfn_ensemble_decode_beam(C.byref(m0._dims), M, pm, ptr_array(tv), ptr_array(hs), ptr_array(cs), n,
                                                 beam, ptr(seq[k0:k1]), ptr(slp[k0:k1]), ptr(done_seq[k0:k1]),
                                                 ptr(done_lp[k0:k1]), ptr(done_p[k0:k1]), ptr(n_done[k0:k1]), ptr(ws),
                                                 ws.numel(), stream()), "rfn_ensemble_decode_beam")
        from .model import _TopList
        nl = n_done.cpu().tolist()
        return seq, slp, _TopList(done_seq.cpu(), nl, "seq"), _TopList(done_p.cpu(), nl, "prob")


def ensemble_sample_greedy(models, fc_feats, att_feats, opt={}):
    """Greedy decode of the logit-mean ensemble (eval_utils.py:729-975): all images of the batch advance together.
    Returns (seq (B,T) int64, seqLogprobs (B,T)) with T <= seq_length as the reference's early break leaves it."""
    m0 = models[0]
    M = len(models)
    if M > 8:
        raise _capi.RfnError("at most 8 ensemble members are supported")
    with torch.no_grad():
        fc, att, rows = m0._check_feats(fc_feats, att_feats)
        dev = fc[0].device
        L = m0.seq_length
        seq = torch.zeros(rows, L, dtype=torch.int64, device=dev)
        slp = torch.zeros(rows, L, dtype=torch.float32, device=dev)
        dT = torch.zeros(1, dtype=torch.int32, device=dev)
        tv, hs, cs = [], [], []
        for model in models:
            TVc, _, h, c = model._thought_vectors(fc, att, rows, want_reason=False)
            tv.append(TVc); hs.append(h); cs.append(c)
        nbytes = lib().rfn_ensemble_workspace_bytes(C.byref(m0._dims), M, rows, 1)
        ws = m0._wsobj.get(nbytes, dev)
        pm = (C.POINTER(C.c_void_p) * M)(*[C.cast(model._params(), C.POINTER(C.c_void_p)) for model in models])
        check(lib().rfn_ensemble_decode_greedy(C.byref(m0._dims), M, pm, ptr_array(tv), ptr_array(hs), ptr_array(cs), rows,
                                               ptr(seq), ptr(slp), ptr(dT), ptr(ws), ws.numel(), stream()),
              "rfn_ensemble_decode_greedy")
        T = int(dT.item())
        if T == 0:
            raise RuntimeError("ensemble greedy: every row emitted <eos> at t=1; the reference fails here too "
                               "(torch.cat of an empty list, eval_utils.py:947)")
        return seq[:, :T], slp[:, :T]
