"""Gradient-enabled versions of the path (train.py:154-163, train_rl.py:160-191): the same Python loop
structure as the reference's forward()/sample(), every operation an autograd.Function over our kernels."""
from __future__ import annotations

import torch

from . import autograd as AG


def thought_vectors(model, att, state_list):
    """Stages 1-2 with the tape on (misc/RecurrentFusionModel.py:283-331)."""
    J = model.num_feat_array
    tv_list = [[] for _ in range(J)]
    reason_mat = [[] for _ in range(J)]
    for i in range(model.num_review_steps_0):
        output_list, state_list = model.review_steps_individual[i](att, state_list)
        for j in range(J):
            tv_list[j].append(output_list[j])
            reason_mat[j].append(AG.linear([(output_list[j], model.reason_linear_individual[j])]))
    thought_vectors_ = [torch.stack(tv_list[j], 1).contiguous() for j in range(J)]
    reason_pred = [AG.MaxOverStepsFn.apply(torch.stack(reason_mat[j], 1).contiguous()) for j in range(J)]
    h = AG.MeanFn.apply(*[st[0].squeeze(0) for st in state_list])
    c = AG.MeanFn.apply(*[st[1].squeeze(0) for st in state_list])
    state = (h.unsqueeze(0), c.unsqueeze(0))
    comb, reason_comb = [], []
    for i in range(model.num_review_steps):
        output, state = model.review_steps[i](thought_vectors_, state)
        comb.append(output)
        reason_comb.append(AG.linear([(output, model.reason_linear)]))
    TVc = torch.stack(comb, 1).contiguous()
    reason_pred.append(AG.MaxOverStepsFn.apply(torch.stack(reason_comb, 1).contiguous()))
    return TVc, reason_pred, state


def _stages(model, fc, att):
    """get_init_state + thought vectors; with model.dedup_rows = g > 1 (rows are g consecutive replicas of each
    image and no stage-1/2 dropout) the stages run on the unique rows only and their outputs are expanded."""
    g = int(getattr(model, "dedup_rows", 1) or 1)
    rows = fc[0].shape[0]
    AG.clear_transposed_cache()
    unique = bool(getattr(model, "unique_feature_rows", False)) and g > 1
    if unique and (model._dropout_active(model.drop_prob_fusion) or model._dropout_active(model.drop_prob_reason)):
        raise RuntimeError("unique_feature_rows needs drop_prob_fusion = drop_prob_reason = 0 in training mode (each replica "
                           "would draw its own dropout mask in the reference); pass FeatureBatch.expanded() instead")
    if unique or (g > 1 and rows % g == 0 and not (model._dropout_active(model.drop_prob_fusion) or
                                                   model._dropout_active(model.drop_prob_reason))):
        # unique: the caller ships one feature row per image (ingest.FeatureBatch.fc / .att) and g label rows per image
        fcu = fc if unique else [f[::g].contiguous() for f in fc]
        attu = att if unique else [a[::g].contiguous() for a in att]
        TVc, reason_pred, (h, c) = thought_vectors(model, attu, model.get_init_state(fcu))
        ex = lambda t: AG.ExpandRowsFn.apply(t, g)
        return ex(TVc), [ex(r) for r in reason_pred], (ex(h.squeeze(0)).unsqueeze(0), ex(c.squeeze(0)).unsqueeze(0))
    return thought_vectors(model, att, model.get_init_state(fc))


def _step(model, it, TVc, state, xt=None):
    if xt is None:
        xt = AG.EmbedFn.apply(it, model.embed.weight)
    output, state = model.decoder(xt, TVc, state)
    lp = AG.LogSoftmaxFn.apply(AG.linear([(output, model.logit)]))
    return lp, state


def forward_xe(model, fc_feats, att_feats, seq, col_any=None):
    """RecurrentFusionModel.forward with gradients (misc/RecurrentFusionModel.py:198-281).  `col_any` (which label
    columns hold a non-zero token) may be supplied by a caller that already knows it on the host, so that the
    call contains no device->host synchronisation (CUDA-graph capture)."""
    fc, att, rows = model._check_feats(fc_feats, att_feats)
    if getattr(model, "unique_feature_rows", False) and int(getattr(model, "dedup_rows", 1) or 1) > 1:
        rows *= int(model.dedup_rows)
    seq = seq.to(device=fc[0].device, dtype=torch.int64)
    if seq.shape[0] != rows:
        raise RuntimeError(f"labels have {seq.shape[0]} rows, the features describe {rows}")
    TVc, reason_pred, state = _stages(model, fc, att)
    outputs = []
    if col_any is None:
        col_any = (seq != 0).any(dim=0).cpu().tolist()
    xts = None
    if model.ss_prob <= 0.0:
        # teacher forcing only: all input tokens are known, one embedding gather (and one dense dE in backward) for the
        # whole sequence instead of one per step
        T = seq.size(1)
        xts = AG.EmbedFn.apply(seq.t().reshape(-1), model.embed.weight).view(T, rows, -1).unbind(0)
    for i in range(seq.size(1)):
        it = seq[:, i].clone()
        if i >= 1 and model.ss_prob > 0.0:                              # scheduled sampling (:260-270)
            sample_mask = torch.rand(rows, device=it.device) < model.ss_prob
            if bool(sample_mask.any()):
                u = torch.rand(rows, device=it.device)
                sampled, _ = AG.select_token(outputs[-1], uniforms=u)
                it = torch.where(sample_mask, sampled, it)
        if i >= 1 and not col_any[i]:                                   # :274-275
            break
        lp, state = _step(model, it, TVc, state, xt=None if xts is None else xts[i])
        outputs.append(lp)
    return torch.stack(outputs, 1).contiguous(), [r.squeeze() for r in reason_pred]


def sample_with_grad(model, fc_feats, att_feats, opt):
    """RecurrentFusionModel.sample with gradients through the log-probs (train_rl.py:160;
    misc/RecurrentFusionModel.py:545-658)."""
    sample_max = opt.get("sample_max", 1)
    temperature = opt.get("temperature", 1.0)
    fc, att, rows = model._check_feats(fc_feats, att_feats)
    if getattr(model, "unique_feature_rows", False) and int(getattr(model, "dedup_rows", 1) or 1) > 1:
        rows *= int(model.dedup_rows)
    dev = fc[0].device
    L = model.seq_length
    uniforms = None
    if not sample_max:
        uniforms = opt.get("uniforms")
        if uniforms is None:
            uniforms = torch.rand(rows, L, device=dev)
        uniforms = uniforms.to(dev).float()
    TVc, reason_pred, state = _stages(model, fc, att)
    seq, slps, lp_all = [], [], []
    lp = None
    unfinished = None
    for t in range(L + 1):
        if t == 0:
            it = torch.zeros(rows, dtype=torch.int64, device=dev)
        else:
            it, _ = AG.select_token(lp, uniforms=None if sample_max else uniforms[:, t - 1].contiguous(),
                                    temperature=temperature)
            s_lp = AG.GatherColsFn.apply(lp, it)
        x_tok = it
        if t >= 1:
            unfinished = (it > 0) if t == 1 else unfinished & (it > 0)
            if not bool(unfinished.any()):
                break
            seq.append(it * unfinished.to(it.dtype))
            slps.append(s_lp)
        lp, state = _step(model, x_tok, TVc, state)
        lp_all.append(lp)
    if not seq:
        raise RuntimeError("sample(): every row emitted <eos> at t=1 (the reference fails here too)")
    return (torch.stack(seq, 1), torch.stack(slps, 1), torch.stack(lp_all, 1).contiguous(),
            [r.squeeze() for r in reason_pred])


def xe_criterion(crit, log_prob, target, mask, top_pred, top_true, reason_weight):
    eps = float(crit.label_smoothing_epsilon) if crit.use_label_smoothing else 0.0
    terms = [AG.XeLossFn.apply(log_prob, target, mask, eps)]
    for p in top_pred:
        terms.append(AG.MarginFn.apply(p, top_true, reason_weight / len(top_pred)))
    return AG.AddScalarsFn.apply(*terms)[0]


def rl_criterion(crit, input, seq, reward, logprobs_all, entropy_reg, top_pred, top_true, reason_weight):
    terms = [AG.RlLossFn.apply(input, seq, reward, logprobs_all, entropy_reg)]
    for p in top_pred:
        terms.append(AG.MarginFn.apply(p, top_true, reason_weight / len(top_pred)))
    return AG.AddScalarsFn.apply(*terms)[0]



class GraphedXEStep:
    """One XE training step (train.py:154-163: zero_grad, forward, criterion, backward, clip_gradient, Adam) captured as
    CUDA graphs and replayed: the eager step issues ~2,600 kernels from Python and is bound by the host, not the device.

    Two graphs -- forward+backward, and the optimizer -- so that the data-parallel gradient all-reduce
    (dist.average_gradients, `between`) can run between them; or pass grad_sync=dist.OverlappedGradSync(params) to have
    the all-reduce captured inside the first graph, overlapped with the backward pass.  Inputs are copied into static buffers before each replay.  The decode
    loop is captured over seq_length + 1 label columns (the reference stops at the first all-zero column, :274-275,
    which is at most that many; columns it would have skipped carry mask 0 and contribute nothing to the loss or the
    gradients).  Needs ss_prob == 0 (scheduled sampling
    reads a mask back to the host) and a FusedAdam(capturable=True)."""

    def __init__(self, model, crit, optimizer, fc, att, labels, masks, top_true, reason_weight, warmup=3, between=None,
                 grad_sync=None):
        if model.ss_prob > 0.0:
            raise RuntimeError("GraphedXEStep: scheduled sampling is not capturable (ss_prob must be 0)")
        if not all(g.get("capturable") for g in optimizer.param_groups):
            raise RuntimeError("GraphedXEStep needs FusedAdam(capturable=True)")
        self.model, self.crit, self.opt, self.between = model, crit, optimizer, between
        self.reason_weight = float(reason_weight)
        self.fc = [f.clone() for f in fc]
        self.att = [a.clone() for a in att]
        self.labels, self.masks, self.top = labels.clone(), masks.clone(), top_true.clone()
        self.col_any = [True] * (labels.size(1) - 1) + [False]   # the last label column is always <pad>: stop there as :274-275 does
        self.loss = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        optimizer.init_state()
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # allocator and lazy kernel attributes; parameters are not touched
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_fb = torch.cuda.CUDAGraph()
        # grad_sync (dist.OverlappedGradSync): the data-parallel all-reduce is captured INSIDE the forward+backward graph,
        # bucket by bucket on a side stream while the backward of the earlier layers still runs (then `between` is unused)
        with torch.cuda.graph(self.g_fb):
            if grad_sync is not None:
                grad_sync.install()
            self.loss = self._fwd_bwd()
            if grad_sync is not None:
                grad_sync.finish()
        if grad_sync is not None:
            grad_sync.remove()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            optimizer.step()

    def _fwd_bwd(self):
        self.opt.zero_grad(set_to_none=True)
        lp, rp = forward_xe(self.model, self.fc, self.att, self.labels, col_any=self.col_any)
        loss = self.crit(lp, self.labels[:, 1:], self.masks[:, 1:], rp, self.top, self.reason_weight)
        loss.backward()
        return loss.detach()

    def __call__(self, fc=None, att=None, labels=None, masks=None, top_true=None):
        """Replays the step on new data (same shapes); returns the loss tensor (device, overwritten by the next call)."""
        if fc is not None:
            for d, s in zip(self.fc, fc):
                d.copy_(s, non_blocking=True)
            for d, s in zip(self.att, att):
                d.copy_(s, non_blocking=True)
            self.labels.copy_(labels, non_blocking=True)
            self.masks.copy_(masks, non_blocking=True)
            self.top.copy_(top_true, non_blocking=True)
        self.g_fb.replay()
        if self.between is not None:
            self.between()
        self.g_opt.replay()
        from . import _capi
        _capi.WEIGHTS_EPOCH[0] += 1   # the replayed optimizer kernel rewrote the parameters
        return self.loss
