"""Gradient-enabled versions of the path (train.py:154-163, train_rl.py:160-191): the same Python loop
structure as the reference's forward()/sample(), every operation an autograd.Function over our kernels."""
from __future__ import annotations

import torch

from . import autograd as AG
from . import tape


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
        if tape.usable(model):
            TVc, reason_pred, (h, c) = tape.thought_vectors(model, fcu, attu)
        else:
            TVc, reason_pred, (h, c) = thought_vectors(model, attu, model.get_init_state(fcu))
        ex = lambda t: AG.ExpandRowsFn.apply(t, g)
        return ex(TVc), [ex(r) for r in reason_pred], (ex(h.squeeze(0)).unsqueeze(0), ex(c.squeeze(0)).unsqueeze(0))
    if tape.usable(model):      # hand-scheduled multi-stream tape (tape.py); the per-op tape below is the general case
        return tape.thought_vectors(model, fc, att)
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
    if model.ss_prob <= 0.0 and tape.usable(model):
        # teacher forcing only: the whole decoder loop is one Function (T-batched weight gradients, hoisted invariants)
        Tn = seq.size(1)
        for i in range(1, seq.size(1)):
            if not col_any[i]:                                          # :274-275
                Tn = i
                break
        lp = tape.decode_teacher_forced(model, seq[:, :Tn], TVc, state)
        return lp, [r.squeeze() for r in reason_pred]
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



def _decode_tokens(model, TVc, h, c, uniforms, temperature, second_workspace=False):
    """Tokens of a greedy (uniforms None) or multinomial decode from given stage-2 outputs, all on the device and without
    gradients (rfn_decode_sample): seq (rows, L) int64 with finished rows zeroed (:647), T on the device.
    second_workspace: scratch of its own, for a decode that runs side by side with another one of the same model."""
    import ctypes as C
    from ._capi import check, lib, ptr, stream
    rows, dev, L = h.shape[0], h.device, model.seq_length
    seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
    slp = torch.empty(rows, L, dtype=torch.float32, device=dev)
    dT = torch.zeros(1, dtype=torch.int32, device=dev)
    if second_workspace:
        from .model import _Workspace
        wsb = model.__dict__.get("_wsobj_b")
        if wsb is None:
            wsb = model.__dict__["_wsobj_b"] = _Workspace()
        ws = wsb.get(lib().rfn_workspace_bytes(C.byref(model._dims), rows, rows), dev)
    else:
        ws = model._ws(rows, rows, dev)
    check(lib().rfn_decode_sample(C.byref(model._dims), model._params(), ptr(TVc), ptr(h), ptr(c), rows, ptr(uniforms),
                                  float(temperature), ptr(seq), ptr(slp), None, ptr(dT), ptr(ws), ws.numel(), stream()),
          "rfn_decode_sample")
    return seq, dT


def rl_forward_loss(model, crit, fc_feats, att_feats, uniforms, reward_fn, top_true, reason_weight, entropy_reg=0.0,
                    temperature=1.0):
    """One self-critical iteration's forward (train_rl.py:150-169) without any host round trip:

      stages 1-2 with the tape on, ONCE                      (the reference runs them again for the greedy baseline,
                                                              get_rewards.py:119-124, on the same rows)
      multinomial decode + greedy baseline decode, no tape   (rfn_decode_sample on the detached thought vectors; the sampled
                                                              tokens are what :622-635 draws, given `uniforms`)
      reward = reward_fn(sampled, greedy) on the device      (reward.compute_reward_packed)
      teacher-forced taped decoder over the SAMPLED tokens   -> the log-probs the reference's sample() carries with grad
      ReviewNetRewardCriterion                               (misc/utils.py:50-84)

    Feeding the masked tokens (0 after <eos>) instead of the raw ones (:637 vs :647) changes only positions the criterion
    masks out (mask0 = seq > 0 for the entropy term, its right shift for the sampled-token term), so the loss and every
    gradient equal those of the reference-shaped step (tests/test_gpu_training.py).  Fixed length L: capturable.
    Needs drop_prob_lm inactive (a no-tape decode cannot replay the taped pass's dropout masks).
    Returns (loss, seq (rows, L), greedy (rows, L), reward (rows, L))."""
    if model._dropout_active(model.drop_prob_lm):
        raise RuntimeError("rl_forward_loss needs drop_prob_lm = 0 in training mode (the shipped RL scripts); use "
                           "model.sample(..., {'sample_max': 0}) + the criterion for the general case")
    fc, att, rows = model._check_feats(fc_feats, att_feats)
    if getattr(model, "unique_feature_rows", False) and int(getattr(model, "dedup_rows", 1) or 1) > 1:
        rows *= int(model.dedup_rows)
    L = model.seq_length
    tape.mark("rl_begin")
    TVc, reason_pred, state = _stages(model, fc, att)
    tape.mark("rl_stages_end")
    with torch.no_grad():
        TVd = TVc.detach().contiguous()
        hd, cd = state[0].detach().squeeze(0).contiguous(), state[1].detach().squeeze(0).contiguous()
        from ._capi import lib
        prev = lib().rfn_get_splitk()
        lib().rfn_set_splitk(1)      # a few hundred rows: split-K GEMMs fill the GPU (the tokens are samples and a baseline)
        # no split-weight cache for these decodes: the weights change with every optimizer step, and a cache built while a
        # CUDA graph of the iteration is captured would be replayed stale
        saved_wc, model.weight_cache = model.weight_cache, False
        try:
            # the two decodes are independent chains of small kernels (a few hundred rows): side by side on two streams
            main = torch.cuda.current_stream()
            side = tape._Streams.get(TVd.device, 1).enc[0]
            u = uniforms.to(TVd.device).float().contiguous()
            tape._after(side, main)
            seq, _ = _decode_tokens(model, TVd, hd, cd, u, temperature)
            with torch.cuda.stream(side):
                greedy, _ = _decode_tokens(model, TVd, hd, cd, None, 1.0, second_workspace=True)
            tape._after(main, side)
        finally:
            lib().rfn_set_splitk(prev)
            model.weight_cache = saved_wc
        tape.mark("rl_decodes_end")
        reward = reward_fn(seq, greedy)
        tape.mark("rl_reward_end")
        tokens = torch.cat([torch.zeros(rows, 1, dtype=torch.int64, device=seq.device), seq[:, :L - 1]], 1)
    if tape.usable(model):
        lp_all = tape.decode_teacher_forced(model, tokens, TVc, state)            # (rows, L, V) view of a time-major table
        lp_tm = lp_all.transpose(0, 1)                                            # (L, rows, V), contiguous again
        slp = AG.GatherColsFn.apply(lp_tm.reshape(L * rows, -1), seq.t().reshape(-1)).view(L, rows).t()
    else:
        xts = AG.EmbedFn.apply(tokens.t().reshape(-1), model.embed.weight).view(L, rows, -1).unbind(0)
        lps, slps = [], []
        for i in range(L):
            lp, state = _step(model, None, TVc, state, xt=xts[i])
            lps.append(lp)
            slps.append(AG.GatherColsFn.apply(lp, seq[:, i].contiguous()))
        lp_all = torch.stack(lps, 1).contiguous()
        slp = torch.stack(slps, 1)
    loss = rl_criterion(crit, slp, seq, reward, lp_all, entropy_reg, [r.squeeze() for r in reason_pred], top_true, reason_weight)
    tape.mark("rl_loss_end")
    return loss, seq, greedy, reward


class EarlyStep:
    """tape.GRAD_SINK for the graphed training steps: as soon as tape.Stage1Fn's backward has issued the weight gradients of a
    fusion step it (1) all-reduces them in place (data parallel: grad_sync = dist.OverlappedGradSync, on its NCCL side stream)
    and (2) applies the clamp + Adam update to those parameters on the same side stream (FusedAdam.apply_to), while the
    backward pass of the earlier fusion steps is still running.  Stage 1 holds 95 % of the parameters (distinct weights per
    step) and is the last thing a backward pass computes; nothing reads a step's weights after its backward, so the update
    cannot race with the rest of the pass.  What is left for optimizer.step() after the backward pass is the 5 % of the
    parameters the decoder, stage 2 and the heads own."""

    def __init__(self, optimizer, grad_sync=None):
        self.opt, self.sync = optimizer, grad_sync
        self.side = grad_sync.side if grad_sync is not None else torch.cuda.Stream()

    def install(self):
        if self.sync is not None:
            self.sync.install()
        tape.GRAD_SINK = self

    def remove(self):
        if self.sync is not None:
            self.sync.remove()
        if tape.GRAD_SINK is self:
            tape.GRAD_SINK = None

    def early(self, pairs, streams):
        if self.sync is not None:
            self.sync.early(pairs, streams)          # all-reduce (SUM) in place on self.side, ordered after `streams`
        else:
            for st in streams:
                self.side.wait_stream(st)
        with torch.cuda.stream(self.side):
            self.opt.apply_to(pairs)

    def join(self):
        torch.cuda.current_stream().wait_stream(self.side)

    def finish(self):
        if self.sync is not None:
            self.sync.finish()
        self.join()


class GraphedRLStep:
    """One self-critical RL iteration (train_rl.py:150-191: sample with grad, greedy baseline, CIDEr-D reward, criterion,
    backward, clip_gradient, Adam) replayed from CUDA graphs: rl_forward_loss + backward in one graph (with the bucketed
    gradient all-reduce captured inside when grad_sync is given), the fused clamp + Adam in a second.  Per call the new
    inputs -- features, uniforms, top-word targets and the references packed to a fixed (n_images, max_refs, max_ref_len)
    shape on the host -- are copied into static buffers; nothing is read back."""

    def __init__(self, model, crit, optimizer, fc, att, uniforms, top_true, gts, table, reward_opt, seq_per_img, reason_weight,
                 entropy_reg=0.0, temperature=1.0, max_refs=5, max_ref_len=None, warmup=2, grad_sync=None, early_optimizer=False):
        from . import reward as RW
        if not all(g.get("capturable") for g in optimizer.param_groups):
            raise RuntimeError("GraphedRLStep needs FusedAdam(capturable=True)")
        self.model, self.crit, self.opt = model, crit, optimizer
        self.reason_weight, self.entropy_reg, self.temperature = float(reason_weight), float(entropy_reg), float(temperature)
        self.table, self.ropt, self.spi = table, reward_opt, int(seq_per_img)
        self.fc = [f.clone() for f in fc]
        self.att = [a.clone() for a in att]
        self.uniforms, self.top = uniforms.clone().float(), top_true.clone()
        self.n_images = len(gts)
        self.max_refs = max(max_refs, max(len(g) for g in gts))
        self.max_ref_len = max_ref_len or min(RW.MAXLEN, model.seq_length + 2)
        dev = self.fc[0].device
        r, n = RW.pack_references_static(gts, self.n_images, self.max_refs, self.max_ref_len)
        self.refs, self.n_refs = torch.from_numpy(r).to(dev), torch.from_numpy(n).to(dev)
        self.loss = self.seq = self.greedy = self.reward = None
        from .model import _Workspace
        self._ws = _Workspace()            # the no-tape decodes' scratch: private, the graph holds raw pointers into it
        self._ws_b = _Workspace()
        saved_ws, model._wsobj = model._wsobj, self._ws
        saved_ws_b = model.__dict__.get("_wsobj_b")
        model.__dict__["_wsobj_b"] = self._ws_b
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            optimizer.init_state()
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.g_fb = torch.cuda.CUDAGraph()
            if tape.MARKS is not None:
                tape.MARKS.clear()          # keep the markers of the captured pass only
            sink = EarlyStep(optimizer, grad_sync) if (early_optimizer and tape.usable(model)) else grad_sync
            with torch.cuda.graph(self.g_fb):
                if isinstance(sink, EarlyStep):
                    optimizer.begin_step()          # the update count of this step, before its first early optimizer pass
                if sink is not None:
                    sink.install()
                self.loss, self.seq, self.greedy, self.reward = self._fwd_bwd()
                if sink is not None:
                    sink.finish()
            if sink is not None:
                sink.remove()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
                optimizer.step()
                tape.mark("opt_end")
        finally:
            model._wsobj = saved_ws
            model.__dict__["_wsobj_b"] = saved_ws_b

    def _reward(self, seq, greedy):
        from . import reward as RW
        return RW.compute_reward_packed(seq, greedy, self.refs, self.n_refs, self.table, self.ropt, self.spi)[0]

    def _fwd_bwd(self):
        self.opt.zero_grad(set_to_none=True)
        tape.mark("step_begin")
        loss, seq, greedy, reward = rl_forward_loss(self.model, self.crit, self.fc, self.att, self.uniforms, self._reward, self.top,
                                                    self.reason_weight, self.entropy_reg, self.temperature)
        loss.backward()
        tape.mark("bwd_end")
        return loss.detach(), seq, greedy, reward

    def __call__(self, fc=None, att=None, uniforms=None, top_true=None, gts=None):
        """Replays the iteration on new data (same shapes); returns the loss tensor (device, overwritten by the next call);
        .seq / .greedy / .reward hold the sampled tokens, the baseline tokens and the rewards of the replay."""
        from . import _capi, reward as RW
        if fc is not None:
            for d, s in zip(self.fc + self.att, list(fc) + list(att)):
                d.copy_(s, non_blocking=True)
        if uniforms is not None:
            self.uniforms.copy_(uniforms, non_blocking=True)
        if top_true is not None:
            self.top.copy_(top_true, non_blocking=True)
        if gts is not None:
            r, n = RW.pack_references_static(gts, self.n_images, self.max_refs, self.max_ref_len)
            self.refs.copy_(torch.from_numpy(r), non_blocking=False)
            self.n_refs.copy_(torch.from_numpy(n), non_blocking=False)
        self.g_fb.replay()
        self.g_opt.replay()
        _capi.WEIGHTS_EPOCH[0] += 1
        return self.loss


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
                 grad_sync=None, early_optimizer=False):
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
        if tape.MARKS is not None:
            tape.MARKS.clear()              # keep the markers of the captured pass only
        # grad_sync (dist.OverlappedGradSync): the data-parallel all-reduce is captured INSIDE the forward+backward graph,
        # bucket by bucket on a side stream while the backward of the earlier layers still runs (then `between` is unused)
        # early_optimizer: the stage-1 parameters (95 % of the model) are updated from inside the backward pass (EarlyStep); not
        # with `between` (an all-reduce between the two graphs must see every gradient before any update).  Measured on one
        # B200 (profiles/r2_phase_timeline_xe.log): the optimizer graph shrinks from 2.75 to 0.57 ms but stage-1 backward grows
        # by as much -- that pass is bound by GPU throughput, not by its dependent chain -- so it is off by default.
        early = early_optimizer and between is None and tape.usable(model)
        sink = EarlyStep(optimizer, grad_sync) if early else grad_sync
        with torch.cuda.graph(self.g_fb):
            if early:
                optimizer.begin_step()              # the update count of this step, before its first early optimizer pass
            if sink is not None:
                sink.install()
            self.loss = self._fwd_bwd()
            if sink is not None:
                sink.finish()
        if sink is not None:
            sink.remove()
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            optimizer.step()
            tape.mark("opt_end")

    def _fwd_bwd(self):
        self.opt.zero_grad(set_to_none=True)
        tape.mark("step_begin")
        lp, rp = forward_xe(self.model, self.fc, self.att, self.labels, col_any=self.col_any)
        loss = self.crit(lp, self.labels[:, 1:], self.masks[:, 1:], rp, self.top, self.reason_weight)
        tape.mark("loss_end")
        loss.backward()
        tape.mark("bwd_end")
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
