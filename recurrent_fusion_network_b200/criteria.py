"""Criteria of the path (misc/utils.py): same class names and call signatures as the reference.

Without autograd (evaluation) the sequence terms run in the fused reduction kernels, which need only
lp[y], sum_v lp_v and sum_v p log p per (row, t) -- the reference materialises a (rows,T,V) one-hot.
With autograd enabled the same math runs through recurrent_fusion_network_b200.training."""

import torch
import torch.nn as nn

from ._capi import check, lib, ptr, stream


def _f32c(t):
    return t.float().contiguous()


def _margin_terms(top_pred, top_true, weight_each, out):
    top_true = top_true.to(torch.int64).contiguous()
    for p in top_pred:
        p = _f32c(p)
        if p.dim() == 1:
            p = p.unsqueeze(0)
        check(lib().rfn_multilabel_margin_f32(ptr(p), ptr(top_true), p.shape[0], p.shape[1], float(weight_each), 1,
                                              ptr(out), stream()), "rfn_multilabel_margin_f32")


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


class ReviewNetEnsembleCriterion(nn.Module):
    """misc/utils.py:154-192"""

    def __init__(self, opt):
        super().__init__()
        self.use_label_smoothing = opt.use_label_smoothing
        self.label_smoothing_epsilon = opt.label_smoothing_epsilon
        self.use_cuda = getattr(opt, "use_cuda", 1)

    def forward(self, log_prob, target, mask, top_pred, top_true, reason_weight):
        if _needs_grad(log_prob, *top_pred):
            from . import training
            return training.xe_criterion(self, log_prob, target, mask, top_pred, top_true, reason_weight)
        log_prob = _f32c(log_prob)
        rows, T, V = log_prob.shape
        # the reference truncates target and mask to log_prob.size(1) separately (misc/utils.py:163-164), so their widths
        # may differ: slice before making them contiguous (the kernel addresses both with one leading dimension)
        target = target[:, :T].to(torch.int64).contiguous()
        mask = _f32c(mask[:, :T])
        out = torch.zeros(1, dtype=torch.float32, device=log_prob.device)
        eps = float(self.label_smoothing_epsilon) if self.use_label_smoothing else 0.0
        check(lib().rfn_xe_loss_f32(ptr(log_prob), ptr(target), ptr(mask), target.stride(0), rows, T, V, eps, ptr(out),
                                    stream()), "rfn_xe_loss_f32")
        _margin_terms(top_pred, top_true, reason_weight / len(top_pred), out)
        return out[0]


class ReviewNetRewardCriterion(nn.Module):
    """misc/utils.py:44-84 (non-PPO branch; use_ppo defaults to 0)."""

    def __init__(self, opt):
        super().__init__()
        self.use_label_smoothing = opt.use_label_smoothing
        self.label_smoothing_epsilon = opt.label_smoothing_epsilon

    def forward(self, input, seq, reward, logprobs_all, entropy_reg, top_pred, top_true, reason_weight,
                sample_logprobs_old, opt):
        if getattr(opt, "use_ppo", 0):
            # the reference's own branch cannot execute for this model: it flattens `input` to (rows*T) but divides by the
            # un-flattened (rows, T) sample_logprobs_old (misc/utils.py:53,63-65; called from train_rl.py:162-169), which
            # raises a broadcasting error for every T != 1 -- probed in the build container with the imported reference
            raise NotImplementedError("use_ppo=1: the reference's PPO branch of ReviewNetRewardCriterion raises a shape error "
                                      "(misc/utils.py:65) and is not built; default 0 in every shipped script")
        tp = top_pred if isinstance(top_pred, list) else [top_pred]
        if _needs_grad(input, logprobs_all, *tp):
            from . import training
            return training.rl_criterion(self, input, seq, reward, logprobs_all, entropy_reg, tp, top_true, reason_weight)
        input, reward, logprobs_all = _f32c(input), _f32c(reward), _f32c(logprobs_all)
        seq = seq.to(torch.int64).contiguous()
        rows, T = input.shape
        V = logprobs_all.shape[2]
        out = torch.zeros(1, dtype=torch.float32, device=input.device)
        check(lib().rfn_rl_loss_f32(ptr(input), ptr(seq), ptr(reward), ptr(logprobs_all), logprobs_all.stride(0), rows, T,
                                    V, float(entropy_reg), ptr(out), stream()), "rfn_rl_loss_f32")
        _margin_terms(tp, top_true, reason_weight / len(tp), out)
        return out[0]


def clip_gradient(optimizer, grad_clip):
    """Element-wise clamp of every gradient (misc/utils.py:292-296)."""
    for group in optimizer.param_groups:
        for param in group["params"]:
            if param.grad is not None:
                param.grad.data.clamp_(-grad_clip, grad_clip)


def set_lr(optimizer, lr):
    for group in optimizer.param_groups:
        group["lr"] = lr
