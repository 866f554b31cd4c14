"""GPU: the hand-scheduled training tape (tape.py: one autograd.Function per stage, side streams, T-batched decoder weight
gradients) against the op-by-op tape (autograd.py) on the same inputs -- loss, log-probs and all parameter gradients; the
op-by-op tape is the one tests/test_gpu_training.py pins to oracle autograd (and those tests now run the fused tape too)."""
from types import SimpleNamespace

import pytest
import torch

from oracle import rfnet_oracle as O
from tests._gpu_util import build_model, cuda_list, maxdiff

pytestmark = pytest.mark.gpu


def _grads(m):
    return {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}


def _close(a, b, rel):
    for k in a:
        scale = float(b[k].abs().max()) + 1e-6
        assert maxdiff(a[k], b[k]) <= rel * scale + 1e-6, f"{k}: {maxdiff(a[k], b[k]):.3g} vs scale {scale:.3g}"


@pytest.mark.parametrize("J,rows,drop", [(1, 3, 0.0), (3, 5, 0.0), (2, 4, 0.3)])
def test_fused_tape_equals_per_op_tape_xe(J, rows, drop):
    from recurrent_fusion_network_b200 import autograd as AG
    from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
    cfg = O.tiny_config(J)
    sd = O.make_state_dict(cfg, seed=70 + J, init_range=0.5, logit_scale=3.0)
    fc, att = O.make_inputs(cfg, rows, seed=J)
    labels, masks, top = O.make_labels(cfg, rows, seed=J + 5)
    crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
    m = build_model(cfg, sd, drop_prob_lm=drop).train()
    g = torch.Generator(device="cuda").manual_seed(3)
    T = labels.shape[1]
    keep = [(torch.rand(rows, cfg.rnn_size, device="cuda", generator=g) >= drop).float() for _ in range(T)]
    out = {}
    for fused in (False, True):
        m.fused_tape = fused
        m.zero_grad(set_to_none=True)
        AG.MASK_QUEUE = [k.clone() for k in keep] if drop > 0 else None
        lp, rp = m(cuda_list(fc), cuda_list(att), labels.cuda())
        AG.MASK_QUEUE = None
        loss = crit(lp, labels[:, 1:].cuda(), masks[:, 1:].cuda(), rp, top.cuda(), 10.0)
        loss.backward()
        out[fused] = (float(loss.detach()), lp.detach().contiguous().clone(), _grads(m))
    assert out[True][1].shape == out[False][1].shape
    assert maxdiff(out[True][1], out[False][1]) <= 2e-5
    assert abs(out[True][0] - out[False][0]) <= 1e-5 * max(1.0, abs(out[False][0]))
    _close(out[True][2], out[False][2], 1e-4)


def test_fused_tape_rl_forward_loss_equals_per_op():
    from recurrent_fusion_network_b200 import training as TR
    from recurrent_fusion_network_b200.criteria import ReviewNetRewardCriterion
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    rows = 6
    fc, att = O.make_inputs(cfg, rows, seed=8)
    _, _, top = O.make_labels(cfg, rows, seed=3)
    L = cfg.seq_length
    u = torch.rand(rows, L, generator=torch.Generator().manual_seed(1)).cuda()
    rw = torch.randn(rows, 1, generator=torch.Generator().manual_seed(2)).expand(rows, L).contiguous().cuda()
    crit = ReviewNetRewardCriterion(SimpleNamespace(use_label_smoothing=0, label_smoothing_epsilon=0.1))
    m = build_model(cfg, sd).train()
    out = {}
    for fused in (False, True):
        m.fused_tape = fused
        m.zero_grad(set_to_none=True)
        loss, seq, greedy, reward = TR.rl_forward_loss(m, crit, cuda_list(fc), cuda_list(att), u, lambda s, g: rw, top.cuda(), 10.0,
                                                       entropy_reg=0.01)
        loss.backward()
        out[fused] = (float(loss.detach()), seq.clone(), _grads(m))
    assert torch.equal(out[True][1], out[False][1])
    assert abs(out[True][0] - out[False][0]) <= 1e-5 * max(1.0, abs(out[False][0]))
    _close(out[True][2], out[False][2], 1e-4)


@pytest.mark.parametrize("review_maxout,decoder_maxout", [(1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("fused", [True, False], ids=["fused_tape", "per_op_tape"])
def test_maxout_gradients_match_oracle(review_maxout, decoder_maxout, fused):
    """opt.review_maxout / opt.maxout = 1 (5R-wide cells, in_transform = max of the last two R-blocks,
    misc/LSTMSoftAttentionCore.py:25,89; misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:25,60): loss and all gradients
    against autograd through the oracle, whose maxout cells are pinned to the reference by the tiny_j2_maxout fixtures."""
    import dataclasses
    from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
    cfg = dataclasses.replace(O.tiny_config(2), review_maxout=review_maxout, decoder_maxout=decoder_maxout)
    sd = O.make_state_dict(cfg, seed=91, init_range=0.5, logit_scale=3.0)
    rows = 5
    fc, att = O.make_inputs(cfg, rows, seed=4)
    labels, masks, top = O.make_labels(cfg, rows, seed=6)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lp_o, rp_o = O.forward_xe(leaves, cfg, fc, att, labels)
    want_loss = O.xe_loss(lp_o, labels[:, 1:], masks[:, 1:], rp_o, top, 10.0, 0.1)
    want_loss.backward()
    crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
    m = build_model(cfg, sd).train()
    m.fused_tape = fused
    lp, rp = m(cuda_list(fc), cuda_list(att), labels.cuda())
    loss = crit(lp, labels[:, 1:].cuda(), masks[:, 1:].cuda(), rp, top.cuda(), 10.0)
    loss.backward()
    assert maxdiff(lp, lp_o) <= 2e-5
    assert abs(float(loss.detach()) - float(want_loss.detach())) <= 1e-4 * max(1.0, abs(float(want_loss.detach())))
    for k, p in m.named_parameters():
        w = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        scale = float(w.abs().max()) + 1e-6
        assert maxdiff(g, w) <= 2e-4 * scale + 1e-6, f"{k}: {maxdiff(g, w):.3g} vs scale {scale:.3g}"
