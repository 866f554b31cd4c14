"""Shared helpers of the GPU parity tests: build the CUDA model from an oracle config + state_dict."""
import torch

from oracle import rfnet_oracle as O
from recurrent_fusion_network_b200 import RecurrentFusionModel, make_opt

LP_TOL = 1e-4  # north star: per-step log-probs within 1e-4 abs in fp32 mode


def opt_from_cfg(cfg: O.RFNConfig, **kw):
    return make_opt(
        feat_array_info=[dict(fc_feat_size=e.fc_feat_size, att_feat_size=e.att_feat_size, att_num=e.att_num)
                         for e in cfg.encoders],
        vocab_size=cfg.vocab_size, input_encoding_size=cfg.input_encoding_size, rnn_size=cfg.rnn_size,
        seq_length=cfg.seq_length, num_review_steps=cfg.num_review_steps,
        num_review_steps_0=cfg.num_review_steps_0, top_words_count=cfg.top_words_count,
        att_hid_size=cfg.att_hid_size, review_maxout=cfg.review_maxout, maxout=cfg.decoder_maxout, **kw)


def build_model(cfg, sd, **kw):
    m = RecurrentFusionModel(opt_from_cfg(cfg, **kw))
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    return m.cuda().eval()


def cuda_list(ts):
    return [t.cuda() for t in ts]


def maxdiff(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def assert_tokens_match_with_tie_policy(got, want, oracle_lp_all, what, margin=1e-5):
    """Exact token match; a first divergence where the ORACLE's top-2 margin is < `margin` is a
    tie-break, not a failure.  Returns the number of tie-broken rows."""
    got = got.cpu()
    ties = 0
    for r in range(want.shape[0]):
        if torch.equal(got[r], want[r]):
            continue
        t = int((got[r] != want[r]).nonzero()[0])
        m = float(O.top2_margin(oracle_lp_all[r, t]))
        assert m < margin, f"{what}: row {r} diverges at t={t} with oracle margin {m:.3g}"
        ties += 1
    return ties


def assert_beam_match_with_tie_policy(got_seq, got_lp, want_seq, want_lp, margins, what, margin=1e-5, lp_tol=LP_TOL):
    """Beam captions: exact token match per image; an image whose caption differs is excused only when the
    ORACLE's own search had a decision (a merge step or the final ranking of the finished beams) whose margin is
    < `margin` (oracle.sample_beam(margins_out=...)).  Log-probs of equal captions must agree
    to `lp_tol`.  Returns the number of tie-broken images."""
    got_seq, got_lp = got_seq.cpu(), got_lp.cpu()
    ties = 0
    for k in range(want_seq.shape[0]):
        if torch.equal(got_seq[k], want_seq[k]):
            d = float((got_lp[k] - want_lp[k]).abs().max())
            assert d <= lp_tol, f"{what}: image {k} log-probs differ by {d:.3g}"
            continue
        m = min(min(margins[k]["steps"], default=float("inf")), margins[k]["final"])
        assert m < margin, f"{what}: image {k} caption differs and the oracle's smallest decision margin is {m:.3g}"
        ties += 1
    return ties
