"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and against the
golden fixtures written from the REAL reference.  Tolerances are the north star's: tokens exact
(near-tie policy of SURVEY.md 4.3), per-step log-probs <= 1e-4 abs (fp32 mode)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import rfnet_oracle as O
from tests import _golden as G
from tests._gpu_util import (LP_TOL, assert_tokens_match_with_tie_policy, build_model, cuda_list, maxdiff)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device_ok():
    from recurrent_fusion_network_b200 import _capi
    _capi.check(_capi.lib().rfn_check_device(), "rfn_check_device")
    torch.set_num_threads(16)


# ---- operator level ---------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,Ks", [(1, 8, [4]), (3, 20, [24, 40]), (16, 2048, [2560, 2208]), (80, 512, [512]),
                                    (200, 9488, [512]), (700, 130, [36, 64, 128]), (33, 33, [12])])
def test_linear_matches_torch(M, N, Ks):
    from recurrent_fusion_network_b200.model import linear
    g = torch.Generator().manual_seed(M * 131 + N)
    xs = [torch.randn(M, k, generator=g) for k in Ks]
    lins = [torch.nn.Linear(k, N) for k in Ks]
    want = sum(l(x).double() for l, x in zip(lins, xs))
    want = sum(torch.nn.functional.linear(x.double(), l.weight.double(), l.bias.double()) for l, x in zip(lins, xs))
    got = linear([(x.cuda(), l.cuda()) for x, l in zip(xs, lins)], M, N)
    scale = float(want.abs().max()) + 1.0
    assert maxdiff(got, want) <= 2e-5 * scale


@pytest.mark.parametrize("rows,N,D,R,Ah", [(2, 5, 16, 32, 16), (7, 196, 2048, 512, 512), (3, 49, 2208, 512, 512),
                                           (300, 8, 512, 512, 512)])
def test_attention_core_matches_oracle(rows, N, D, R, Ah):
    from recurrent_fusion_network_b200 import AttentionModelCore
    torch.manual_seed(rows + N)
    mod = AttentionModelCore(R, D, N, Ah)
    sd = {"m." + k: v for k, v in mod.state_dict().items()}
    h = torch.randn(rows, R)
    A = torch.randn(rows, N, D)
    want = O.attention(sd, "m", h, A)
    with torch.no_grad():
        got = mod.cuda()(h.cuda(), A.cuda())
    assert maxdiff(got, want) <= 2e-5


def test_core_modules_match_oracle_steps():
    """FeatArrayFusionNoInputCore / LSTMSoftMultiAttention... / LSTMSoftAttentionCore module mirrors."""
    cfg = O.tiny_config(3)
    sd = O.make_state_dict(cfg, seed=5, init_range=0.5)
    m = build_model(cfg, sd)
    fc, att = O.make_inputs(cfg, 6, seed=3)
    with torch.no_grad():
        st_o = O.get_init_state(sd, cfg, fc)
        st_g = m.get_init_state(cuda_list(fc))
        for (ho, co), (hg, cg) in zip(st_o, st_g):
            assert hg.shape == (1, 6, cfg.rnn_size)
            assert maxdiff(hg[0], ho) <= 1e-5 and maxdiff(cg[0], co) <= 1e-5
        # stage-1 step
        new_o = O.stage1_step(sd, cfg, 0, att, st_o)
        out_g, new_g = m.review_steps_individual[0](cuda_list(att), list(st_g))
        for j in range(cfg.J):
            assert maxdiff(out_g[j], new_o[j][0]) <= 2e-5
            assert maxdiff(new_g[j][1][0], new_o[j][1]) <= 2e-5
        # stage-2 step on random thought vectors
        TV = [torch.randn(6, cfg.num_review_steps_0, cfg.rnn_size) for _ in range(cfg.J)]
        h = torch.randn(6, cfg.rnn_size); c = torch.randn(6, cfg.rnn_size)
        ho, co = O.stage2_step(sd, cfg, 1, TV, (h, c))
        og, (hg, cg) = m.review_steps[1](cuda_list(TV), (h.cuda().unsqueeze(0), c.cuda().unsqueeze(0)))
        assert maxdiff(og, ho) <= 2e-5 and maxdiff(cg[0], co) <= 2e-5
        # decoder step
        TVc = torch.randn(6, cfg.num_review_steps, cfg.rnn_size)
        x = torch.randn(6, cfg.input_encoding_size)
        ho, co = O.decoder_step(sd, x, TVc, (h, c))
        og, (hg, cg) = m.decoder(x.cuda(), TVc.cuda(), (h.cuda().unsqueeze(0), c.cuda().unsqueeze(0)))
        assert maxdiff(og, ho) <= 2e-5 and maxdiff(cg[0], co) <= 2e-5
        # one_time_step returns logits
        lo, _ = O.one_time_step(sd, x, TVc, (h, c))
        lg, _ = m.one_time_step(x.cuda(), None, TVc.cuda(), (h.cuda().unsqueeze(0), c.cuda().unsqueeze(0)))
        assert maxdiff(lg, lo) <= 5e-5
        # get_thought_vectors (takes the state_list from get_init_state)
        TVc_o, rp_o, st2_o = O.get_thought_vectors(sd, cfg, att, st_o)
        TVc_g, rp_g, st2_g = m.get_thought_vectors(cuda_list(fc), cuda_list(att), m.get_init_state(cuda_list(fc)))
        assert maxdiff(TVc_g, TVc_o) <= 2e-5
        assert maxdiff(st2_g[0][0], st2_o[0]) <= 2e-5 and maxdiff(st2_g[1][0], st2_o[1]) <= 2e-5
        for a, b in zip(rp_g, rp_o):
            assert maxdiff(a, b) <= 5e-5


def test_review_net_core_alias():
    """LSTMSoftAttentionNoInputCore == J=1 stage-1 math with h2h(pre_h)."""
    from recurrent_fusion_network_b200 import LSTMSoftAttentionNoInputCore
    torch.manual_seed(0)
    mod = LSTMSoftAttentionNoInputCore(32, 24, 7, 16, 0.0)
    sd = {"m." + k: v.clone() for k, v in mod.state_dict().items()}
    h = torch.randn(4, 32); c = torch.randn(4, 32); A = torch.randn(4, 7, 24)
    z = O.attention(sd, "m", h, A)
    Gm = torch.nn.functional.linear(h, sd["m.h2h.weight"], sd["m.h2h.bias"]) + \
        torch.nn.functional.linear(z, sd["m.z2h.weight"], sd["m.z2h.bias"])
    ho, co = O.lstm_cell(Gm, c)
    with torch.no_grad():
        og, (hg, cg) = mod.cuda()(A.cuda(), None, None, (h.cuda().unsqueeze(0), c.cuda().unsqueeze(0)))
    assert maxdiff(og, ho) <= 2e-5 and maxdiff(cg[0], co) <= 2e-5


def test_review_net_core_matches_reference_module_fixture():
    """SURVEY 8(a) row a4 against the REFERENCE class itself: tests/golden/review_core_cases.npz holds the state_dict, inputs and
    outputs of misc/LSTMSoftAttentionNoInputCore.py (oracle/gen_golden_cores.py: three chained steps, plain and maxout)."""
    import os
    import numpy as np
    from recurrent_fusion_network_b200 import LSTMSoftAttentionNoInputCore
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "review_core_cases.npz"))
    for name in fx["names"]:
        R, D, N, A, maxout, rows = (int(v) for v in fx[f"{name}.dims"])
        mod = LSTMSoftAttentionNoInputCore(R, D, N, A, 0.0, maxout)
        mod.load_state_dict({k[len(name) + 4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith(f"{name}.sd.")})
        mod = mod.cuda().eval()
        att = torch.from_numpy(fx[f"{name}.att"]).cuda()
        state = (torch.from_numpy(fx[f"{name}.h0"]).cuda(), torch.from_numpy(fx[f"{name}.c0"]).cuda())
        with torch.no_grad():
            for step in range(fx[f"{name}.h"].shape[0]):
                out, state = mod(att, None, None, state)
                assert maxdiff(out, fx[f"{name}.h"][step]) <= 2e-5 and maxdiff(state[0][0], fx[f"{name}.h"][step]) <= 2e-5
                assert maxdiff(state[1][0], fx[f"{name}.c"][step]) <= 2e-5


# ---- path level vs the REAL reference's outputs (golden fixtures) -----------------------------------
@pytest.fixture(params=[4, 1, 3, 0], ids=["tc_fp16x3", "tc3xtf32", "tc_tf32_bf16x", "simt"])
def gemm_mode(request):
    """Run under the fp32-grade engines: tcgen05 split-fp16 (mode 4, the default), 3xTF32, TF32 + BF16 cross terms (mode 3)
    and fp32 SIMT."""
    from recurrent_fusion_network_b200 import _capi
    prev = _capi.lib().rfn_get_gemm_mode()
    _capi.check(_capi.lib().rfn_set_gemm_mode(request.param))
    yield request.param
    _capi.check(_capi.lib().rfn_set_gemm_mode(prev))


@pytest.mark.parametrize("name", G.TINY + G.FULL)
def test_path_matches_reference_fixture(name, gemm_mode):
    cfg, sd, fc, att, labels, masks, top_words, d = G.load_case(name)
    stride = int(d["stride"])
    m = build_model(cfg, sd)
    fcg, attg = cuda_list(fc), cuda_list(att)
    with torch.no_grad():
        # forward (teacher forced)
        lp, rp = m(fcg, attg, labels.cuda())
        assert lp.shape[1] == int(d["xe_T"])
        assert maxdiff(lp[:, :, ::stride], d["xe_lp_strided"]) <= LP_TOL
        tgt = lp.gather(2, labels[:, 1:lp.shape[1] + 1].cuda().unsqueeze(2)).squeeze(2)
        assert maxdiff(tgt, d["xe_lp_target"]) <= LP_TOL
        rps = torch.stack([r.reshape(len(fc[0]), -1)[:, ::max(1, stride // 8)] for r in rp])
        assert maxdiff(rps, d["reason_pred_strided"]) <= LP_TOL
        # greedy sample
        s, sl, la, _ = m.sample(fcg, attg, {"sample_max": 1, "beam_size": 1})
        want = torch.from_numpy(d["greedy_seq"])
        assert s.shape == want.shape
        if not torch.equal(s.cpu(), want):
            _, _, la_o, _ = O.sample(sd, cfg, fc, att)
            assert_tokens_match_with_tie_policy(s, want, la_o, name + " greedy")
        else:
            assert maxdiff(sl, d["greedy_slp"]) <= LP_TOL
            assert maxdiff(la[:, :, ::stride], d["greedy_lp_all_strided"]) <= LP_TOL
        # beam search
        beam = int(d["beam"])
        bs, bl, ts, tp, rpb = m.sample_beam(fcg, attg, {"beam_size": beam})
        assert np.array_equal(bs.cpu().numpy(), d["beam_seq"]), name + " beam tokens"
        assert maxdiff(bl, d["beam_lp"]) <= LP_TOL
        gs, gp = G.top_lists(d)
        assert [tuple(t.shape) for t in ts] == [tuple(t.shape) for t in gs]
        for a, b, pa, pb in zip(ts, gs, tp, gp):
            assert torch.equal(a, b)
            assert maxdiff(torch.tensor(pa), torch.tensor(pb)) <= 2 * LP_TOL
        assert len(m.done_beams) == len(fc[0]) and len(m.done_beams[0]) == int(d["beam_n_done"][0])
        assert len(rpb) == len(fc[0]) and rpb[0][0].shape == (beam, cfg.top_words_count)


@pytest.mark.parametrize("name", ["tiny_j2_eos_b", "tiny_j3_eos_a", "config1_n49"])
def test_multinomial_sample_with_shared_uniforms(name):
    cfg, sd, fc, att, *_ = G.load_case(name)
    rows = fc[0].shape[0]
    g = torch.Generator().manual_seed(99)
    u = torch.rand(rows, cfg.seq_length, generator=g)
    m = build_model(cfg, sd)
    with torch.no_grad():
        for temp in (1.0, 0.7):
            s, sl, la, _ = m.sample(cuda_list(fc), cuda_list(att), {"sample_max": 0, "temperature": temp, "uniforms": u.cuda()})
            so, slo, lao, _ = O.sample(sd, cfg, fc, att, sample_max=0, temperature=temp, uniforms=u)
            if s.shape == so.shape and torch.equal(s.cpu(), so):
                assert maxdiff(sl, slo) <= LP_TOL and maxdiff(la, lao) <= LP_TOL
            else:
                # a uniform landed within float rounding of a CDF boundary: pin log-probs given OUR tokens
                sf, slf, laf, _ = O.sample(sd, cfg, fc, att, forced_tokens=torch.nn.functional.pad(
                    s.cpu(), (0, cfg.seq_length - s.shape[1])))
                T = min(s.shape[1], sf.shape[1])
                assert maxdiff(la[:, :T + 1], laf[:, :T + 1]) <= LP_TOL


def test_multinomial_distribution():
    """Sampling frequencies follow exp(lp) (distributional parity)."""
    cfg = O.tiny_config(1)
    sd = O.make_state_dict(cfg, seed=3, init_range=0.5, logit_scale=3.0)
    fc, att = O.make_inputs(cfg, 1, seed=1)
    rows = 4096
    fcr = [f.expand(rows, -1).contiguous().cuda() for f in fc]
    attr = [a.expand(rows, -1, -1).contiguous().cuda() for a in att]
    m = build_model(cfg, sd)
    with torch.no_grad():
        s, sl, la, _ = m.sample(fcr, attr, {"sample_max": 0})
    p = la[0, 0].exp().double().cpu()
    freq = torch.bincount(s[:, 0].cpu(), minlength=cfg.V1).double() / rows
    assert float((freq - p).abs().max()) < 0.04


def test_beam_batching_is_chunk_invariant():
    """Batched device beam == the same images decoded in smaller chunks (and == oracle per image)."""
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, 37, seed=21)
    m = build_model(cfg, sd)
    with torch.no_grad():
        a = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        m.chunk_images = 8
        b = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        m.chunk_images = 1024
        o = O.sample_beam(sd, cfg, fc, att, beam_size=3)
    assert torch.equal(a[0], b[0]) and maxdiff(a[1], b[1]) <= 2e-5   # chunk size selects different GEMM kernels
    assert torch.equal(a[0].cpu(), o[0]) and maxdiff(a[1], o[1]) <= LP_TOL
    assert [t.shape for t in a[2]] == [t.shape for t in o[2]]
    for x, y in zip(a[2], o[2]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("beam", [1, 2, 5, 10])
def test_other_beam_widths(beam):
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1247, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, 5, seed=8)
    m = build_model(cfg, sd)
    with torch.no_grad():
        if beam == 1:
            a = m._beam_tensors(cuda_list(fc), cuda_list(att), 5, 1)
            seq, slp = a[0], a[1]
        else:
            seq, slp, *_ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": beam})
        o = O.sample_beam(sd, cfg, fc, att, beam_size=beam)
    assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= LP_TOL


def test_ensemble_beam_matches_oracle():
    from recurrent_fusion_network_b200.ensemble import ensemble_sample_beam
    cfg = O.tiny_config(2)
    sds = [O.make_state_dict(cfg, seed=1250 + i, init_range=0.5, logit_scale=3.0, eos_bias=0.8) for i in range(3)]
    fc, att = O.make_inputs(cfg, 6, seed=8)
    models = [build_model(cfg, sd) for sd in sds]
    with torch.no_grad():
        seq, slp, ts, tp = ensemble_sample_beam(models, cuda_list(fc), cuda_list(att), {"beam_size": 3})
        o = O.ensemble_sample_beam(sds, cfg, fc, att, beam_size=3)
    assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= LP_TOL
    assert [t.shape for t in ts] == [t.shape for t in o[2]]


def test_fused_criteria_match_oracle():
    from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion, ReviewNetRewardCriterion
    from types import SimpleNamespace
    cfg, sd, fc, att, labels, masks, top_words, d = G.load_case("tiny_j2_eos_b")
    lp, rp = O.forward_xe(sd, cfg, fc, att, labels)
    for ls in (1, 0):
        crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=ls, label_smoothing_epsilon=0.1, use_cuda=1))
        with torch.no_grad():
            got = crit(lp.cuda(), labels[:, 1:].cuda(), masks[:, 1:].cuda(), cuda_list(rp), top_words.cuda(), 10.0)
        want = d["xe_loss_ls"] if ls else d["xe_loss_nols"]
        assert abs(float(got) - float(want)) <= 1e-4 * max(1.0, abs(float(want)))
    s, sl, la, _ = O.sample(sd, cfg, fc, att)
    g = torch.Generator().manual_seed(int(d["rl_reward_seed"]))
    reward = torch.randn(s.shape[0], 1, generator=g).expand(s.shape[0], s.shape[1]).contiguous()
    rl = ReviewNetRewardCriterion(SimpleNamespace(use_label_smoothing=0, label_smoothing_epsilon=0.1))
    with torch.no_grad():
        got = rl(sl.cuda(), s.cuda(), reward.cuda(), la.cuda(), 0.01, cuda_list(rp), top_words.cuda(), 10.0, None,
                 SimpleNamespace(use_ppo=0))
    assert abs(float(got) - float(d["rl_loss"])) <= 1e-4 * max(1.0, abs(float(d["rl_loss"])))


# ---- full-size, size-independent properties ---------------------------------------------------------
def test_full_size_properties():
    """At BASELINE sizes (five encoders, 9488-way vocab): replicated rows decode identically,
    batched beam == per-image beam, log-probs normalise, and a permutation of the images permutes
    the captions."""
    cfg = O.RFNConfig()
    sd = O.make_state_dict(cfg, seed=1234, sharpen=True)
    fc, att = O.make_inputs(cfg, 24, seed=9)
    m = build_model(cfg, sd)
    fcg, attg = cuda_list(fc), cuda_list(att)
    with torch.no_grad():
        seq, slp, *_ = m.sample_beam(fcg, attg, {"beam_size": 3})
        perm = torch.randperm(24, generator=torch.Generator().manual_seed(1))
        seq_p, slp_p, *_ = m.sample_beam([f[perm.cuda()] for f in fcg], [a[perm.cuda()] for a in attg], {"beam_size": 3})
        assert torch.equal(seq_p.cpu(), seq.cpu()[perm]) and maxdiff(slp_p.cpu(), slp.cpu()[perm]) <= 1e-5
        m.chunk_images = 7
        seq_c, slp_c, *_ = m.sample_beam(fcg, attg, {"beam_size": 3})
        m.chunk_images = 1024
        assert torch.equal(seq_c, seq) and maxdiff(slp_c, slp) <= 5e-5   # different chunking -> different GEMM engines/tiles
        s, sl, la, _ = m.sample(fcg, attg, {"sample_max": 1})
        assert maxdiff(la.exp().sum(-1), torch.ones(la.shape[:2])) <= 1e-4
        # finished rows stay zero, and seqLogprobs is the log-prob of the chosen token
        z = (s == 0)
        assert bool((z[:, 1:] | ~z[:, :-1]).all())   # once a row emits 0 it stays 0
        first = O.sample(sd, cfg, [f[:2] for f in fc], [a[:2] for a in att])
        assert_tokens_match_with_tie_policy(s[:2, :first[0].shape[1]], first[0], first[2], "full greedy")


# ---- edge cases ----------------------------------------------------------------------------------------
def test_single_row_and_ragged_shapes():
    """rows == 1 (the reference's .squeeze() collapses (1,K) reason_pred to (K,)), a batch that is not a
    multiple of anything, and N_j == 1."""
    cfg = O.RFNConfig(encoders=(O.Encoder(1, 16, 24), O.Encoder(7, 24, 16)), rnn_size=32, att_hid_size=16,
                      input_encoding_size=24, vocab_size=59, seq_length=5, num_review_steps_0=2, num_review_steps=3,
                      top_words_count=20)
    sd = O.make_state_dict(cfg, seed=11, init_range=0.5, logit_scale=3.0, eos_bias=0.5)
    m = build_model(cfg, sd)
    for rows in (1, 13):
        fc, att = O.make_inputs(cfg, rows, seed=rows)
        labels, masks, top = O.make_labels(cfg, rows, seed=rows + 1)
        with torch.no_grad():
            lp, rp = m(cuda_list(fc), cuda_list(att), labels.cuda())
            lp_o, rp_o = O.forward_xe(sd, cfg, fc, att, labels)
            assert maxdiff(lp, lp_o) <= LP_TOL
            assert rp[0].shape == ((cfg.top_words_count,) if rows == 1 else (rows, cfg.top_words_count))
            bs, bl, ts, tp, _ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
            o = O.sample_beam(sd, cfg, fc, att, beam_size=3)
            assert torch.equal(bs.cpu(), o[0]) and maxdiff(bl, o[1]) <= LP_TOL
            try:
                s, sl, la, _ = m.sample(cuda_list(fc), cuda_list(att), {"sample_max": 1})
                so, slo, lao, _ = O.sample(sd, cfg, fc, att)
                assert torch.equal(s.cpu(), so) and maxdiff(la, lao) <= LP_TOL
            except RuntimeError:
                with pytest.raises(RuntimeError):   # all rows emitted <eos> at t=1: the reference fails too
                    O.sample(sd, cfg, fc, att)


def test_max_beam_width_and_bad_arguments():
    from recurrent_fusion_network_b200 import _capi
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1247, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, 3, seed=8)
    m = build_model(cfg, sd)
    with torch.no_grad():
        for beam in (8, 16):
            seq, slp, ts, tp, _ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": beam})
            o = O.sample_beam(sd, cfg, fc, att, beam_size=beam)
            assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= LP_TOL
            assert [t.shape for t in ts] == [t.shape for t in o[2]]
        # the reference's DEFAULT beam width (sample_beam's opt.get('beam_size', 10), :353)
        seq, slp, ts, tp, _ = m.sample_beam(cuda_list(fc), cuda_list(att), {})
        o = O.sample_beam(sd, cfg, fc, att, beam_size=10)
        assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= LP_TOL
        with pytest.raises(_capi.RfnError):
            m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 17})
        with pytest.raises(_capi.RfnError):
            m.sample(cuda_list(fc)[:1], cuda_list(att), {})                      # wrong encoder count
        with pytest.raises(_capi.RfnError):
            m.sample(cuda_list(fc), [a[:, :-1] for a in cuda_list(att)], {})     # wrong att_num


def test_full_size_ensemble_two_models():
    """BASELINE.json configs[4] at reference sizes (two five-encoder models, logit-mean ensemble, beam 3)."""
    from recurrent_fusion_network_b200.ensemble import ensemble_sample_beam, model_ensemble_feat_array_one_step
    cfg = O.RFNConfig()
    sds = [O.make_state_dict(cfg, seed=1235 + i, sharpen=True) for i in range(2)]
    fc, att = O.make_inputs(cfg, 3, seed=12)
    models = [build_model(cfg, sd) for sd in sds]
    torch.set_num_threads(16)
    with torch.no_grad():
        seq, slp, ts, tp = ensemble_sample_beam(models, cuda_list(fc), cuda_list(att), {"beam_size": 3})
        o = O.ensemble_sample_beam(sds, cfg, fc, att, beam_size=3)
        assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= LP_TOL
        assert [t.shape for t in ts] == [t.shape for t in o[2]]
        # one ensemble step through the model-level API (eval_utils.py:268-290)
        states, tvs, xts = [], [], []
        for m in models:
            TVc, _, st = m.get_thought_vectors(cuda_list(fc), cuda_list(att), m.get_init_state(cuda_list(fc)))
            tvs.append(TVc); states.append(st)
            xts.append(m.embed.weight[torch.zeros(3, dtype=torch.int64, device="cuda")])
        _, _, lp = model_ensemble_feat_array_one_step(models, xts, states, tvs)
        lg = []
        for sd in sds:
            TVc_o, _, st_o = O.get_thought_vectors(sd, cfg, att, O.get_init_state(sd, cfg, fc))
            l, _ = O.one_time_step(sd, sd["embed.weight"][torch.zeros(3, dtype=torch.int64)], TVc_o, st_o)
            lg.append(l)
        want = torch.log_softmax(sum(lg) / 2, dim=1)
        assert maxdiff(lp, want) <= LP_TOL
