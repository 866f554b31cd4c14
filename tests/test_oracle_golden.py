"""CPU: the oracle restatement reproduces the REAL reference's outputs stored in tests/golden/
(written by oracle/gen_golden.py from /root/reference).  Tokens exact; floats <= 2e-5 abs
(reference fp32-vs-fp64 drift is ~1.2e-6; 2e-5 leaves room for BLAS thread counts)."""
import numpy as np
import pytest
import torch

from oracle import rfnet_oracle as O
from tests import _golden as G

TOL = 2e-5


def _close(a, b, tol=TOL):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert float((a - b).abs().max()) <= tol, float((a - b).abs().max())


@pytest.mark.parametrize("name", G.TINY + G.FULL)
def test_oracle_matches_reference_fixture(name):
    torch.set_num_threads(8)
    cfg, sd, fc, att, labels, masks, top_words, d = G.load_case(name)
    stride = int(d["stride"])
    tol = TOL * (10 if "sharp" in name else 1)
    with torch.no_grad():
        lp, rp = O.forward_xe(sd, cfg, fc, att, labels)
        assert lp.shape[1] == int(d["xe_T"])
        _close(lp[:, :, ::stride], d["xe_lp_strided"], tol)
        _close(torch.stack([r[:, ::max(1, stride // 8)] for r in rp]), d["reason_pred_strided"], tol)
        s, sl, la, _ = O.sample(sd, cfg, fc, att, sample_max=1)
        assert np.array_equal(s.numpy(), d["greedy_seq"])
        _close(sl, d["greedy_slp"], tol)
        _close(la[:, :, ::stride], d["greedy_lp_all_strided"], tol)
        bs, bl, ts, tp, _ = O.sample_beam(sd, cfg, fc, att, beam_size=int(d["beam"]))
        assert np.array_equal(bs.numpy(), d["beam_seq"])
        _close(bl, d["beam_lp"], tol)
        gs, gp = G.top_lists(d)
        assert [t.shape for t in ts] == [t.shape for t in gs]
        for a, b, pa, pb in zip(ts, gs, tp, gp):
            assert torch.equal(a, b)
            _close(pa, pb, tol)
        # criteria
        l1 = O.xe_loss(lp, labels[:, 1:], masks[:, 1:], rp, top_words, 10.0, 0.1)
        l0 = O.xe_loss(lp, labels[:, 1:], masks[:, 1:], rp, top_words, 10.0, 0.0)
        assert abs(float(l1) - float(d["xe_loss_ls"])) <= 1e-4 * max(1.0, abs(float(l1)))
        assert abs(float(l0) - float(d["xe_loss_nols"])) <= 1e-4 * max(1.0, abs(float(l0)))
        g = torch.Generator().manual_seed(int(d["rl_reward_seed"]))
        reward = torch.randn(s.shape[0], 1, generator=g).expand(s.shape[0], s.shape[1]).contiguous()
        r = O.rl_loss(sl, s, reward, la, 0.01, rp, top_words, 10.0)
        assert abs(float(r) - float(d["rl_loss"])) <= 1e-4 * max(1.0, abs(float(r)))


def test_state_dict_has_773_tensors_for_full_model():
    shapes = O.state_dict_shapes(O.RFNConfig())
    assert len(shapes) == 773
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 489_465_041


def test_inverse_cdf_sampler_distribution():
    g = torch.Generator().manual_seed(3)
    p = torch.tensor([[0.1, 0.0, 0.6, 0.3]]).expand(20000, 4).contiguous()
    u = torch.rand(20000, generator=g)
    tok = O.inverse_cdf_sample(p, u)
    freq = torch.bincount(tok, minlength=4).double() / 20000
    assert freq[1] == 0
    assert float((freq - p[0].double()).abs().max()) < 0.02


def test_beam_merge_edge_cases():
    # t == 1: only beam 0 is live; ties keep (c, q) insertion order
    L, b = 4, 3
    seq = torch.zeros(L, b, dtype=torch.int64)
    lp = torch.zeros(L, b)
    sm = torch.zeros(b)
    done = []
    ys = torch.tensor([[-1.0, -1.0, -2.0]] * 3)
    ix = torch.tensor([[5, 0, 7]] * 3)
    src = O.beam_merge(b, 1, L, ys, ix, seq, lp, sm, done)
    assert src == [0, 0, 0]
    assert seq[0].tolist() == [5, 0, 7]
    assert len(done) == 1 and done[0]["seq"].tolist() == [0, 0, 0, 0]
    # t == 2: beam 1 ended (token 0) and is skipped, its slot is refilled from live beams
    src = O.beam_merge(b, 2, L, ys, ix, seq, lp, sm, done)
    assert 1 not in src
    # all beams ended -> None
    seq[1] = 0
    assert O.beam_merge(b, 3, L, ys, ix, seq, lp, sm, done) is None


def test_eval_split_oracle_reproduces_reference_driver_fixture():
    """oracle/eval_oracle.eval_split against the return values of the REFERENCE's own eval_split (eval_utils.py:66-265), run from
    its source text by oracle/gen_golden_eval.py on the reference model: prediction lists (ids, captions, what gets popped, both
    break conditions, the val_images_use = -1 quirk) equal, mean loss within 2e-6."""
    import json
    import os
    from oracle import eval_oracle as EO
    from tests.test_eval_utils import FakeLoader
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_split_cases.json")))
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=fx["weights_seed"], init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    for c in fx["cases"]:
        kw = {"eval_split": "val", "val_images_use": c["val_images_use"], "beam_size": c["beam_size"], "reason_weight": 10}
        loss, preds = EO.eval_split(sd, cfg, FakeLoader(cfg, c["n_images"], c["batch"], fx["seq_per_img"], seed=fx["loader_seed"]), kw)
        assert preds == c["predictions"], c
        assert abs(loss - c["loss"]) <= 2e-6 * max(1.0, abs(c["loss"])), (loss, c["loss"])


def test_ensemble_step_oracle_reproduces_reference_hook_fixture():
    """The ensemble step (per-model one_time_step, logit mean, log_softmax) against the output of the reference's own hook
    model_ensemble_feat_array_one_step (eval_utils.py:268-290), run from its source text on three reference models by
    oracle/gen_golden_eval.py: two consecutive greedy steps, log-probs within 2e-6 (measured difference: 0)."""
    import os
    import numpy as np
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "ensemble_step_case.npz"))
    cfg = O.tiny_config(2)
    sds = [O.make_state_dict(cfg, seed=int(s), init_range=0.5, logit_scale=3.0, eos_bias=0.8) for s in fx["seeds"]]
    rows = int(fx["rows"])
    fc, att = O.make_inputs(cfg, rows, seed=int(fx["input_seed"]))
    with torch.no_grad():
        tvs, sts = [], []
        for sd in sds:
            tv, _, st = O.get_thought_vectors(sd, cfg, att, O.get_init_state(sd, cfg, fc))
            tvs.append(tv); sts.append(st)
        tok = torch.zeros(rows, dtype=torch.int64)
        for step in range(fx["logprobs"].shape[0]):
            logits = []
            for k, sd in enumerate(sds):
                lg, sts[k] = O.one_time_step(sd, sd["embed.weight"][tok], tvs[k], sts[k])
                logits.append(lg)
            lp = torch.log_softmax(sum(logits) / len(sds), dim=1)
            assert float((lp - torch.from_numpy(fx["logprobs"][step])).abs().max()) <= 2e-6
            tok = lp.argmax(1)


def test_review_core_oracle_math_reproduces_reference_module_fixture():
    """misc/LSTMSoftAttentionNoInputCore.py run by oracle/gen_golden_cores.py (the reference class itself, three
    chained steps, plain and maxout).  The oracle's attention + cell composition with h2h(pre_h) must reproduce its outputs."""
    import os
    import numpy as np
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "review_core_cases.npz"))
    for name in fx["names"]:
        R, D, N, A, maxout, rows = (int(v) for v in fx[f"{name}.dims"])
        sd = {"m." + k[len(name) + 4:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith(f"{name}.sd.")}
        att = torch.from_numpy(fx[f"{name}.att"])
        h, c = torch.from_numpy(fx[f"{name}.h0"])[0], torch.from_numpy(fx[f"{name}.c0"])[0]
        for step in range(fx[f"{name}.h"].shape[0]):
            z = O.attention(sd, "m", h, att)
            G = torch.nn.functional.linear(h, sd["m.h2h.weight"], sd["m.h2h.bias"]) + \
                torch.nn.functional.linear(z, sd["m.z2h.weight"], sd["m.z2h.bias"])
            h, c = O.lstm_cell(G, c)          # (5R-wide G selects the maxout form)
            assert float((h - torch.from_numpy(fx[f"{name}.h"][step])).abs().max()) <= 2e-6
            assert float((c - torch.from_numpy(fx[f"{name}.c"][step])).abs().max()) <= 2e-6
