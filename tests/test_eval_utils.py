"""eval_utils drivers (eval_utils.py:66-265, 387-719) over a loader double with the reference's protocol."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import rfnet_oracle as O


class FakeLoader:
    """Protocol of dataloader.DataLoader used by the drivers: batches of `batch_size` images, each replicated seq_per_img
    times, bounds with it_pos_now / it_max / wrapped."""

    def __init__(self, cfg, n_images, batch_size, seq_per_img, seed=0):
        self.cfg, self.n, self.batch_size, self.seq_per_img = cfg, n_images, batch_size, seq_per_img
        self.fc, self.att = O.make_inputs(cfg, n_images, seed=seed)
        self.labels, self.masks, self.top = O.make_labels(cfg, n_images * seq_per_img, seed=seed + 1)
        self.vocab = {str(i): "w%d" % i for i in range(1, cfg.V1)}
        self.pos = 0

    def reset_iterator(self, split):
        self.pos = 0

    def get_vocab(self):
        return self.vocab

    def get_batch(self, split, batch_size=None):
        b = batch_size or self.batch_size
        idx, wrapped = [], False
        for _ in range(b):
            idx.append(self.pos)
            self.pos += 1
            if self.pos >= self.n:
                self.pos, wrapped = 0, True
        rep = np.repeat(np.array(idx), self.seq_per_img)
        rows = np.concatenate([np.arange(i * self.seq_per_img, (i + 1) * self.seq_per_img) for i in idx])
        return {"fc_feats_array": [f.numpy()[rep] for f in self.fc], "att_feats_array": [a.numpy()[rep] for a in self.att],
                "labels": self.labels.numpy()[rows], "masks": self.masks.numpy()[rows], "top_words": self.top.numpy()[rows],
                "infos": [{"id": 1000 + i} for i in idx],
                "bounds": {"it_pos_now": self.pos, "it_max": self.n, "wrapped": wrapped}}


def test_decode_sequence_stops_at_first_zero():
    from recurrent_fusion_network_b200.eval_utils import decode_sequence
    vocab = {"1": "a", "2": "b", "3": "c"}
    assert decode_sequence(vocab, torch.tensor([[1, 2, 0, 3], [0, 1, 1, 1], [3, 3, 3, 3]])) == ["a b", "", "c c c c"]


@pytest.mark.gpu
def test_eval_split_matches_direct_calls():
    from recurrent_fusion_network_b200 import eval_utils as EU
    from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
    from tests._gpu_util import build_model, cuda_list
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    m = build_model(cfg, sd)
    crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=0, label_smoothing_epsilon=0.1, use_cuda=1))
    loader = FakeLoader(cfg, n_images=7, batch_size=3, seq_per_img=2)
    kw = {"eval_split": "val", "val_images_use": 7, "beam_size": 3, "language_eval": 0, "verbose": False,
          "feature_type": "feat_array", "reason_weight": 10}
    loss, preds, stats = EU.eval_split(m, crit, loader, kw)
    assert stats is None and m.training                                   # switched back to train mode (:261)
    assert [p["image_id"] for p in preds] == [1000 + i for i in range(7)]  # 9 decoded, 2 beyond the split end popped
    with torch.no_grad():
        m.eval()
        seq = m.sample(cuda_list(loader.fc), cuda_list(loader.att), {"beam_size": 3})[0]
        want = EU.decode_sequence(loader.vocab, seq)
        assert [p["caption"] for p in preds] == want
        # loss of the first batch through the same public calls
        loader.reset_iterator("val")
        d = loader.get_batch("val")
        lp, tp = m([torch.from_numpy(a).cuda() for a in d["fc_feats_array"]], [torch.from_numpy(a).cuda() for a in d["att_feats_array"]],
                   torch.from_numpy(d["labels"]).cuda())
        l0 = float(crit(lp, torch.from_numpy(d["labels"]).cuda()[:, 1:], torch.from_numpy(d["masks"]).cuda()[:, 1:], tp,
                        torch.from_numpy(d["top_words"]).cuda(), 10))
    loader2 = FakeLoader(cfg, n_images=7, batch_size=3, seq_per_img=2)
    loss1, preds1, _ = EU.eval_split(m, crit, loader2, dict(kw, val_images_use=3))
    assert abs(loss1 - l0) <= 1e-5 * max(1.0, abs(l0)) and len(preds1) == 3
    with pytest.raises(RuntimeError):
        EU.eval_split(m, crit, FakeLoader(cfg, 3, 3, 2), dict(kw, language_eval=1, val_images_use=3))
    seen = {}
    EU.eval_split(m, crit, FakeLoader(cfg, 3, 3, 2), dict(kw, language_eval=1, val_images_use=3, id="x",
                                                          language_eval_fn=lambda ds, p, mid, sp: seen.update(n=len(p), id=mid) or {"CIDEr": 0.0}))
    assert seen == {"n": 3, "id": "eval_split_x_0"}


@pytest.mark.gpu
def test_eval_ensemble_matches_ensemble_beam():
    from recurrent_fusion_network_b200 import eval_utils as EU
    from recurrent_fusion_network_b200.ensemble import ensemble_sample_beam
    from tests._gpu_util import build_model, cuda_list
    cfg = O.tiny_config(2)
    models = [build_model(cfg, O.make_state_dict(cfg, seed=s, init_range=0.5, logit_scale=3.0, eos_bias=0.8)) for s in (1250, 1251)]
    loader = FakeLoader(cfg, n_images=5, batch_size=2, seq_per_img=2)
    _, preds, _ = EU.eval_ensemble(models, loader, {"eval_split": "test", "num_images": 5, "beam_size": 3, "batch_size": 2,
                                                    "verbose": False, "language_eval": 0})
    with torch.no_grad():
        seq, slp = ensemble_sample_beam(models, cuda_list(loader.fc), cuda_list(loader.att), {"beam_size": 3})[:2]
    assert [p["caption"] for p in preds] == EU.decode_sequence(loader.vocab, seq)
    want_lp = (slp * (seq > 0).float()).sum(1).cpu()
    assert max(abs(p["log_prob"] - float(w)) for p, w in zip(preds, want_lp)) <= 1e-4


# ---- against the restated oracle of the reference's drivers (oracle/eval_oracle.py) ---------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("beam_size,val_images_use,n_images,batch", [(3, 7, 7, 3), (1, 4, 7, 3), (3, -1, 5, 2), (1, 100, 5, 2), (3, 2, 6, 4)])
def test_eval_split_matches_reference_driver_oracle(beam_size, val_images_use, n_images, batch):
    """Mean loss, the prediction list (ids, captions, what gets popped) and both break conditions of eval_utils.py:66-265,
    incl. the quirk that val_images_use = -1 stops after the first batch (`n >= -1`).  The oracle's results for these very cases
    are pinned against the reference's own eval_split (tests/golden/eval_split_cases.json, oracle/gen_golden_eval.py)."""
    from oracle import eval_oracle as EO
    from recurrent_fusion_network_b200 import eval_utils as EU
    from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
    from tests._gpu_util import build_model
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    m = build_model(cfg, sd)
    crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=0, label_smoothing_epsilon=0.1, use_cuda=1))
    kw = {"eval_split": "val", "val_images_use": val_images_use, "beam_size": beam_size, "language_eval": 0, "verbose": False,
          "feature_type": "feat_array", "reason_weight": 10, "sample_max": 1}
    loss, preds, _ = EU.eval_split(m, crit, FakeLoader(cfg, n_images, batch, 2, seed=3), kw)
    want_loss, want_preds = EO.eval_split(sd, cfg, FakeLoader(cfg, n_images, batch, 2, seed=3), kw)
    assert preds == want_preds
    assert abs(loss - want_loss) <= 1e-4 * max(1.0, abs(want_loss))
    import json, os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_split_cases.json")))
    ref = [c for c in fx["cases"] if (c["beam_size"], c["val_images_use"], c["n_images"], c["batch"]) == (beam_size, val_images_use, n_images, batch)]
    assert len(ref) == 1 and preds == ref[0]["predictions"]            # the REFERENCE driver's own output
    assert abs(loss - ref[0]["loss"]) <= 1e-4 * max(1.0, abs(ref[0]["loss"]))


@pytest.mark.gpu
@pytest.mark.parametrize("beam_size,num_images,n_images,batch", [(3, 5, 5, 2), (1, 5, 5, 2), (1, -1, 6, 3), (3, 2, 6, 3)])
def test_eval_ensemble_beam_and_greedy_match_reference_driver_oracle(beam_size, num_images, n_images, batch):
    """eval_ensemble (beam, eval_utils.py:387-719) and eval_ensemble_greedy (:729-975, what eval_ensemble.sh runs)."""
    from oracle import eval_oracle as EO
    from recurrent_fusion_network_b200 import eval_utils as EU
    from tests._gpu_util import build_model
    cfg = O.tiny_config(2)
    sds = [O.make_state_dict(cfg, seed=s, init_range=0.5, logit_scale=3.0, eos_bias=0.8) for s in (1250, 1251, 1252)]
    models = [build_model(cfg, sd) for sd in sds]
    kw = {"eval_split": "test", "num_images": num_images, "beam_size": beam_size, "batch_size": batch, "verbose": False,
          "language_eval": 0}
    _, preds, _ = EU.eval_ensemble(models, FakeLoader(cfg, n_images, batch, 2, seed=5), kw)
    want = EO.eval_ensemble(sds, cfg, FakeLoader(cfg, n_images, batch, 2, seed=5), kw)
    assert [(p["image_id"], p["caption"]) for p in preds] == [(p["image_id"], p["caption"]) for p in want]
    assert max(abs(p["log_prob"] - q["log_prob"]) for p, q in zip(preds, want)) <= 2e-4
    if beam_size == 1:
        _, preds_g, _ = EU.eval_ensemble_greedy(models, FakeLoader(cfg, n_images, batch, 2, seed=5), dict(kw, beam_size=7))
        assert [p["caption"] for p in preds_g] == [p["caption"] for p in want]


@pytest.mark.gpu
def test_ensemble_greedy_full_size_two_models():
    """Greedy logit-mean ensemble at reference sizes (two five-encoder models), tokens and log-probs vs the oracle."""
    from oracle import eval_oracle as EO
    from recurrent_fusion_network_b200.ensemble import ensemble_sample_greedy
    from tests._gpu_util import LP_TOL, assert_tokens_match_with_tie_policy, build_model, cuda_list, maxdiff
    cfg = O.RFNConfig()
    sds = [O.make_state_dict(cfg, seed=1235 + i, sharpen=True) for i in range(2)]
    fc, att = O.make_inputs(cfg, 4, seed=12)
    models = [build_model(cfg, sd) for sd in sds]
    torch.set_num_threads(16)
    seq, slp = ensemble_sample_greedy(models, cuda_list(fc), cuda_list(att))
    with torch.no_grad():
        oseq, oslp = EO.ensemble_sample_greedy(sds, cfg, fc, att)
    assert seq.shape == oseq.shape and torch.equal(seq.cpu(), oseq)
    assert maxdiff(slp, oslp) <= LP_TOL


@pytest.mark.gpu
def test_ensemble_step_matches_reference_hook_fixture():
    """ensemble.model_ensemble_feat_array_one_step (per-model one_time_step + rfn_mean_log_softmax_f32) against the log-probs the
    REFERENCE's own hook (eval_utils.py:268-290) produced on three reference models, two consecutive greedy steps
    (tests/golden/ensemble_step_case.npz, written by oracle/gen_golden_eval.py)."""
    import os
    from recurrent_fusion_network_b200.ensemble import model_ensemble_feat_array_one_step
    from tests._gpu_util import LP_TOL, build_model, cuda_list, maxdiff
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "ensemble_step_case.npz"))
    cfg = O.tiny_config(2)
    sds = [O.make_state_dict(cfg, seed=int(s), init_range=0.5, logit_scale=3.0, eos_bias=0.8) for s in fx["seeds"]]
    models = [build_model(cfg, sd) for sd in sds]
    rows = int(fx["rows"])
    fc, att = O.make_inputs(cfg, rows, seed=int(fx["input_seed"]))
    with torch.no_grad():
        tvs, sts = [], []
        for m in models:
            tv, _, st = m.get_thought_vectors(cuda_list(fc), cuda_list(att), m.get_init_state(cuda_list(fc)))
            tvs.append(tv); sts.append(st)
        tok = torch.zeros(rows, dtype=torch.int64, device="cuda")
        for step in range(fx["logprobs"].shape[0]):
            xts = [m.embed.weight[tok] for m in models]
            _, sts, lp = model_ensemble_feat_array_one_step(models, xts, sts, tvs)
            want = torch.from_numpy(fx["logprobs"][step])
            assert maxdiff(lp, want) <= LP_TOL
            assert torch.equal(lp.argmax(1).cpu(), want.argmax(1))
            tok = lp.argmax(1)
