"""CIDEr-D reward scorer: oracle vs the reference scorer's outputs (CPU), CUDA kernel vs both (GPU)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ciderd_oracle as CD
from tests._golden import GOLDEN


def _load():
    d = dict(np.load(os.path.join(GOLDEN, "ciderd.npz")))
    n_img, spi = int(d["n_img"]), int(d["spi"])
    gts = [[d["gts"][i, j] for j in range(int(d["n_refs"][i]))] for i in range(n_img)]
    dfi = {tuple(int(t) for t in k if t >= 0): float(v) for k, v in zip(d["df_keys"], d["df_vals"])}
    return d, gts, dfi, n_img, spi


def test_oracle_matches_reference_scorer_fixture():
    d, gts, dfi, n_img, spi = _load()
    rows = d["gen"].shape[0]
    hyps = [CD.caption_tokens(x) for x in list(d["gen"]) + list(d["greedy"])]
    gt_tok = [[CD.caption_tokens(g) for g in gts[i]] for i in range(n_img)]
    refs = [gt_tok[(i % rows) // spi] for i in range(2 * rows)]
    a = CD.ciderd_scores(hyps, refs, CD.corpus_document_frequency(refs), np.log(float(len(refs))))
    b = CD.ciderd_scores(hyps, refs, dfi, np.log(113287.0))
    assert np.abs(a - d["scores_corpus"]).max() <= 1e-12
    assert np.abs(b - d["scores_table"]).max() <= 1e-12
    rew, sc = CD.self_critical_reward(d["gen"], d["greedy"], gts, None, None, spi)
    assert rew.shape == d["gen"].shape and np.abs(rew[:, 0] - (d["scores_corpus"][:rows] - d["scores_corpus"][rows:])).max() <= 1e-12


@pytest.mark.gpu
def test_device_scorer_matches_reference_scorer_fixture():
    from recurrent_fusion_network_b200 import reward as RW
    d, gts, dfi, n_img, spi = _load()
    rows, T = d["gen"].shape
    gen = torch.from_numpy(d["gen"]).cuda()
    greedy = torch.from_numpy(d["greedy"]).cuda()
    opt = SimpleNamespace(cider_weight=1.0, bleu4_weight=0, spice_weight=0, use_baseline=1)
    # corpus document frequencies (CiderD(df='corpus'))
    rew, sc = RW.compute_reward(gen, greedy, gts, None, opt, seq_per_img=spi)
    assert np.abs(sc.cpu().numpy() - d["scores_corpus"]).max() <= 1e-11
    want = (d["scores_corpus"][:rows] - d["scores_corpus"][rows:]).astype(np.float32)
    assert rew.shape == (rows, T) and np.abs(rew.cpu().numpy() - want[:, None]).max() <= 1e-6
    # injected table, 'coco-train' reference length
    table = RW.DocumentFrequency(dfi, 113287, gen.device)
    rew2, sc2 = RW.compute_reward(gen, greedy, gts, table, opt, seq_per_img=spi)
    assert np.abs(sc2.cpu().numpy() - d["scores_table"]).max() <= 1e-11
    # the string interface of ciderD.py
    hyps = [CD.caption_tokens(x) for x in list(d["gen"]) + list(d["greedy"])]
    tostr = lambda t: " ".join(str(x) for x in t)
    res = [{"image_id": i, "caption": [tostr(h)]} for i, h in enumerate(hyps)]
    gts_s = {i: [tostr(CD.caption_tokens(g)) for g in gts[(i % rows) // spi]] for i in range(2 * rows)}
    # hypothesis 3 has no terminating 0 (16 real tokens): the string API requires the array_to_str convention
    ok = [i for i, h in enumerate(hyps) if h[-1] == 0]
    mean, scores = RW.CiderD(df="coco-train", document_frequency=dfi, n_documents=113287).compute_score(
        gts_s, [res[i] for i in ok])
    assert np.abs(scores - d["scores_table"][ok]).max() <= 1e-11 and abs(mean - scores.mean()) < 1e-12


@pytest.mark.gpu
def test_self_critical_reward_end_to_end():
    """get_self_critical_reward_feat_array: greedy decode on the device + CIDEr-D on the device == oracle pipeline."""
    from oracle import rfnet_oracle as O
    from recurrent_fusion_network_b200 import reward as RW
    from tests._gpu_util import build_model, cuda_list
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    imgs, spi = 3, 2
    fc, att = O.make_inputs(cfg, imgs, seed=8)
    fc = [f.repeat_interleave(spi, 0) for f in fc]
    att = [a.repeat_interleave(spi, 0) for a in att]
    rows = imgs * spi
    g = torch.Generator().manual_seed(1)
    u = torch.rand(rows, cfg.seq_length, generator=g)
    rng = np.random.RandomState(2)
    gts = []
    for _ in range(imgs):
        refs = []
        for _ in range(3):
            a = np.zeros(cfg.seq_length + 1, dtype=np.int64)
            n = rng.randint(1, cfg.seq_length)
            a[:n] = rng.randint(1, cfg.V1, size=n)
            refs.append(a)
        gts.append(refs)
    m = build_model(cfg, sd)
    with torch.no_grad():
        gen, *_ = m.sample(cuda_list(fc), cuda_list(att), {"sample_max": 0, "uniforms": u.cuda()})
        gen_o, *_ = O.sample(sd, cfg, fc, att, sample_max=0, uniforms=u)
        greedy_o, *_ = O.sample(sd, cfg, fc, att, sample_max=1)
    assert torch.equal(gen.cpu(), gen_o)
    opt = SimpleNamespace(cider_weight=1.0, bleu4_weight=0, spice_weight=0, use_baseline=1)
    got = RW.get_self_critical_reward_feat_array(None, m, cuda_list(fc), cuda_list(att), {"gts": gts}, gen, opt)
    T = gen_o.shape[1]
    greedy_p = torch.nn.functional.pad(greedy_o, (0, max(0, T - greedy_o.shape[1])))[:, :T]
    want, _ = CD.self_critical_reward(gen_o.numpy(), greedy_p.numpy(), gts, None, None, spi)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-5


def _reward_cases():
    d = np.load(os.path.join(GOLDEN, "reward_cases.npz"))
    for name in d["names"]:
        n_img, spi, use_baseline = (int(v) for v in d[f"{name}.meta"])
        gts = [[d[f"{name}.gts"][i, j] for j in range(int(d[f"{name}.n_refs"][i]))] for i in range(n_img)]
        yield name, d[f"{name}.gen"], d[f"{name}.greedy"], gts, spi, bool(use_baseline), float(d[f"{name}.cider_weight"]), d[f"{name}.rewards"]


def test_oracle_reward_assembly_matches_reference_compute_reward_fixture():
    """get_rewards.py:39-112 compute_reward run from the reference's source text (oracle/gen_golden_reward.py): array_to_str, the
    hypothesis / reference bookkeeping, sampled minus greedy (or sampled alone), the weight, the repeat along T."""
    for name, gen, greedy, gts, spi, use_baseline, w, want in _reward_cases():
        got, _ = CD.self_critical_reward(gen, greedy, gts, None, None, spi, cider_weight=w, use_baseline=use_baseline)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-12, name


@pytest.mark.gpu
def test_device_reward_matches_reference_compute_reward_fixture():
    from recurrent_fusion_network_b200 import reward as RW
    for name, gen, greedy, gts, spi, use_baseline, w, want in _reward_cases():
        opt = SimpleNamespace(cider_weight=w, bleu4_weight=0, spice_weight=0, use_baseline=1 if use_baseline else 0)
        rew, _ = RW.compute_reward(torch.from_numpy(gen).cuda(), torch.from_numpy(greedy).cuda(), gts, None, opt, seq_per_img=spi)
        assert rew.shape == want.shape and np.abs(rew.cpu().numpy() - want).max() <= 1e-6, name
