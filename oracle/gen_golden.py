"""Pin the oracle against the REAL reference and write the golden fixtures.  TEST INFRASTRUCTURE.

Run in the build container only (it needs /root/reference, which never travels):

    python oracle/gen_golden.py            # validates + (re)writes tests/golden/*.npz

What it does
  1. imports the reference's own modules read-only from /root/reference (no bytecode written),
     builds ``RecurrentFusionModel(opt)`` through ``opts.parse_opt()`` exactly as main.py does,
  2. loads the oracle's deterministic synthetic weights into it with ``load_state_dict`` (this also
     proves the 773 state_dict key names / shapes match),
  3. runs the reference's ``forward``, greedy ``sample``, ``sample_beam`` and both criteria on seeded
     synthetic inputs, and checks the oracle restatement against them,
  4. stores the REFERENCE's outputs (never the oracle's) as fixtures; weights and inputs are
     regenerated from their seeds at test time and guarded by checksums stored in the fixture.

``sample_beam`` does not run on torch >= 0.4 as written: the four 0-dim indexing
expressions at misc/RecurrentFusionModel.py:477-478 and :513 are rewritten to ``.item()`` in the
source TEXT at import time (nothing is copied into this repo).
"""
from __future__ import annotations

import argparse
import dataclasses
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import rfnet_oracle as O  # noqa: E402

warnings.filterwarnings("ignore")


def import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    path = os.path.join(REF, "misc", "RecurrentFusionModel.py")
    src = open(path).read()
    edits = [
        ("'c': ix.data[q, c],", "'c': ix.data[q, c].item(),"),
        ("'p': candidate_logprob.data[0],", "'p': candidate_logprob.item(),"),
        ("'r': local_logprob.data[0]}", "'r': local_logprob.item()}"),
        ("'p': beam_logprobs_sum[vix]\n", "'p': beam_logprobs_sum[vix].item()\n"),
    ]
    for a, b in edits:
        assert src.count(a) == 1, a
        src = src.replace(a, b)
    mod = types.ModuleType("rfm_patched")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    import opts  # reference's opts.py
    import misc.utils as ref_utils
    return mod, opts, ref_utils


def build_reference_model(mod, opts, cfg: O.RFNConfig, fusion_maxout: int = 0):
    argv = sys.argv
    sys.argv = ["x", "--caption_model", "recurrent_fusion_model", "--feature_type", "feat_array",
                "--use_cuda", "0"]
    try:
        opt = opts.parse_opt()
    finally:
        sys.argv = argv
    opt.vocab_size = cfg.vocab_size
    opt.seq_length = cfg.seq_length
    opt.rnn_size = cfg.rnn_size
    opt.att_hid_size = cfg.att_hid_size
    opt.input_encoding_size = cfg.input_encoding_size
    opt.num_review_steps_0 = cfg.num_review_steps_0
    opt.num_review_steps = cfg.num_review_steps
    opt.top_words_count = cfg.top_words_count
    opt.review_maxout = cfg.review_maxout      # opts.py:182
    opt.maxout = cfg.decoder_maxout            # opts.py:180
    opt.fusion_maxout = fusion_maxout          # opts.py:184: read by the model, never forwarded to the cells (:93-97)
    opt.feat_array_info = [dict(fc_feat_size=e.fc_feat_size, att_feat_size=e.att_feat_size,
                                att_num=e.att_num) for e in cfg.encoders]
    model = mod.RecurrentFusionModel(opt)
    model.eval()
    return model, opt


def checksum(tensors) -> float:
    return float(sum(t.double().abs().sum() for t in tensors))


def maxdiff(a, b) -> float:
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0


def pad_top(top_seq, top_prob, L):
    n = max(t.shape[0] for t in top_seq)
    B = len(top_seq)
    ts = np.zeros((B, n, L), dtype=np.int64)
    tp = np.full((B, n), np.nan, dtype=np.float64)
    nd = np.zeros((B,), dtype=np.int64)
    for k in range(B):
        m = top_seq[k].shape[0]
        ts[k, :m] = top_seq[k].numpy()
        tp[k, :m] = np.asarray(top_prob[k], dtype=np.float64)
        nd[k] = m
    return ts, tp, nd


def run_case(mod, opts, ref_utils, name, cfg, rows, wseed, iseed, wkw, stride, beam=3,
             do_beam=True, do_loss=True, log=print):
    sd = O.make_state_dict(cfg, seed=wseed, **wkw)
    sharpen = bool(wkw)
    model, opt = build_reference_model(mod, opts, cfg)
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key order/name mismatch"
    model.load_state_dict(sd)
    fc, att = O.make_inputs(cfg, rows, seed=iseed)
    labels, masks, top_words = O.make_labels(cfg, rows, seed=iseed + 100)
    out = dict(rows=rows, wseed=wseed, iseed=iseed, wkw=repr(sorted(wkw.items())), stride=stride, beam=beam,
               w_checksum=checksum(sd.values()), in_checksum=checksum(fc + att),
               lab_checksum=float(labels.sum()))
    worst = {}
    with torch.no_grad():
        # ---- forward (XE, teacher forced) ------------------------------------------------
        lp_ref, rp_ref = model(fc, att, labels)
        lp_o, rp_o = O.forward_xe(sd, cfg, fc, att, labels)
        worst["xe_lp"] = maxdiff(lp_ref, lp_o)
        worst["xe_reason"] = max(maxdiff(a.reshape(rows, -1), b) for a, b in zip(rp_ref, rp_o))
        out["xe_lp_strided"] = lp_ref[:, :, ::stride].numpy()
        out["xe_lp_target"] = lp_ref.gather(2, labels[:, 1:lp_ref.shape[1] + 1].unsqueeze(2)).squeeze(2).numpy()
        out["xe_T"] = lp_ref.shape[1]
        out["reason_pred_strided"] = np.stack([r.reshape(rows, -1)[:, ::max(1, stride // 8)].numpy() for r in rp_ref])
        # ---- greedy sample ----------------------------------------------------------------
        s_ref, sl_ref, la_ref, _ = model.sample(fc, att, {"sample_max": 1, "beam_size": 1})
        s_o, sl_o, la_o, _ = O.sample(sd, cfg, fc, att, sample_max=1)
        assert s_ref.shape == s_o.shape, (s_ref.shape, s_o.shape)
        worst["greedy_tok_mismatch"] = int((s_ref != s_o).sum())
        worst["greedy_slp"] = maxdiff(sl_ref, sl_o)
        worst["greedy_lp_all"] = maxdiff(la_ref, la_o)
        out["greedy_seq"] = s_ref.numpy()
        out["greedy_slp"] = sl_ref.numpy()
        out["greedy_lp_all_strided"] = la_ref[:, :, ::stride].numpy()
        out["greedy_min_margin"] = float(O.top2_margin(la_ref[:, :s_ref.shape[1]]).min())
        # ---- beam -------------------------------------------------------------------------
        if do_beam:
            bs_ref, bl_ref, ts_ref, tp_ref, _ = model.sample_beam(fc, att, {"beam_size": beam})
            bs_o, bl_o, ts_o, tp_o, _ = O.sample_beam(sd, cfg, fc, att, beam_size=beam)
            worst["beam_tok_mismatch"] = int((bs_ref != bs_o).sum())
            worst["beam_lp"] = maxdiff(bl_ref, bl_o)
            assert [t.shape for t in ts_ref] == [t.shape for t in ts_o]
            worst["beam_top_seq_mismatch"] = int(sum((a != b).sum() for a, b in zip(ts_ref, ts_o)))
            worst["beam_top_prob"] = max(abs(x - y) for a, b in zip(tp_ref, tp_o) for x, y in zip(a, b))
            out["beam_seq"] = bs_ref.numpy()
            out["beam_lp"] = bl_ref.numpy()
            ts, tp, nd = pad_top(ts_ref, tp_ref, cfg.seq_length)
            out["beam_top_seq"], out["beam_top_prob"], out["beam_n_done"] = ts, tp, nd
    # ---- criteria (with autograd in the reference; values only here) --------------------------
    if do_loss:
        class _O:  # the criterion reads these off ``opt``
            use_label_smoothing = 1
            label_smoothing_epsilon = 0.1
            use_cuda = 0
            use_ppo = 0
        with torch.no_grad():
            crit = ref_utils.ReviewNetEnsembleCriterion(_O)
            rp_list = [r.reshape(rows, -1) for r in rp_ref]
            l_ref = crit(lp_ref, labels[:, 1:], masks[:, 1:], rp_list, top_words, 10.0)
            l_o = O.xe_loss(lp_o, labels[:, 1:], masks[:, 1:], rp_o, top_words, 10.0, 0.1)
            worst["xe_loss"] = abs(float(l_ref) - float(l_o))
            out["xe_loss_ls"] = float(l_ref)
            _O.use_label_smoothing = 0
            crit = ref_utils.ReviewNetEnsembleCriterion(_O)
            l_ref0 = crit(lp_ref, labels[:, 1:], masks[:, 1:], rp_list, top_words, 10.0)
            l_o0 = O.xe_loss(lp_o, labels[:, 1:], masks[:, 1:], rp_o, top_words, 10.0, 0.0)
            worst["xe_loss_nols"] = abs(float(l_ref0) - float(l_o0))
            out["xe_loss_nols"] = float(l_ref0)
            g = torch.Generator().manual_seed(iseed + 200)
            reward = torch.randn(rows, 1, generator=g).expand(rows, s_ref.shape[1]).contiguous()
            rl = ref_utils.ReviewNetRewardCriterion(_O)
            r_ref = rl(sl_ref, s_ref, reward, la_ref, 0.01, rp_list, top_words, 10.0, None, _O)
            r_o = O.rl_loss(sl_o, s_o, reward, la_o, 0.01, rp_o, top_words, 10.0)
            worst["rl_loss"] = abs(float(r_ref) - float(r_o))
            out["rl_loss"] = float(r_ref)
            out["rl_reward_seed"] = iseed + 200
    log(f"[{name}] oracle-vs-reference: " + ", ".join(f"{k}={v:.3g}" for k, v in worst.items()))
    # gates: tokens exact, floats within 2e-6 (reference fp32-vs-fp64 drift is ~1.2e-6)
    for k, v in worst.items():
        if k.endswith("mismatch"):
            assert v == 0, (name, k, v)
        else:
            assert v < 5e-5 if sharpen else v < 5e-6, (name, k, v)
    out["oracle_vs_reference"] = np.array([f"{k}={v:.3g}" for k, v in worst.items()])
    return out


def ciderd_case(log):
    """CIDEr-D: reference scorer (cider/pyciderevalcap/ciderD/ciderD_scorer.py) vs oracle/ciderd_oracle.py on random
    token captions, 'corpus' document frequencies and an injected 'coco-train'-style table."""
    sys.path.insert(0, os.path.join(REF, "cider"))
    from pyciderevalcap.ciderD.ciderD_scorer import CiderScorer
    from oracle import ciderd_oracle as CD
    from collections import defaultdict
    rng = np.random.RandomState(0)
    L, V, n_img, spi = 16, 30, 6, 5

    def cap(maxlen):
        n = rng.randint(1, maxlen)
        a = np.zeros(maxlen, dtype=np.int64)
        a[:n] = rng.randint(1, V, size=n)
        return a

    rows = n_img * spi
    gen = np.stack([cap(L) for _ in range(rows)]); gen[3] = rng.randint(1, V, size=L)   # one caption without any 0
    greedy = np.stack([cap(L) for _ in range(rows)]); greedy[7, 0] = 0                   # one empty caption
    gts = [[cap(L + 1) for _ in range(rng.randint(1, 6))] for _ in range(n_img)]
    hyps = [CD.caption_tokens(x) for x in list(gen) + list(greedy)]
    gt_tok = [[CD.caption_tokens(g) for g in gts[i]] for i in range(n_img)]
    refs = [gt_tok[(i % rows) // spi] for i in range(2 * rows)]
    tostr = lambda t: " ".join(str(x) for x in t)
    out = dict(gen=gen, greedy=greedy, n_img=n_img, spi=spi,
               gts=np.stack([np.stack([np.pad(g, (0, 0)) for g in (gts[i] + [np.zeros(L + 1, dtype=np.int64)] * (5 - len(gts[i])))]) for i in range(n_img)]),
               n_refs=np.array([len(g) for g in gts]))
    # corpus mode
    sc = CiderScorer(df_mode="corpus")
    for h, r in zip(hyps, refs):
        sc += (tostr(h), [tostr(x) for x in r])
    _, ref_scores = sc.compute_score()
    mine = CD.ciderd_scores(hyps, refs, CD.corpus_document_frequency(refs), np.log(float(len(refs))))
    d1 = float(np.abs(mine - ref_scores).max())
    out["scores_corpus"] = ref_scores
    # injected document frequencies, 'coco-train' reference length
    dfi = {}
    for ng in list(CD.corpus_document_frequency(refs).keys())[::3]:
        dfi[ng] = float(rng.randint(1, 5000))
    sc2 = CiderScorer(df_mode="corpus")
    sc2.df_mode = "coco-train-synthetic"
    d = defaultdict(float)
    for ng, v in dfi.items():
        d[tuple(str(t) for t in ng)] = v
    sc2.document_frequency = d
    for h, r in zip(hyps, refs):
        sc2 += (tostr(h), [tostr(x) for x in r])
    _, ref_scores2 = sc2.compute_score()
    mine2 = CD.ciderd_scores(hyps, refs, dfi, np.log(float(113287)))
    d2 = float(np.abs(mine2 - ref_scores2).max())
    out["scores_table"] = ref_scores2
    out["df_keys"] = np.array([list(k) + [-1] * (4 - len(k)) for k in dfi.keys()], dtype=np.int64)
    out["df_vals"] = np.array(list(dfi.values()))
    log(f"[ciderd] oracle-vs-reference: corpus={d1:.3g}, table={d2:.3g}")
    assert d1 < 1e-12 and d2 < 1e-12
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--skip-full", action="store_true")
    ap.add_argument("--only", default=None, help="write this case only (and append its line to PIN_LOG.txt)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(8)
    mod, opts, ref_utils = import_reference()
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    TINY = dict(init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    SHARP = dict(sharpen=True)
    cases = [
        ("tiny_j2", O.tiny_config(2), 3, 1234, 7, {}, 1),
        ("tiny_j2_eos_a", O.tiny_config(2), 4, 1248, 8, TINY, 1),
        ("tiny_j2_eos_b", O.tiny_config(2), 4, 1250, 8, TINY, 1),
        ("tiny_j3_eos_a", O.tiny_config(3), 5, 1243, 8, TINY, 1),
        ("tiny_j3_eos_b", O.tiny_config(3), 5, 1246, 8, TINY, 1),
        ("tiny_j1", O.tiny_config(1), 2, 1237, 10, dict(init_range=0.5, logit_scale=3.0), 1),
        # the maxout variants of the stage-2 and decoder cells (opts.py:180-183; default 0 in every shipped script)
        ("tiny_j2_maxout", dataclasses.replace(O.tiny_config(2), review_maxout=1, decoder_maxout=1), 4, 1248, 8, TINY, 1),
        ("tiny_j3_maxout_dec", dataclasses.replace(O.tiny_config(3), decoder_maxout=1), 5, 1243, 8, TINY, 1),
    ]
    if not args.skip_full:
        cases += [
            ("config1_n49", O.config1(49), 16, 1234, 7, {}, 97),
            ("config1_n196_sharp", O.config1(196), 4, 1234, 8, SHARP, 97),
            ("full_j5", O.RFNConfig(), 4, 1234, 7, {}, 97),
            ("full_j5_sharp", O.RFNConfig(), 4, 1234, 9, SHARP, 97),
        ]
    if args.only:
        cases = [c for c in cases if c[0] == args.only]
        assert cases, args.only
        # fusion_maxout = 1 changes nothing: FeatArrayFusionNoInputCore does not pass it on to its cells
        m1, _ = build_reference_model(mod, opts, cases[0][1], fusion_maxout=1)
        m0, _ = build_reference_model(mod, opts, cases[0][1], fusion_maxout=0)
        assert {k: tuple(v.shape) for k, v in m1.state_dict().items()} == {k: tuple(v.shape) for k, v in m0.state_dict().items()}
        log(f"[{args.only}] fusion_maxout=1 leaves every state_dict shape unchanged (never forwarded, misc/RecurrentFusionModel.py:93-97)")
    for name, cfg, rows, wseed, iseed, wkw, stride in cases:
        out = run_case(mod, opts, ref_utils, name, cfg, rows, wseed, iseed, wkw, stride, log=log)
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **out)
    if args.only:
        with open(os.path.join(args.out, "PIN_LOG.txt"), "a") as f:
            f.write("\n".join(lines) + "\n")
        return
    np.savez_compressed(os.path.join(args.out, "ciderd.npz"), **ciderd_case(log))
    with open(os.path.join(args.out, "PIN_LOG.txt"), "w") as f:
        f.write("oracle/gen_golden.py -- oracle restatement vs the imported reference modules "
                f"(torch {torch.__version__}, CPU fp32)\n")
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
