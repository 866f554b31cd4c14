"""Pin oracle/eval_oracle.eval_split against the REAL reference driver and write the fixture.  TEST INFRASTRUCTURE.

Run in the build container only (it needs /root/reference, which never travels):

    python oracle/gen_golden_eval.py        # validates + (re)writes tests/golden/eval_split_cases.json

The reference's eval_split (eval_utils.py:66-265) is executed from its SOURCE TEXT with two kinds of text edits, nothing
else (nothing is copied into this repo): `.data[0]` on 0-dim tensors -> `.item()` (the same torch >= 0.4 repair as SURVEY
D10; also `ix = seq[i, j]` -> `.item()` in misc/utils.py:25 decode_sequence) and `torch.cuda.FloatTensor` ->
`torch.FloatTensor` (no GPU here; use_cuda = 0 is the reference's own CPU switch).  It
runs on the reference's RecurrentFusionModel (imported as oracle/gen_golden.py does) with the oracle's deterministic weights,
the reference's ReviewNetEnsembleCriterion and decode_sequence, and the loader double of tests/test_eval_utils.py, which
follows the dataloader.DataLoader protocol the driver uses (reset_iterator / get_batch / get_vocab / batch_size /
seq_per_img, bounds with it_pos_now / it_max / wrapped).  What is stored is the REFERENCE's return value (mean loss and the
prediction list); the oracle restatement must reproduce it: prediction lists equal, loss within 2e-6.

eval_ensemble / eval_ensemble_greedy cannot be run this way: they call the model with stale signatures and read loader keys that do
not exist.  What CAN be pinned is the ensemble STEP they share (ensemble_step_case below)."""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types
import warnings

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import eval_oracle as EO  # noqa: E402
from oracle import rfnet_oracle as O  # noqa: E402
from oracle.gen_golden import build_reference_model, import_reference  # noqa: E402

warnings.filterwarnings("ignore")

CASES = [(3, 7, 7, 3), (1, 4, 7, 3), (3, -1, 5, 2), (1, 100, 5, 2), (3, 2, 6, 4)]   # beam, val_images_use, images, batch
WSEED, LSEED = 1250, 3


def import_reference_eval_utils():
    path = os.path.join(REF, "eval_utils.py")
    src = open(path).read()
    assert src.count(".data[0]") >= 3
    src = src.replace(".data[0]", ".item()").replace("torch.cuda.FloatTensor", "torch.FloatTensor")
    mod = types.ModuleType("ref_eval_utils_patched")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    # misc/utils.py:19-33 decode_sequence: `ix = seq[i, j]` is a 0-dim tensor on torch >= 0.4 and str(ix) is 'tensor(26)'
    upath = os.path.join(REF, "misc", "utils.py")
    usrc = open(upath).read()
    assert usrc.count("ix = seq[i, j]\n") == 1
    usrc = usrc.replace("ix = seq[i, j]\n", "ix = seq[i, j].item()\n")
    umod = types.ModuleType("ref_misc_utils_patched")
    umod.__file__ = upath
    exec(compile(usrc, upath, "exec"), umod.__dict__)
    mod.utils = umod
    return mod


def main():
    from tests.test_eval_utils import FakeLoader
    mod, opts, ref_utils = import_reference()
    ref_eval = import_reference_eval_utils()
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=WSEED, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    model, _ = build_reference_model(mod, opts, cfg)
    model.load_state_dict(sd)

    class _Opt:
        use_label_smoothing = 0
        label_smoothing_epsilon = 0.1
        use_cuda = 0
        use_ppo = 0
    crit = ref_utils.ReviewNetEnsembleCriterion(_Opt)
    out = []
    for beam, use, n_images, batch in CASES:
        kw = {"eval_split": "val", "val_images_use": use, "beam_size": beam, "language_eval": 0, "verbose": False,
              "feature_type": "feat_array", "reason_weight": 10, "sample_max": 1, "caption_model": "recurrent_fusion_model",
              "use_cuda": 0, "id": "pin"}
        with contextlib.redirect_stdout(io.StringIO()):
            loss, preds, stats = ref_eval.eval_split(model, crit, FakeLoader(cfg, n_images, batch, 2, seed=LSEED), kw)
        assert stats is None and model.training
        loss = float(loss)
        o_loss, o_preds = EO.eval_split(sd, cfg, FakeLoader(cfg, n_images, batch, 2, seed=LSEED), kw)
        d = abs(loss - o_loss)
        same = preds == o_preds
        print(f"[eval_split beam={beam} val_images_use={use} images={n_images} batch={batch}] reference: loss {loss:.6f}, "
              f"{len(preds)} predictions; oracle-vs-reference: loss diff {d:.3g}, predictions equal: {same}")
        assert same and d <= 2e-6 * max(1.0, abs(loss))
        out.append(dict(beam_size=beam, val_images_use=use, n_images=n_images, batch=batch, loss=loss, predictions=preds))
    path = os.path.join(ROOT, "tests", "golden", "eval_split_cases.json")
    json.dump(dict(weights_seed=WSEED, loader_seed=LSEED, seq_per_img=2, cases=out), open(path, "w"), indent=1)
    print("wrote", path)


def ensemble_step_case():
    """The ensemble step hook model_ensemble_feat_array_one_step (eval_utils.py:268-290), executed from the reference's source
    text on three reference models.  The hook calls model.one_time_step with a stale six-argument signature; a
    signature shim forwards (xt, fc, thought_vectors, mil, matching, state) to the model's own four-argument method -- the
    arithmetic (per-model step, running sum of the logits, division, log_softmax) is entirely the reference's.  Two consecutive
    steps are compared with the oracle's ensemble step (one_time_step per model, log_softmax(sum(logits) / M))."""
    mod, opts, _ = import_reference()
    ref_eval = import_reference_eval_utils()
    cfg = O.tiny_config(2)
    seeds = (1250, 1251, 1252)
    sds = [O.make_state_dict(cfg, seed=s, init_range=0.5, logit_scale=3.0, eos_bias=0.8) for s in seeds]
    models = []
    for sd in sds:
        m, _ = build_reference_model(mod, opts, cfg)
        m.load_state_dict(sd)
        models.append(m)

    class Shim:
        def __init__(self, m):
            self.m = m

        def one_time_step(self, xt, fc, tv, mil, matching, state):
            return self.m.one_time_step(xt, fc, tv, state)

    rows = 5
    fc, att = O.make_inputs(cfg, rows, seed=9)
    worst = 0.0
    ref_lps = []
    with torch.no_grad():
        r_tv, r_st, o_tv, o_st = [], [], [], []
        for m, sd in zip(models, sds):
            tv, _, st = m.get_thought_vectors(fc, att, m.get_init_state(fc))
            r_tv.append(tv); r_st.append(st)
            tvo, _, sto = O.get_thought_vectors(sd, cfg, att, O.get_init_state(sd, cfg, fc))
            o_tv.append(tvo); o_st.append(sto)
        tok = torch.zeros(rows, dtype=torch.int64)
        for step in range(2):
            xt_list = [m.embed(tok) for m in models]
            _, r_st, r_lp = ref_eval.model_ensemble_feat_array_one_step([Shim(m) for m in models], xt_list, fc, att, None, None,
                                                                        r_st, r_tv, "recurrent_fusion_model")
            logits = []
            for k, sd in enumerate(sds):
                lg, o_st[k] = O.one_time_step(sd, sd["embed.weight"][tok], o_tv[k], o_st[k])
                logits.append(lg)
            o_lp = torch.log_softmax(sum(logits) / len(sds), dim=1)
            d = float((r_lp - o_lp).abs().max())
            ds = max(float((a[0].reshape(rows, -1) - b[0]).abs().max()) for a, b in zip(r_st, o_st))
            worst = max(worst, d, ds)
            assert torch.equal(r_lp.argmax(1), o_lp.argmax(1))
            ref_lps.append(r_lp.numpy())
            tok = r_lp.argmax(1)
    print(f"[ensemble step hook, 3 models, 2 steps] oracle-vs-reference: max |log-prob / state diff| = {worst:.3g}")
    assert worst <= 2e-6
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ensemble_step_case.npz")
    np.savez_compressed(path, seeds=np.array(seeds), rows=rows, input_seed=9, logprobs=np.stack(ref_lps))
    print("wrote", path)
    return worst


if __name__ == "__main__":
    main()
    ensemble_step_case()
