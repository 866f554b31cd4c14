"""Pin the self-critical reward ASSEMBLY (get_rewards.py:39-112 compute_reward) against the reference's own function.  TEST INFRASTRUCTURE.

Run in the build container only (it needs /root/reference):

    python oracle/gen_golden_reward.py      # validates + (re)writes tests/golden/reward_cases.npz

tests/golden/ciderd.npz pins the CIDEr-D SCORER (written from the reference's CiderScorer by oracle/gen_golden.py).  What sits
around it -- array_to_str (tokens up to and including the first 0), the reference / hypothesis bookkeeping
`gts[i % batch_size // seq_per_img]`, sampled minus greedy scores, the weights, the repeat along T -- is compute_reward.  It is
executed here from the SOURCE TEXT of get_rewards.py with ONE edit: `CiderD(df='coco-train-idxs')` -> `CiderD(df='corpus')`
(data/coco-train-idxs.p is a git-LFS stub in the reference checkout; 'corpus' is the scorer's other built-in mode).  BleuD / SpiceD
are imported by the module but not called (bleu4_weight = spice_weight = 0, as in the shipped RL scripts).  Stored: the inputs
and the REFERENCE's rewards; oracle/ciderd_oracle.self_critical_reward must reproduce them to 1e-12."""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ciderd_oracle as CD  # noqa: E402

warnings.filterwarnings("ignore")


def import_reference_get_rewards():
    sys.dont_write_bytecode = True
    for p in (REF, os.path.join(REF, "cider")):
        if p not in sys.path:
            sys.path.insert(0, p)
    path = os.path.join(REF, "get_rewards.py")
    src = open(path).read()
    assert src.count("CiderD(df='coco-train-idxs')") == 1
    src = src.replace("CiderD(df='coco-train-idxs')", "CiderD(df='corpus')")
    mod = types.ModuleType("ref_get_rewards_patched")
    mod.__file__ = path
    cwd = os.getcwd()
    os.chdir(REF)            # the module appends the relative path "cider" to sys.path
    try:
        exec(compile(src, path, "exec"), mod.__dict__)
    finally:
        os.chdir(cwd)
    return mod


def main():
    ref = import_reference_get_rewards()
    rng = np.random.RandomState(5)
    out = {}
    names = []
    for name, (n_img, spi, T, V, use_baseline, w) in {"baseline": (4, 5, 16, 25, 1, 1.0), "no_baseline_w": (3, 2, 12, 12, 0, 0.5)}.items():
        rows = n_img * spi

        def cap(maxlen):
            a = np.zeros(maxlen, dtype=np.int64)
            n = rng.randint(1, maxlen)
            a[:n] = rng.randint(1, V, size=n)
            return a
        gen = np.stack([cap(T) for _ in range(rows)]); gen[1] = rng.randint(1, V, size=T)     # a caption without any 0
        greedy = np.stack([cap(T) for _ in range(rows)]); greedy[2, 0] = 0                    # an empty caption
        n_refs = rng.randint(1, 6, size=n_img)
        gts = [[cap(T + 1) for _ in range(n_refs[i])] for i in range(n_img)]
        vocab = {str(i): "w%d" % i for i in range(1, V)}     # (the SPICE strings are built even when spice_weight = 0)
        opt = SimpleNamespace(cider_weight=w, bleu4_weight=0, spice_weight=0, use_baseline=use_baseline)
        with contextlib.redirect_stdout(io.StringIO()):
            rewards = ref.compute_reward(vocab, torch.from_numpy(gen), torch.from_numpy(greedy), {"gts": gts}, opt)
        mine, _ = CD.self_critical_reward(gen, greedy, gts, None, None, spi, cider_weight=w, use_baseline=bool(use_baseline))
        d = float(np.abs(mine - rewards).max())
        print(f"[compute_reward {name}] reference: rewards {rewards.shape}, |r| max {np.abs(rewards).max():.4f}; oracle-vs-reference: {d:.3g}")
        assert rewards.shape == gen.shape and d <= 1e-12
        names.append(name)
        out[f"{name}.gen"], out[f"{name}.greedy"], out[f"{name}.rewards"] = gen, greedy, rewards
        out[f"{name}.gts"] = np.stack([np.stack(gts[i] + [np.zeros(T + 1, dtype=np.int64)] * (5 - len(gts[i]))) for i in range(n_img)])
        out[f"{name}.n_refs"] = n_refs
        out[f"{name}.meta"] = np.array([n_img, spi, use_baseline], dtype=np.int64)
        out[f"{name}.cider_weight"] = np.array(w)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "reward_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
