"""Fixture for the review cell of ReviewNet, LSTMSoftAttentionNoInputCore, written by the REFERENCE module itself.  TEST INFRASTRUCTURE.

Run in the build container only (it needs /root/reference):

    python oracle/gen_golden_cores.py       # (re)writes tests/golden/review_core_cases.npz

The full RecurrentFusionModel never instantiates misc/LSTMSoftAttentionNoInputCore.py (stage 1 uses the
feat-array variant), so the path-level fixtures of oracle/gen_golden.py do not exercise it; SURVEY 8(a) row a4 keeps it for API
parity.  Here the reference class is imported as is (no text edits), initialised by its own constructor under a fixed seed and
run for three chained steps on seeded inputs, plain and with maxout; the fixture holds its state_dict, the inputs and ITS
outputs.  tests/test_gpu_parity.py loads the same state_dict into our mirror module and compares."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
warnings.filterwarnings("ignore")


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from misc.LSTMSoftAttentionNoInputCore import LSTMSoftAttentionNoInputCore as RefCore
    out = {}
    names = []
    for name, (R, D, N, A, maxout, rows) in {"plain": (32, 24, 7, 16, 0, 6), "maxout": (48, 40, 9, 16, 1, 5)}.items():
        torch.manual_seed(11 if maxout else 7)
        core = RefCore(R, D, N, A, 0.0, maxout).eval()
        core.att_h_2_out.bias.data.fill_(0.3)          # (cancels in the softmax; kept non-zero to check that it does)
        with torch.no_grad():                          # the constructor's init (weights in +-0.1, gate biases -1) keeps |h| < 0.04:
            core.h2h.weight.mul_(4.0); core.z2h.weight.mul_(2.0)          # widen it so that the gates leave their linear range
            core.h2h.bias.uniform_(-0.5, 0.5); core.z2h.bias.uniform_(-0.5, 0.5)
            core.att_2_att_h.weight.mul_(3.0); core.h_2_att_h.weight.mul_(3.0); core.att_h_2_out.weight.mul_(10.0)
        g = torch.Generator().manual_seed(21 + maxout)
        att = torch.randn(rows, N, D, generator=g)
        h = torch.randn(1, rows, R, generator=g) * 0.5
        c = torch.randn(1, rows, R, generator=g) * 0.5
        hs, cs = [], []
        with torch.no_grad():
            state = (h, c)
            for _ in range(3):
                o, state = core(att, None, None, state)
                assert torch.equal(o, state[0][0])
                hs.append(state[0][0].clone()); cs.append(state[1][0].clone())
        names.append(name)
        out[f"{name}.dims"] = np.array([R, D, N, A, maxout, rows])
        for k, v in core.state_dict().items():
            out[f"{name}.sd.{k}"] = v.numpy()
        out[f"{name}.att"], out[f"{name}.h0"], out[f"{name}.c0"] = att.numpy(), h.numpy(), c.numpy()
        out[f"{name}.h"], out[f"{name}.c"] = torch.stack(hs).numpy(), torch.stack(cs).numpy()
        print(f"[review core {name}] reference module: R={R} D={D} N={N} A={A} maxout={maxout}, 3 steps, |h| max {float(torch.stack(hs).abs().max()):.4f}")
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "review_core_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
