"""Helpers shared by the CPU and GPU parity tests: load a golden fixture (outputs of the REAL
reference, written by oracle/gen_golden.py), regenerate its weights/inputs from their seeds and
verify them against the checksums stored beside the outputs."""
import ast
import dataclasses
import os

import numpy as np
import torch

from oracle import rfnet_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CONFIGS = {
    "tiny_j1": lambda: O.tiny_config(1),
    "tiny_j2": lambda: O.tiny_config(2),
    "tiny_j2_eos_a": lambda: O.tiny_config(2),
    "tiny_j2_eos_b": lambda: O.tiny_config(2),
    "tiny_j3_eos_a": lambda: O.tiny_config(3),
    "tiny_j3_eos_b": lambda: O.tiny_config(3),
    "tiny_j2_maxout": lambda: dataclasses.replace(O.tiny_config(2), review_maxout=1, decoder_maxout=1),
    "tiny_j3_maxout_dec": lambda: dataclasses.replace(O.tiny_config(3), decoder_maxout=1),
    "config1_n49": lambda: O.config1(49),
    "config1_n196_sharp": lambda: O.config1(196),
    "full_j5": lambda: O.RFNConfig(),
    "full_j5_sharp": lambda: O.RFNConfig(),
}
TINY = [k for k in CONFIGS if k.startswith("tiny")]
FULL = [k for k in CONFIGS if not k.startswith("tiny")]


def checksum(tensors) -> float:
    return float(sum(t.double().abs().sum() for t in tensors))


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    cfg = CONFIGS[name]()
    wkw = dict(ast.literal_eval(str(d["wkw"])))
    sd = O.make_state_dict(cfg, seed=int(d["wseed"]), **wkw)
    rows = int(d["rows"])
    fc, att = O.make_inputs(cfg, rows, seed=int(d["iseed"]))
    labels, masks, top_words = O.make_labels(cfg, rows, seed=int(d["iseed"]) + 100)
    # the regenerated tensors must be the ones the reference saw
    assert abs(checksum(sd.values()) - float(d["w_checksum"])) <= 1e-9 * float(d["w_checksum"])
    assert abs(checksum(fc + att) - float(d["in_checksum"])) <= 1e-9 * float(d["in_checksum"])
    assert float(labels.sum()) == float(d["lab_checksum"])
    return cfg, sd, fc, att, labels, masks, top_words, d


def top_lists(d):
    """Un-pad beam_top_seq / beam_top_prob into per-image lists."""
    out_s, out_p = [], []
    for k, n in enumerate(d["beam_n_done"]):
        out_s.append(torch.from_numpy(d["beam_top_seq"][k, :n]))
        out_p.append([float(x) for x in d["beam_top_prob"][k, :n]])
    return out_s, out_p
