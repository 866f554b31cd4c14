"""CPU: the C-ABI library builds/loads and exports every symbol include/rfn_b200.h declares; the
Python boundary mirrors the reference's state_dict layout; the product path refuses to run without
CUDA (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

from oracle import rfnet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from recurrent_fusion_network_b200 import build
    build.build()
    from recurrent_fusion_network_b200 import _capi
    return _capi


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "rfn_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rfn_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    dll = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in rfn_b200.h but not exported"
    assert declared == capi.exported_symbols(), "ctypes table and header disagree"
    assert capi.lib().rfn_version() >= 100


def test_num_params_matches_state_dict(capi):
    from tests._gpu_util import opt_from_cfg
    from recurrent_fusion_network_b200 import RecurrentFusionModel
    for cfg in (O.tiny_config(1), O.tiny_config(3)):
        m = RecurrentFusionModel(opt_from_cfg(cfg))
        shapes = O.state_dict_shapes(cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
        assert capi.lib().rfn_num_params(ctypes.byref(m._dims)) == len(shapes)
    full = capi.make_dims([(e.att_num, e.att_feat_size, e.fc_feat_size) for e in O.FULL_ENCODERS], 512, 512, 512, 9488,
                          1000, 8, 8, 16)
    assert capi.lib().rfn_num_params(ctypes.byref(full)) == 773
    assert capi.lib().rfn_workspace_bytes(ctypes.byref(full), 16, 48) > 0


def test_no_cpu_fallback(capi):
    from tests._gpu_util import opt_from_cfg
    from recurrent_fusion_network_b200 import RecurrentFusionModel
    cfg = O.tiny_config(2)
    m = RecurrentFusionModel(opt_from_cfg(cfg)).eval()
    fc, att = O.make_inputs(cfg, 2)
    with torch.no_grad(), pytest.raises(capi.RfnError):
        m.sample(fc, att, {"beam_size": 3})
    with torch.no_grad(), pytest.raises(capi.RfnError):
        m.get_init_state(fc)


def test_argument_validation_without_gpu(capi):
    # bad dims are rejected before any CUDA call
    bad = capi.make_dims([(5, 18, 16)], 32, 16, 24, 60, 20, 3, 2, 6)  # att_feat_size not a multiple of 4
    assert capi.lib().rfn_workspace_bytes(ctypes.byref(bad), 4, 4) == 0
    assert b"multiples of 4" in capi.lib().rfn_last_error()
    assert capi.lib().rfn_set_gemm_mode(7) != 0
