"""ctypes binding of librfn_b200.so (the C ABI declared in include/rfn_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this module
raises.  PyTorch is used only for device memory and streams; every tensor crosses the boundary as a
raw device pointer."""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librfn_b200.so")

MAX_ENCODERS = 8
MAX_BEAM = 16


class RfnDims(C.Structure):
    """struct rfn_dims -- the opt fields RecurrentFusionModel reads (misc/RecurrentFusionModel.py:120-151)."""
    _fields_ = [
        ("J", C.c_int32),
        ("att_num", C.c_int32 * MAX_ENCODERS),
        ("att_feat_size", C.c_int32 * MAX_ENCODERS),
        ("fc_feat_size", C.c_int32 * MAX_ENCODERS),
        ("rnn_size", C.c_int32),
        ("att_hid_size", C.c_int32),
        ("input_encoding_size", C.c_int32),
        ("vocab_plus1", C.c_int32),
        ("top_words_count", C.c_int32),
        ("num_review_steps_0", C.c_int32),
        ("num_review_steps", C.c_int32),
        ("seq_length", C.c_int32),
        ("review_maxout", C.c_int32),
        ("decoder_maxout", C.c_int32),
    ]


_vp = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t
_pp = C.POINTER(C.c_void_p)
_dims = C.POINTER(RfnDims)

_SIGNATURES = {
    "rfn_last_error": (C.c_char_p, []),
    "rfn_version": (_i, []),
    "rfn_check_device": (_i, []),
    "rfn_num_params": (_i, [_dims]),
    "rfn_num_param_slots": (_i, [_dims]),
    "rfn_wcache_bytes": (_sz, [_dims, _i]),
    "rfn_wcache_build": (_i, [_dims, _pp, _i, _vp, _sz, _vp]),
    "rfn_launch_count": (C.c_uint64, []),
    "rfn_engine_num": (_i, []),
    "rfn_engine_name": (C.c_char_p, [_i]),
    "rfn_engine_launch_counts": (_i, [C.POINTER(C.c_uint64), _i]),
    "rfn_set_gemm_mode": (_i, [_i]),
    "rfn_get_gemm_mode": (_i, []),
    "rfn_set_splitk": (_i, [_i]),
    "rfn_get_splitk": (_i, []),
    "rfn_set_tc_cluster": (_i, [_i]),
    "rfn_get_tc_cluster": (_i, []),
    "rfn_set_h3_cluster": (_i, [_i]),
    "rfn_get_h3_cluster": (_i, []),
    "rfn_set_att_bf16_variant": (_i, [_i]),
    "rfn_get_att_bf16_variant": (_i, []),
    "rfn_debug_set_timeline": (_i, [_vp, _i]),
    "rfn_set_concurrency": (_i, [_i]),
    "rfn_set_pdl": (_i, [_i]),
    "rfn_get_pdl": (_i, []),
    "rfn_set_persistent_decoder": (_i, [_i]),
    "rfn_get_persistent_decoder": (_i, []),
    "rfn_debug_set_pd_timeline": (_i, [_vp]),
    "rfn_profile_enable": (_i, [_i]),
    "rfn_profile_num_tags": (_i, []),
    "rfn_profile_tag_name": (C.c_char_p, [_i]),
    "rfn_profile_read": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_uint64), _i]),
    "rfn_linear_f32": (_i, [_i, _pp, C.POINTER(_i), _pp, C.POINTER(_i), _pp, _vp, _i, _i, _i, _i, _vp]),
    "rfn_linear_f32_engine": (_i, [_i, _i, _pp, C.POINTER(_i), _pp, C.POINTER(_i), _pp, _vp, _i, _i, _i, _i, _vp]),
    "rfn_split_bytes": (_sz, [_i, _i, C.POINTER(_i), _i]),
    "rfn_split_rows_f32": (_i, [_i, _pp, C.POINTER(_i), C.POINTER(_i), _i, _i, _vp, _sz, _vp]),
    "rfn_linear_split": (_i, [_i, _i, _vp, _vp, C.POINTER(_i), _pp, _vp, _i, _i, _i, _i, _vp]),
    "rfn_attention_step_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "rfn_attention_core_f32": (_i, [_vp] * 10 + [_i] * 5 + [_vp, _sz, _vp]),
    "rfn_lstm_cell_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "rfn_log_softmax_f32": (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "rfn_workspace_bytes": (_sz, [_dims, _i, _i]),
    "rfn_ensemble_workspace_bytes": (_sz, [_dims, _i, _i, _i]),
    "rfn_thought_vectors": (_i, [_dims, _pp, _pp, _pp, _pp, _pp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rfn_one_time_step": (_i, [_dims, _pp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "rfn_decode_teacher_forced": (_i, [_dims, _pp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "rfn_decode_sample": (_i, [_dims, _pp, _vp, _vp, _vp, _i, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rfn_decode_beam": (_i, [_dims, _pp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rfn_ensemble_decode_beam": (_i, [_dims, _i, C.POINTER(_pp), _pp, _pp, _pp, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _sz, _vp]),
    "rfn_ensemble_decode_greedy": (_i, [_dims, _i, C.POINTER(_pp), _pp, _pp, _pp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rfn_xe_loss_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "rfn_multilabel_margin_f32": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp]),
    "rfn_mean_log_softmax_f32": (_i, [_i, _pp, _i, _i, _vp, _vp, _vp]),
    "rfn_gemm_general_f32": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "rfn_gemm_general_f32_engine": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "rfn_colsum_f32": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "rfn_attention_step_bwd_f32": (_i, [_vp] * 6 + [_i] + [_vp] * 5 + [_i] * 5 + [_vp]),
    "rfn_lstm_cell_bwd_f32": (_i, [_vp] * 6 + [_i, _i, _vp]),
    "rfn_log_softmax_bwd_f32": (_i, [_vp, _sz, _vp, _sz, _vp, _sz, _i, _i, _vp]),
    "rfn_embed_f32": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "rfn_embed_bwd_f32": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "rfn_max_over_steps_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rfn_max_over_steps_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rfn_axpby_f32": (_i, [_f, _vp, _f, _vp, _vp, _sz, _vp]),
    "rfn_mul_scale_f32": (_i, [_f, _vp, _vp, _vp, _sz, _vp]),
    "rfn_select_token_f32": (_i, [_vp, _sz, _i, _i, _vp, _f, _vp, _vp, _vp]),
    "rfn_gather_cols_f32": (_i, [_vp, _sz, _vp, _vp, _i, _vp]),
    "rfn_scatter_cols_f32": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "rfn_expand_rows_f32": (_i, [_vp, _i, _vp, _i, _i, _vp]),
    "rfn_group_sum_f32": (_i, [_vp, _i, _vp, _i, _i, _vp]),
    "rfn_xe_loss_bwd_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "rfn_rl_loss_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "rfn_multilabel_margin_bwd_f32": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "rfn_ciderd_scores_f64": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, C.c_double, C.c_double, _vp, _vp, _vp]),
    "rfn_ciderd_reward_f32": (_i, [_vp, _i, _i, C.c_double, _i, _vp, _vp]),
    "rfn_ciderd_hash": (C.c_uint64, [C.POINTER(C.c_int32)]),
    "rfn_transpose_f32": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "rfn_adam_step_f32": (_i, [_i, _pp, _pp, _pp, _pp, C.POINTER(C.c_int64), _f, _f, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "rfn_rl_loss_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "rfn_lstm_cell_drop_f32": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "rfn_lstm_cell_ex_f32": (_i, [_vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "rfn_lstm_cell_bwd_ex_f32": (_i, [_vp, _vp, _i, _pp, C.POINTER(_i), _vp, _f, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "rfn_sum_strided_f32": (_i, [_i, _pp, C.POINTER(_i), _f, _vp, _i, _i, _i, _vp]),
    "rfn_mean_tensors_f32": (_i, [_vp, _sz, _i, _vp, _i, _i, _i, _i, _vp]),
    "rfn_xe_loss_strided_f32": (_i, [_vp, _sz, _sz, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "rfn_rl_loss_strided_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _i, _i, _i, _f, _vp, _vp]),
    "rfn_lstm_cell_bwd_multi_f32": (_i, [_vp, _vp, _i, _pp, C.POINTER(_i), _vp, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "rfn_linear_bwd_x_f32": (_i, [_i, _pp, C.POINTER(_i), _pp, C.POINTER(_i), C.POINTER(_i), _vp, _i, _i, _i, _i, _vp]),
    "rfn_xe_loss_bwd_strided_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _sz, _sz, _vp]),
    "rfn_rl_loss_bwd_strided_f32": (_i, [_vp, _vp, _vp, _sz, _sz, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
}

_lib = None
#: bumped whenever this package modifies parameters behind autograd's back (the fused optimizer kernel, CUDA-graph replays
#: of a training step): derived weight caches (model._params) are keyed on it
WEIGHTS_EPOCH = [0]


class RfnError(RuntimeError):
    pass


def exported_symbols():
    """Every symbol include/rfn_b200.h declares (the CPU test-suite checks they all resolve)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RfnError(
                f"{LIB_PATH} is missing: build it with `python -m recurrent_fusion_network_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
        mode = os.environ.get("RFN_GEMM_MODE")
        if mode is not None:
            check(_lib.rfn_set_gemm_mode(int(mode)), "rfn_set_gemm_mode")
        pdl = os.environ.get("RFN_PDL")
        if pdl is not None:
            check(_lib.rfn_set_pdl(int(pdl)), "rfn_set_pdl")
        h3c = os.environ.get("RFN_H3_CLUSTER")
        if h3c is not None:
            check(_lib.rfn_set_h3_cluster(int(h3c)), "rfn_set_h3_cluster")
        aw = os.environ.get("RFN_ATT_BF16_WIDE")
        if aw is not None:
            check(_lib.rfn_set_att_bf16_variant(int(aw)), "rfn_set_att_bf16_variant")
        cl = os.environ.get("RFN_TC_CLUSTER")
        if cl is not None:
            check(_lib.rfn_set_tc_cluster(int(cl)), "rfn_set_tc_cluster")
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().rfn_last_error().decode("utf-8", "replace")
        raise RfnError(f"{what or 'librfn_b200'} failed with status {status}: {msg}")


def ptr(t) -> int:
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "librfn_b200 takes device tensors only"
    return t.data_ptr()


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
    return arr


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def make_dims(encoders, rnn_size, att_hid_size, input_encoding_size, vocab_plus1, top_words_count,
              num_review_steps_0, num_review_steps, seq_length, review_maxout=0, decoder_maxout=0) -> RfnDims:
    """encoders: sequence of (att_num, att_feat_size, fc_feat_size)."""
    if len(encoders) > MAX_ENCODERS:
        raise RfnError(f"at most {MAX_ENCODERS} encoders are supported")
    d = RfnDims()
    d.J = len(encoders)
    for j, (n, a, f) in enumerate(encoders):
        d.att_num[j], d.att_feat_size[j], d.fc_feat_size[j] = n, a, f
    d.rnn_size, d.att_hid_size, d.input_encoding_size = rnn_size, att_hid_size, input_encoding_size
    d.vocab_plus1, d.top_words_count = vocab_plus1, top_words_count
    d.num_review_steps_0, d.num_review_steps, d.seq_length = num_review_steps_0, num_review_steps, seq_length
    d.review_maxout, d.decoder_maxout = (1 if review_maxout else 0), (1 if decoder_maxout else 0)
    return d


def engine_launch_counts():
    """-> {GEMM kernel family: launches since the process started}."""
    n = lib().rfn_engine_num()
    cnt = (C.c_uint64 * n)()
    check(lib().rfn_engine_launch_counts(cnt, n), "rfn_engine_launch_counts")
    return {lib().rfn_engine_name(i).decode(): int(cnt[i]) for i in range(n)}


def profile_enable(on: bool) -> None:
    check(lib().rfn_profile_enable(1 if on else 0))


def profile_read():
    """-> {class name: (milliseconds, launches)} accumulated since the last read."""
    n = lib().rfn_profile_num_tags()
    ms = (C.c_float * n)()
    cnt = (C.c_uint64 * n)()
    check(lib().rfn_profile_read(ms, cnt, n), "rfn_profile_read")
    return {lib().rfn_profile_tag_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
