"""GPU: the tcgen05 GEMM engines (3xTF32 and single-pass TF32) against the fp32 SIMT engine and an
fp64 reference of the same contraction; the fused attention-score epilogue against the oracle."""
import ctypes as C

import pytest
import torch

from oracle import rfnet_oracle as O
from tests._gpu_util import build_model, cuda_list, maxdiff

pytestmark = pytest.mark.gpu


def _linear_engine(engine, xs, ws, bs, M, N, accumulate_into=None):
    from recurrent_fusion_network_b200 import _capi
    from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
    n = len(xs)
    y = accumulate_into if accumulate_into is not None else torch.empty(M, N, device="cuda")
    ld = (C.c_int * n)(*[x.stride(0) for x in xs])
    ks = (C.c_int * n)(*[x.shape[1] for x in xs])
    check(lib().rfn_linear_f32_engine(engine, n, ptr_array(xs), ld, ptr_array(ws), ks, ptr_array(bs), ptr(y), y.stride(0),
                                      M, N, 1 if accumulate_into is not None else 0, stream()), "rfn_linear_f32_engine")
    return y


SHAPES = [(128, 256, [32]), (128, 128, [64]), (256, 512, [2048]), (200, 512, [2208]), (1000, 2048, [2560, 1280]),
          (300, 9488, [512]), (129, 132, [36, 64, 128]), (64, 40, [24]), (5, 512, [512])]


@pytest.mark.parametrize("M,N,Ks", SHAPES)
@pytest.mark.parametrize("engine,tol", [(1, 3e-6), (2, 3e-3)])
def test_tc_engine_matches_fp64(engine, tol, M, N, Ks):
    g = torch.Generator().manual_seed(M + N)
    xs = [torch.randn(M, k, generator=g) for k in Ks]
    ws = [(torch.rand(N, k, generator=g) * 2 - 1) * 0.1 for k in Ks]
    bs = [torch.randn(N, generator=g) for _ in Ks]
    want = sum(x.double() @ w.double().t() + b.double() for x, w, b in zip(xs, ws, bs))
    got = _linear_engine(engine, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    err = maxdiff(got, want)
    assert err <= tol * scale, f"engine {engine}: err {err:.3g} vs scale {scale:.3g}"
    if engine == 1:
        ref = _linear_engine(0, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N)
        # 3xTF32 must be as good as the fp32 SIMT engine (both ~1e-6 relative to fp64)
        assert err <= 4 * maxdiff(ref, want) + 1e-6 * scale
        acc = _linear_engine(engine, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N, accumulate_into=ref.clone())
        assert maxdiff(acc, 2 * want) <= 2 * tol * scale


def test_strided_operands():
    """x with a leading dimension larger than K (the H-concat slices of stage 1)."""
    g = torch.Generator().manual_seed(1)
    big = torch.randn(300, 2560, generator=g).cuda()
    x = big[:, 512:1024]
    w = ((torch.rand(512, 512, generator=g) * 2 - 1) * 0.1).cuda()
    b = torch.randn(512, generator=g).cuda()
    want = x.double() @ w.double().t() + b.double()
    for engine, tol in ((0, 3e-6), (1, 3e-6)):
        got = _linear_engine(engine, [x], [w], [b], 300, 512)
        assert maxdiff(got, want) <= tol * float(want.abs().max())


@pytest.mark.parametrize("mode,tol", [(1, 2e-5), (3, 2e-5), (2, 5e-3)])
def test_stage1_with_fused_score_epilogue(mode, tol):
    """Full thought-vector pass with the tensor engine (fused tanh-score epilogue) vs the oracle."""
    from recurrent_fusion_network_b200 import _capi
    cfg = O.config1(196)
    sd = O.make_state_dict(cfg, seed=1234)
    fc, att = O.make_inputs(cfg, 6, seed=5)
    m = build_model(cfg, sd)
    _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
    try:
        with torch.no_grad():
            TVc, rp, st = m.get_thought_vectors(cuda_list(fc), cuda_list(att), m.get_init_state(cuda_list(fc)))
            torch.cuda.synchronize()
        TVc_o, rp_o, st_o = O.get_thought_vectors(sd, cfg, att, O.get_init_state(sd, cfg, fc))
        assert maxdiff(TVc, TVc_o) <= tol
        assert maxdiff(st[0][0], st_o[0]) <= tol
    finally:
        _capi.lib().rfn_set_gemm_mode(0)


def test_fused_vocab_epilogue_beam_matches_oracle_and_simt():
    """>= 128 decoder rows: the logits GEMM runs on the tensor engine with the fused
    log-softmax-statistics + top-k epilogue (logits never written); compare with the oracle (tiny vocab)
    and with the fp32 SIMT engine (9488-way vocab)."""
    from recurrent_fusion_network_b200 import _capi
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=1250, init_range=0.5, logit_scale=3.0, eos_bias=0.8)
    fc, att = O.make_inputs(cfg, 50, seed=4)
    m = build_model(cfg, sd)
    with torch.no_grad():
        seq, slp, ts, tp, _ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        o = O.sample_beam(sd, cfg, fc, att, beam_size=3)
    assert torch.equal(seq.cpu(), o[0]) and maxdiff(slp, o[1]) <= 1e-4
    assert [t.shape for t in ts] == [t.shape for t in o[2]]

    cfg = O.config1(49)
    sd = O.make_state_dict(cfg, seed=1234, sharpen=True)
    fc, att = O.make_inputs(cfg, 48, seed=6)
    m = build_model(cfg, sd)
    res = {}
    for mode in (1, 0):
        _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
        with torch.no_grad():
            res[mode] = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
    _capi.check(_capi.lib().rfn_set_gemm_mode(1))
    same = (res[0][0] == res[1][0]).all(dim=1)
    assert int((~same).sum()) <= 1, "tensor-engine and SIMT captions differ on more than a near-tie"
    assert maxdiff(res[0][1][same], res[1][1][same]) <= 1e-4


@pytest.fixture(params=[1, 2], ids=["cluster", "persistent"])
def tc_cluster(request):
    from recurrent_fusion_network_b200 import _capi
    prev = _capi.lib().rfn_get_tc_cluster()
    _capi.check(_capi.lib().rfn_set_tc_cluster(request.param))
    yield
    _capi.check(_capi.lib().rfn_set_tc_cluster(prev))


@pytest.mark.parametrize("M,N,Ks", [(256, 256, [32]), (256, 512, [2048]), (1000, 2048, [2560, 1280]), (300, 9488, [512]),
                                    (777, 260, [36, 64, 128]), (20000, 768, [96])])
@pytest.mark.parametrize("engine,tol", [(1, 3e-6), (3, 4e-6), (2, 3e-3)])
def test_two_cta_cluster_engine(tc_cluster, engine, tol, M, N, Ks):
    """cta_group::2 pairs (256 x 256 tiles, operands split across the two CTAs)."""
    g = torch.Generator().manual_seed(M + N)
    xs = [torch.randn(M, k, generator=g) for k in Ks]
    ws = [(torch.rand(N, k, generator=g) * 2 - 1) * 0.1 for k in Ks]
    bs = [torch.randn(N, generator=g) for _ in Ks]
    want = sum(x.double() @ w.double().t() + b.double() for x, w, b in zip(xs, ws, bs))
    got = _linear_engine(engine, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    assert maxdiff(got, want) <= tol * scale


def test_library_services_profile_concurrency_and_errors():
    """rfn_profile_*, rfn_set_concurrency, rfn_launch_count and the error path of a too-small workspace."""
    import ctypes as C
    from recurrent_fusion_network_b200 import _capi
    from recurrent_fusion_network_b200._capi import lib, ptr, ptr_array, stream
    cfg = O.tiny_config(2)
    sd = O.make_state_dict(cfg, seed=3, init_range=0.5)
    fc, att = O.make_inputs(cfg, 5, seed=1)
    m = build_model(cfg, sd)
    outs = {}
    _capi.check(lib().rfn_set_persistent_decoder(0))   # the per-step launch path is what carries the profile classes checked here
    for conc in (0, 1):
        _capi.check(lib().rfn_set_concurrency(conc))
        n0 = lib().rfn_launch_count()
        _capi.profile_enable(True)
        with torch.no_grad():
            outs[conc] = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        torch.cuda.synchronize()
        prof = _capi.profile_read()
        _capi.profile_enable(False)
        n_prof = sum(v[1] for v in prof.values())
        assert lib().rfn_launch_count() - n0 >= n_prof > 50   # a profiled launcher may issue several kernels
        assert prof["beam_merge"][1] == cfg.seq_length + 1 and prof["lstm_cell"][0] > 0
    _capi.check(lib().rfn_set_concurrency(1))
    _capi.check(lib().rfn_set_persistent_decoder(1))
    assert torch.equal(outs[0][0], outs[1][0]) and maxdiff(outs[0][1], outs[1][1]) == 0.0   # same kernels, same bits
    # workspace too small -> RFN_ERR_WORKSPACE with a message, nothing launched
    TVc = torch.empty(5, cfg.num_review_steps, cfg.rnn_size, device="cuda")
    h = torch.empty(5, cfg.rnn_size, device="cuda")
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    rc = lib().rfn_thought_vectors(C.byref(m._dims), m._params(), ptr_array(cuda_list(fc)), None, None, ptr_array(cuda_list(att)), 5,
                                   ptr(TVc), ptr(h), ptr(h.clone()), None, None, ptr(ws), ws.numel(), stream())
    assert rc == -3 and b"workspace" in lib().rfn_last_error()


@pytest.mark.parametrize("M,N,K", [(512, 2048, 15680), (512, 1536, 3136), (256, 260, 1028), (512, 2208, 784)])
def test_split_k_weight_gradient_gemm(M, N, K):
    """RFN_GEMM_SPLITK: long contraction, few output tiles (dU = dP^T . A) split over clusters, partial tiles added
    atomically; with and without accumulation into y."""
    g = torch.Generator().manual_seed(K)
    x = torch.randn(M, K, generator=g).cuda()
    w = ((torch.rand(N, K, generator=g) * 2 - 1) * 0.1).cuda()
    want = x.double() @ w.double().t()
    scale = float(want.abs().max())
    y = _linear_engine_flags(1, x, w, M, N, 2)
    assert maxdiff(y, want) <= 3e-6 * scale
    y0 = torch.randn(M, N, generator=g).cuda()
    y = _linear_engine_flags(1, x, w, M, N, 3, y0.clone())
    assert maxdiff(y, want + y0.double()) <= 3e-6 * scale


def _linear_engine_flags(engine, x, w, M, N, flags, y=None):
    from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
    y = torch.empty(M, N, device="cuda") if y is None else y
    ld = (C.c_int * 1)(x.stride(0))
    ks = (C.c_int * 1)(x.shape[1])
    check(lib().rfn_linear_f32_engine(engine, 1, ptr_array([x]), ld, ptr_array([w]), ks, ptr_array([None]), ptr(y), N, M, N,
                                      flags, stream()), "rfn_linear_f32_engine")
    return y


def test_transpose_with_zero_padding():
    from recurrent_fusion_network_b200.autograd import _transpose
    g = torch.Generator().manual_seed(3)
    for rows, cols in [(5, 7), (1000, 130), (33, 64), (3137, 512)]:
        big = torch.randn(rows, cols + 3, generator=g).cuda()
        x = big[:, :cols]
        t = _transpose(x)
        assert t.shape == (cols, (rows + 3) // 4 * 4)
        assert torch.equal(t[:, :rows], x.t()) and float(t[:, rows:].abs().sum()) == 0.0


def test_attention_backward_tensor_engine_weight_gradient():
    """AttentionFn.backward: dU = dP^T . A through transposes + the split-K tensor GEMM == the fp32 SIMT general GEMM."""
    from recurrent_fusion_network_b200 import _capi, autograd as AG
    g = torch.Generator().manual_seed(9)
    rows, N, D, R, Ah = 12, 196, 512, 256, 256
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.1).cuda().requires_grad_(True)
    h, U_w, U_b, Wh_w, Wh_b, v_w, v_b = mk(rows, R), mk(Ah, D), mk(Ah), mk(Ah, R), mk(Ah), mk(1, Ah), mk(1)
    A = torch.randn(rows, N, D, generator=g).cuda()
    dz = torch.randn(rows, D, generator=g).cuda()
    grads = {}
    for mode in (1, 0):
        _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
        AG.clear_transposed_cache()
        z = AG.AttentionFn.apply(h, A, U_w, U_b, Wh_w, Wh_b, v_w, v_b)
        grads[mode] = torch.autograd.grad(z, [h, U_w, U_b, Wh_w], dz)
    _capi.check(_capi.lib().rfn_set_gemm_mode(1))
    for a, b in zip(grads[1], grads[0]):
        assert maxdiff(a, b) <= 2e-5 * max(1.0, float(b.abs().max()))
    assert float(grads[1][1].abs().max()) > 0


@pytest.mark.parametrize("M,N,Ks", [(80, 2048, [512, 2560, 2048]), (16, 2048, [2048]), (80, 9488, [512]), (127, 516, [2052]),
                                    (1, 1024, [1024])])
def test_small_row_split_k_linear(M, N, Ks):
    """rfn_linear_f32 with RFN_GEMM_SPLITK and < 128 rows: 1-CTA tensor kernel, contraction split over blockIdx.z."""
    from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
    g = torch.Generator().manual_seed(M + N)
    xs = cuda_list([torch.randn(M, k, generator=g) for k in Ks])
    ws = cuda_list([(torch.rand(N, k, generator=g) * 2 - 1) * 0.1 for k in Ks])
    bs = cuda_list([torch.randn(N, generator=g) for _ in Ks])
    want = sum(x.double() @ w.double().t() + b.double() for x, w, b in zip(xs, ws, bs))
    scale = float(want.abs().max())
    n = len(Ks)
    ld = (C.c_int * n)(*[x.stride(0) for x in xs])
    ks = (C.c_int * n)(*Ks)
    for acc in (0, 1):
        y0 = torch.randn(M, N, generator=g).cuda()
        y = y0.clone()
        check(lib().rfn_linear_f32(n, ptr_array(xs), ld, ptr_array(ws), ks, ptr_array(bs), ptr(y), N, M, N, acc | 2, stream()),
              "rfn_linear_f32")
        assert maxdiff(y, want + (y0.double() if acc else 0)) <= 3e-6 * scale


@pytest.mark.parametrize("M,N,K", [(80, 2560, 2048), (80, 2048, 9488), (16, 2208, 2048), (300, 516, 1028), (80, 132, 4096)])
def test_general_gemm_dx_on_tensor_engine(M, N, K):
    """rfn_gemm_general_f32(1, 0): dX[M,N] = dY[M,K] . W[K,N] with W row-major (K, N) -- the MN-major B operand path --
    against fp64 and against the SIMT kernel."""
    from recurrent_fusion_network_b200 import _capi
    from recurrent_fusion_network_b200.autograd import _gemm_general
    g = torch.Generator().manual_seed(M + N + K)
    dY = torch.randn(M, K, generator=g).cuda()
    W = ((torch.rand(K, N, generator=g) * 2 - 1) * 0.1).cuda()
    want = dY.double() @ W.double()
    scale = float(want.abs().max())
    out = {}
    for mode in (1, 0):
        _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
        for acc in (0, 1):
            y0 = torch.randn(M, N, generator=g).cuda()
            y = y0.clone()
            _gemm_general(1, 0, dY, K, W, N, y, N, M, N, K, accumulate=acc)
            assert maxdiff(y, want + (y0.double() if acc else 0)) <= 3e-6 * scale, (mode, acc)
    _capi.check(_capi.lib().rfn_set_gemm_mode(1))


# ---- split-fp16 ("fp16x3", engine 4) and single-pass bf16 (engine 5) ------------------------------------------------
H3_SHAPES = [(256, 256, [64]), (256, 512, [2048]), (1000, 2048, [2560, 1280]), (300, 9488, [512]), (777, 260, [36, 64, 128]),
             (20000, 768, [96]), (5000, 512, [2208]), (513, 516, [100])]


@pytest.mark.parametrize("M,N,Ks", H3_SHAPES)
@pytest.mark.parametrize("engine,tol", [(4, 3e-6), (5, 2e-2)])
def test_split_fp16_engine_matches_fp64(engine, tol, M, N, Ks):
    """rfn_linear_f32_engine(4 | 5): operands split into scaled fp16 pairs (or bf16), x0.w0 + x1.w0 + x0.w1 on kind::f16 MMAs."""
    g = torch.Generator().manual_seed(M + N)
    xs = [torch.randn(M, k, generator=g) for k in Ks]
    ws = [(torch.rand(N, k, generator=g) * 2 - 1) * 0.1 for k in Ks]
    bs = [torch.randn(N, generator=g) for _ in Ks]
    want = sum(x.double() @ w.double().t() + b.double() for x, w, b in zip(xs, ws, bs))
    got = _linear_engine(engine, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    err = maxdiff(got, want)
    assert err <= tol * scale, f"engine {engine}: err {err:.3g} vs scale {scale:.3g}"
    if engine == 4:
        ref = _linear_engine(0, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N)
        assert err <= 4 * maxdiff(ref, want) + 1e-6 * scale          # as good as the fp32 FFMA engine
        acc = _linear_engine(engine, cuda_list(xs), cuda_list(ws), cuda_list(bs), M, N, accumulate_into=ref.clone())
        assert maxdiff(acc, 2 * want) <= 2 * tol * scale


def test_split_fp16_engine_row_scaling():
    """Rows of x and of W whose magnitudes span 2^-40 .. 2^+40 (and all-zero rows): the power-of-two row scales keep every
    output row at fp32-grade accuracy RELATIVE TO ITS OWN SCALE -- an unscaled fp16 split would flush the small rows."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 512, 384, 520
    x = torch.randn(M, K, generator=g) * torch.pow(2.0, torch.randint(-40, 41, (M, 1), generator=g).float())
    w = torch.randn(N, K, generator=g) * torch.pow(2.0, torch.randint(-30, 31, (N, 1), generator=g).float())
    x[7] = 0
    w[11] = 0
    b = torch.zeros(N)
    want = x.double() @ w.double().t()
    got = _linear_engine(4, [x.cuda()], [w.cuda()], [b.cuda()], M, N).double().cpu()
    ref = (x.abs().double() @ w.abs().double().t()) + 1e-300        # per-element scale of the dot product
    rel = ((got - want).abs() / ref).max()
    assert float(rel) <= 1e-6, float(rel)
    assert float(got[7].abs().max()) == 0.0 and float(got[:, 11].abs().max()) == 0.0


def test_split_operator_api_and_bf16():
    """rfn_split_rows_f32 + rfn_linear_split (the pre-split operator API the path uses: features split once, reused by the
    eight fusion steps); pieces reconstruct x to 2^-22 relative to the row maximum; bf16 flavour within bf16 rounding."""
    from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
    g = torch.Generator().manual_seed(2)
    M, N, K = 640, 512, 264
    x = torch.randn(M, K, generator=g).cuda() * 3
    w = ((torch.rand(N, K, generator=g) * 2 - 1) * 0.1).cuda()
    bias = torch.randn(N, generator=g).cuda()
    want = x.double() @ w.double().t() + bias.double()
    for bf16, tol in ((0, 3e-6), (1, 2e-2)):
        ks = (C.c_int * 1)(K)
        bufs = []
        for t, rows in ((x, M), (w, N)):
            nbytes = lib().rfn_split_bytes(rows, 1, ks, bf16)
            buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            ld = (C.c_int * 1)(t.stride(0))
            check(lib().rfn_split_rows_f32(1, ptr_array([t]), ld, ks, rows, bf16, ptr(buf), nbytes, stream()), "rfn_split_rows_f32")
            bufs.append(buf)
        y = torch.empty(M, N, device="cuda")
        check(lib().rfn_linear_split(bf16, 1, ptr(bufs[0]), ptr(bufs[1]), ks, ptr_array([bias]), ptr(y), N, M, N, 0, stream()),
              "rfn_linear_split")
        assert maxdiff(y, want) <= tol * float(want.abs().max())
        if not bf16:
            inv = bufs[0][:M * 4].view(torch.float32)
            off = (M * 4 + 255) // 256 * 256
            sz = (M * K * 2 + 255) // 256 * 256
            p0 = bufs[0][off:off + M * K * 2].view(torch.float16).view(M, K).double()
            p1 = bufs[0][off + sz:off + sz + M * K * 2].view(torch.float16).view(M, K).double()
            rec = (p0 + p1) * inv.double().unsqueeze(1)
            rowmax = x.abs().max(dim=1, keepdim=True)[0].double()
            assert float(((rec - x.double()).abs() / rowmax).max()) <= 2.0 ** -21
            assert float((p0.abs().max(dim=1)[0]).min()) >= 2.0 ** 14 - 16 and float(p0.abs().max()) <= 2.0 ** 15


@pytest.mark.parametrize("engine", [4, 5])
def test_h3_four_cta_clusters_are_bit_identical(engine):
    """rfn_set_h3_cluster(4): two CTA pairs per cluster share the W tile by TMA multicast (each CTA loads half of its W rows for
    both pairs).  Same MMA sequence per tile, so the output must equal the 2-CTA form bit for bit -- including a last cluster
    step whose second pair lies entirely beyond M.  (Measured: no faster, profiles/r2_h3_cluster4_multicast.log; default 2.)"""
    from recurrent_fusion_network_b200._capi import check, lib
    g = torch.Generator().manual_seed(5)
    try:
        for M, N, Ks in [(1000, 2048, [2560, 1280]), (520, 512, [2048]), (777, 256, [72]), (1300, 768, [512, 512, 512])]:
            xs = cuda_list([torch.randn(M, k, generator=g) for k in Ks])
            ws = cuda_list([(torch.rand(N, k, generator=g) * 2 - 1) * 0.1 for k in Ks])
            bs = cuda_list([torch.randn(N, generator=g) for _ in Ks])
            check(lib().rfn_set_h3_cluster(2))
            y2 = _linear_engine(engine, xs, ws, bs, M, N)
            check(lib().rfn_set_h3_cluster(4))
            y4 = _linear_engine(engine, xs, ws, bs, M, N)
            torch.cuda.synchronize()
            assert torch.equal(y2, y4), (engine, M, N, Ks, maxdiff(y2, y4))
    finally:
        check(lib().rfn_set_h3_cluster(2))
