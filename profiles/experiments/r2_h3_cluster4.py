"""4-CTA clusters of the split-fp16 / bf16 engine (two pairs sharing the W tile by TMA multicast) against the 2-CTA form:
bit-identical outputs expected (same MMA sequence per tile), then timing of the stage-1 att_2_att_h shape and a gates shape."""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream  # noqa: E402


def linear(engine, xs, ws, bs, M, N):
    n = len(xs)
    y = torch.empty(M, N, device="cuda")
    ld = (C.c_int * n)(*[x.stride(0) for x in xs])
    ks = (C.c_int * n)(*[x.shape[1] for x in xs])
    check(lib().rfn_linear_f32_engine(engine, n, ptr_array(xs), ld, ptr_array(ws), ks, ptr_array(bs), ptr(y), N, M, N, 0, stream()),
          "rfn_linear_f32_engine")
    return y


g = torch.Generator(device="cuda").manual_seed(0)
ok = True
for engine in (4, 5):
    for (M, N, Ks) in [(1000, 2048, [2560, 1280]), (520, 512, [2048]), (5000, 2048, [2560, 2048]), (300, 9488, [512]), (777, 256, [72]),
                       (4096, 512, [512, 512, 512])]:
        xs = [torch.randn(M, k, device="cuda", generator=g) for k in Ks]
        ws = [(torch.rand(N, k, device="cuda", generator=g) * 2 - 1) * 0.1 for k in Ks]
        bs = [torch.randn(N, device="cuda", generator=g) for _ in Ks]
        check(lib().rfn_set_h3_cluster(2))
        y2 = linear(engine, xs, ws, bs, M, N)
        check(lib().rfn_set_h3_cluster(4))
        y4 = linear(engine, xs, ws, bs, M, N)
        torch.cuda.synchronize()
        want = sum(x.double() @ w.double().t() + b.double() for x, w, b in zip(xs, ws, bs))
        d = float((y2 - y4).abs().max())
        e = float((y4.double() - want).abs().max() / want.abs().max())
        print(f"engine {engine} M={M} N={N} K={Ks}: |cl4 - cl2| = {d:.3g}, rel err vs fp64 {e:.3g}", flush=True)
        ok &= d == 0.0
print("CLUSTER4_IDENTICAL" if ok else "CLUSTER4_DIFFERS")
# timing: pre-split operands through rfn_linear_split
for bf16, name in ((0, "fp16x3"), (1, "bf16")):
    for (M, N, K) in [(1024 * 196, 512, 2048), (5000, 2048, 4608), (15000, 9488, 512)]:
        x = torch.randn(M, K, device="cuda", generator=g)
        w = (torch.rand(N, K, device="cuda", generator=g) * 2 - 1) * 0.1
        ks = (C.c_int * 1)(K)

        def split(t):
            nb = lib().rfn_split_bytes(t.shape[0], 1, ks, bf16)
            out = torch.empty(int(nb), dtype=torch.uint8, device="cuda")
            ld = (C.c_int * 1)(K)
            check(lib().rfn_split_rows_f32(1, ptr_array([t]), ld, ks, t.shape[0], bf16, ptr(out), out.numel(), stream()))
            return out
        xs, wsp = split(x), split(w)
        y = torch.empty(M, N, device="cuda")
        for cl in (2, 4):
            check(lib().rfn_set_h3_cluster(cl))
            for _ in range(3):
                check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(wsp), ks, ptr_array([None]), ptr(y), N, M, N, 0, stream()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(wsp), ks, ptr_array([None]), ptr(y), N, M, N, 0, stream()))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{name} M={M} N={N} K={K} cluster {cl}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        del x, w, xs, wsp, y
