"""Round 2: the split-fp16 / bf16 GEMM kernels alone (operands pre-split) on the shapes of the decode path, CUDA events.
Usage: python profiles/experiments/r2_h3_perf2.py <tag>  ->  gpurun_out/r2_h3_perf2_<tag>.json"""
import ctypes as C, json, sys, torch
sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream

def split(t, rows, K, bf16):
    ks = (C.c_int * 1)(K)
    n = lib().rfn_split_bytes(rows, 1, ks, bf16)
    buf = torch.empty(n, dtype=torch.uint8, device='cuda')
    check(lib().rfn_split_rows_f32(1, ptr_array([t]), (C.c_int * 1)(t.stride(0)), ks, rows, bf16, ptr(buf), n, stream()))
    return buf

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
out = []
for name, M, N, K in (("att2att_resnet_1024img", 200704, 512, 2048), ("att2att_resnet_5000img", 980000, 512, 2048),
                      ("gates_5000", 5000, 2048, 4608), ("logits_15000", 15000, 9488, 512), ("g_5000", 5000, 512, 512),
                      ("square_8192", 8192, 8192, 8192)):
    x = torch.randn(M, K, device='cuda'); w = (torch.rand(N, K, device='cuda') * 2 - 1) * 0.1; b = torch.zeros(N, device='cuda')
    y = torch.empty(M, N, device='cuda')
    rec = dict(shape=name, M=M, N=N, K=K, gflop=2e-9 * M * N * K)
    ref = None
    if M * N <= 80e6:
        ref = (x.double() @ w.double().t()).float()
    for bf16 in (0, 1):
        xs, ws = split(x, M, K, bf16), split(w, N, K, bf16)
        ks = (C.c_int * 1)(K)
        ms = timeit(lambda: check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(ws), ks, ptr_array([b]), ptr(y), N, M, N, 0, stream())))
        t = "bf16" if bf16 else "fp16x3"
        rec[f"{t}_ms"] = round(ms, 4); rec[f"{t}_tflops"] = round(rec["gflop"] / ms, 1)
        if ref is not None:
            rec[f"{t}_maxerr_vs_fp64"] = float((y - ref).abs().max())
        del xs, ws
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del x, w, y, ref
    torch.cuda.empty_cache()
json.dump(out, open(f'gpurun_out/r2_h3_perf2_{tag}.json', 'w'), indent=1)
