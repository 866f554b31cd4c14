"""Round 2: the split-fp16 engine (mode 4) against engine modes 3 / 1 / 2 on the shapes of the decode path.
Times ONLY the GEMM kernels (operands pre-split, as the path does for the features) with CUDA events."""
import ctypes as C, json, sys, torch
sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream

def split(t, rows, K, bf16):
    ks = (C.c_int * 1)(K)
    n = lib().rfn_split_bytes(rows, 1, ks, bf16)
    buf = torch.empty(n, dtype=torch.uint8, device='cuda')
    check(lib().rfn_split_rows_f32(1, ptr_array([t]), (C.c_int * 1)(t.stride(0)), ks, rows, bf16, ptr(buf), n, stream()))
    return buf

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
for name, M, N, K in (("att2att_resnet_1024img", 200704, 512, 2048), ("att2att_resnet_5000img", 980000, 512, 2048),
                      ("gates_5000", 5000, 2048, 4608), ("logits_15000", 15000, 9488, 512), ("g_5000", 5000, 512, 512)):
    x = torch.randn(M, K, device='cuda'); w = (torch.rand(N, K, device='cuda') * 2 - 1) * 0.1; b = torch.zeros(N, device='cuda')
    y = torch.empty(M, N, device='cuda')
    rec = dict(shape=name, M=M, N=N, K=K, gflop=2e-9 * M * N * K)
    for eng in (3, 1, 2):
        ms = timeit(lambda: check(lib().rfn_linear_f32_engine(eng, 1, ptr_array([x]), (C.c_int * 1)(K), ptr_array([w]), (C.c_int * 1)(K),
                                                              ptr_array([b]), ptr(y), N, M, N, 0, stream())))
        rec[f"engine{eng}_ms"] = round(ms, 4); rec[f"engine{eng}_tflops"] = round(rec["gflop"] / ms, 1)
    ref = y.clone()
    for bf16 in (0, 1):
        xs, ws = split(x, M, K, bf16), split(w, N, K, bf16)
        ks = (C.c_int * 1)(K)
        ms = timeit(lambda: check(lib().rfn_linear_split(bf16, 1, ptr(xs), ptr(ws), ks, ptr_array([b]), ptr(y), N, M, N, 0, stream())))
        tag = "bf16" if bf16 else "fp16x3"
        rec[f"{tag}_ms"] = round(ms, 4); rec[f"{tag}_tflops"] = round(rec["gflop"] / ms, 1)
        rec[f"{tag}_maxdiff_vs_tf32_single"] = float((y - ref).abs().max())
        ms = timeit(lambda: split(x, M, K, bf16), 3)
        rec[f"{tag}_split_x_ms"] = round(ms, 4)
        del xs, ws
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del x, w, y, ref
    torch.cuda.empty_cache()
json.dump(out, open('gpurun_out/r2_h3_perf.json', 'w'), indent=1)
