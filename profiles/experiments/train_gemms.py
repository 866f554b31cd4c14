"""Two launches of the split-K 1-CTA tcgen05 kernel as the training tape uses it: (a) the stage-1 gates GEMM at 80 rows
(three K-major sources, weights streamed once), (b) dX = dY . W with the weights read as an MN-major B operand."""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
g = torch.Generator().manual_seed(0)
M, N, Ks = 80, 2048, [512, 2560, 2048]
xs = [torch.randn(M, k, generator=g).cuda() for k in Ks]
ws = [((torch.rand(N, k, generator=g) * 2 - 1) * 0.1).cuda() for k in Ks]
bs = [torch.randn(N, generator=g).cuda() for _ in Ks]
y = torch.empty(M, N, device='cuda')
ld = (C.c_int * 3)(*Ks); ks = (C.c_int * 3)(*Ks)
dY = torch.randn(M, 2048, generator=g).cuda(); W = ((torch.rand(2048, 5120, generator=g) * 2 - 1) * 0.1).cuda()
dX = torch.empty(M, 5120, device='cuda')
for _ in range(3):
    check(lib().rfn_linear_f32(3, ptr_array(xs), ld, ptr_array(ws), ks, ptr_array(bs), ptr(y), N, M, N, 2, stream()), "linear")
    check(lib().rfn_gemm_general_f32_engine(1, 1, 0, ptr(dY), 2048, ptr(W), 5120, ptr(dX), 5120, M, 5120, 2048, 0, stream()), "dx")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, nbytes in (("gates fwd 80 x 2048 x 5120", lambda: lib().rfn_linear_f32(3, ptr_array(xs), ld, ptr_array(ws), ks, ptr_array(bs), ptr(y), N, M, N, 2, stream()), N * 5120 * 4),
                         ("dX 80 x 5120 x 2048 (MN-major W)", lambda: lib().rfn_gemm_general_f32_engine(1, 1, 0, ptr(dY), 2048, ptr(W), 5120, ptr(dX), 5120, M, 5120, 2048, 0, stream()), 2048 * 5120 * 4)):
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms * 1e3:.1f} us per call (memset + kernel), weight stream {nbytes / 1e6:.1f} MB -> {nbytes / ms / 1e6:.0f} GB/s (L2-warm)")
