import sys, torch, ctypes as C
sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import check, lib, ptr, ptr_array, stream
L = lib()
L.rfn_debug_set_mn_desc.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint]
def run(M, N, K, A, B, eng=2):
    y = torch.full((M, N), -7.0, device='cuda')
    check(L.rfn_gemm_general_f32_engine(eng, 1, 0, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(y), N, M, N, K, 0, stream()), "x")
    torch.cuda.synchronize()
    return y
M, N, K = 128, 128, 32
A = torch.zeros(M, K, device='cuda')
for i in range(32): A[i, i] = 1.0
B = (torch.arange(K, device='cuda').float()[:, None] * 1000 + torch.arange(N, device='cuda').float()[None, :]).contiguous()
want = A @ B
def desc(lbo, sbo, ver=1, layout=1):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (ver << 46) | (layout << 61)
for name, d, kstep, idb in [
    ("default lbo4096 sbo512 layout1", 0, 0, 0),
    ("lbo512 sbo4096", desc(512, 4096), 64, 1 << 16),
    ("lbo4096 sbo1024", desc(4096, 1024), 64, 1 << 16),
    ("lbo1024 sbo4096", desc(1024, 4096), 64, 1 << 16),
]:
    L.rfn_debug_set_mn_desc(d, kstep, idb)
    y = run(M, N, K, A, B)
    nz = int((y != 0).sum())
    print(f"{name}: nonzero {nz}, maxdiff {float((y - want).abs().max()):.1f}; y[0,:6]={y[0,:6].tolist()} y[1,:3]={y[1,:3].tolist()} y[9,33:36]={y[9,33:36].tolist()} y[5,:40:8]={y[5,:40:8].tolist()}")
L.rfn_debug_set_mn_desc(0, 0, 0)
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(80, 2560, 2048), (300, 516, 1028)]:
    dY = torch.randn(M, K, generator=g).cuda(); W = ((torch.rand(K, N, generator=g) * 2 - 1) * 0.1).cuda()
    want = dY.double() @ W.double()
    y = run(M, N, K, dY, W, eng=1)
    print(M, N, K, "3xTF32 rel err", float((y - want).abs().max() / want.abs().max()))
