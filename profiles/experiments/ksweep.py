import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from tests.test_gpu_engines import _linear_engine
from recurrent_fusion_network_b200 import _capi
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M, N = 15000, 9488
b = torch.zeros(N, device='cuda')
for cl in (1, 0):
    _capi.lib().rfn_set_tc_cluster(cl)
    for K in (32, 128, 512, 1024, 2048):
        x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.1
        y = torch.empty(M, N, device='cuda')
        ms = t(lambda: _linear_engine(1, [x], [w], [b], M, N, accumulate_into=None))
        fl = 2.0 * M * N * K
        print(f"cluster={cl} K={K:5d}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
