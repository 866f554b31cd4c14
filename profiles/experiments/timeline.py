import sys, torch
EPI = int(sys.argv[1]) if len(sys.argv) > 1 else -1
sys.path.insert(0, '.')  # run from the repo root
from tests.test_gpu_engines import _linear_engine
from recurrent_fusion_network_b200 import _capi
M, N = 15000, 9488
b = torch.zeros(N, device='cuda')
for K in (32, 512, 2048):
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.1
    nct = 2 * ((N + 255) // 256) * ((M + 255) // 256)
    buf = torch.zeros(nct, 8, dtype=torch.int64, device='cuda')
    _linear_engine(1, [x], [w], [b], M, N); torch.cuda.synchronize()
    _capi.lib().rfn_debug_set_timeline(buf.data_ptr(), EPI)
    _linear_engine(1, [x], [w], [b], M, N); torch.cuda.synchronize()
    _capi.lib().rfn_debug_set_timeline(None, -1)
    t = buf.cpu().double()
    lead = t[0::2]
    d = lambda a, b_: float((lead[:, b_] - lead[:, a]).median())
    print(f"K={K}: init {d(0,1):.0f}  init->firstMMA {d(1,2):.0f}  mainloop(first->last MMA issue) {d(2,3):.0f}  lastMMA->drained {d(3,4):.0f}  "
          f"epilogue {d(4,5):.0f}  teardown {d(5,6):.0f}  total {d(0,6):.0f} clk (median over leader CTAs)")
