"""Host-side cost of one kernel launch on this box (explains the eager small-batch numbers: bench.py config1_latency, xe_train.eager)."""
import time, torch, ctypes as C, sys
sys.path.insert(0, '.')
from recurrent_fusion_network_b200._capi import lib, ptr, stream, check
x = torch.zeros(256, device='cuda'); y = torch.zeros(256, device='cuda')
for name, fn in (("torch x.add_(1)", lambda: x.add_(1)),
                 ("librfn rfn_axpby_f32", lambda: check(lib().rfn_axpby_f32(1.0, ptr(x), 0.0, ptr(y), ptr(y), 256, stream())))):
    for _ in range(200): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5000): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: {1e6 * (t1 - t0) / 5000:.2f} us per launch issued (host), {1e6 * (t2 - t0) / 5000:.2f} us per launch incl. drain")
