import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from tests.test_gpu_engines import _linear_engine
M, N, K = 15000, 9488, int(sys.argv[1]) if len(sys.argv) > 1 else 32
b = torch.zeros(N, device='cuda')
x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') * 0.1
for _ in range(3):
    _linear_engine(1, [x], [w], [b], M, N)
torch.cuda.synchronize()
