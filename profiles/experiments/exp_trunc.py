import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from tests.test_gpu_engines import _linear_engine
torch.manual_seed(0)
M, N, K = 128, 256, 32
x = torch.randn(M, K); w = torch.randn(N, K); b = torch.zeros(N)
def trunc(t): return (t.view(torch.int32) & ~0x1fff).view(torch.float32)
def rn(t):
    i = t.view(torch.int32)
    i = (i + 0xfff + ((i >> 13) & 1)) & ~0x1fff
    return i.view(torch.float32)
got = _linear_engine(2, [x.cuda()], [w.cuda()], [b.cuda()], M, N).double().cpu()
for name, fx in (("trunc", trunc), ("rn", rn)):
    want = fx(x).double() @ fx(w).double().t()
    print(name, "max abs diff", float((got - want).abs().max()), "rms", float((got-want).pow(2).mean().sqrt()))
print("vs exact fp32 inputs", float((got - x.double() @ w.double().t()).abs().max()))
