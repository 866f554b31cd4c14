import ctypes as C, sys, torch
sys.path.insert(0, '.')  # run from the repo root
from tests.test_gpu_engines import _linear_engine
torch.manual_seed(0)
def tf32(x):
    return (x.view(torch.int32) & ~0x1fff).view(torch.float32)
for K in (256, 1024, 2048, 4096):
    M, N = 256, 256
    x = tf32(torch.randn(M, K)); w = tf32((torch.rand(N, K) * 2 - 1) * 0.1); b = torch.zeros(N)
    want = x.double() @ w.double().t()
    for eng in (0, 1, 2):
        got = _linear_engine(eng, [x.cuda()], [w.cuda()], [b.cuda()], M, N).double().cpu()
        err = (got - want)
        rel = err / want.abs().clamp_min(1e-3)
        print(f"K={K} engine={eng} tf32-exact inputs: max_abs={err.abs().max():.3g} rms={err.pow(2).mean().sqrt():.3g} "
              f"mean_signed(err*sign(want))={(err*want.sign()).mean():.3g} scale={want.abs().mean():.3g}")
    # positive-only inputs: accumulator grows monotonically -> RZ bias shows as a negative mean error
    xp, wp = x.abs(), w.abs()
    wantp = xp.double() @ wp.double().t()
    for eng in (0, 2):
        got = _linear_engine(eng, [xp.cuda()], [wp.cuda()], [b.cuda()], M, N).double().cpu()
        err = got - wantp
        print(f"K={K} engine={eng} positive inputs: mean_rel={(err/wantp).mean():.3g} max_rel={(err/wantp).abs().max():.3g}")
