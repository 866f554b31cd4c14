"""The XE / RL training lines of bench.py on their own (one GPU): python profiles/experiments/r2_train_bench.py [xe|rl|both] [steps]"""
import json
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from recurrent_fusion_network_b200 import _capi  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
_capi.check(_capi.lib().rfn_check_device())
model = bench.build_model(dev)
model.fused_tape = bool(fused)


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _capi.lib().rfn_launch_count()
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, int(_capi.lib().rfn_launch_count() - l0), out


res = {"fused_tape": bool(fused)}
if which in ("xe", "both"):
    res["xe"] = bench.xe_train_bench(model, dev, 1, 0, steps, timed)
if which in ("rl", "both"):
    res["rl"] = bench.rl_train_bench(model, dev, 1, 0, steps, timed)
print(json.dumps(res))
