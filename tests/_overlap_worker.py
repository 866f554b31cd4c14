"""Worker for tests/test_gpu_overlap_sync.py (one process, NCCL group of world size 1): GraphedXEStep with the gradient
all-reduce captured inside the backward graph (dist.OverlappedGradSync) against the same step without any collective."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rfnet_oracle as O  # noqa: E402
from recurrent_fusion_network_b200 import dist as D, training as TR  # noqa: E402
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion  # noqa: E402
from recurrent_fusion_network_b200.optim import FusedAdam  # noqa: E402
from tests._gpu_util import build_model, cuda_list  # noqa: E402

torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
cfg = O.tiny_config(2)
sd = O.make_state_dict(cfg, seed=11, init_range=0.5)
rows = 6
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
fc, att = O.make_inputs(cfg, rows, seed=1)
labels, masks, top = O.make_labels(cfg, rows, seed=1)
batch = (cuda_list(fc), cuda_list(att), labels.cuda(), masks.cuda().float(), top.cuda())
out = {}
for mode in ("plain", "overlapped"):
    m = build_model(cfg, sd)
    m.train()
    m.drop_prob_lm = m.decoder.drop_prob_lm = 0.0
    params = list(m.parameters())
    opt = FusedAdam(params, lr=1e-3, weight_decay=1e-5, grad_clip=1.0, capturable=True)
    gs = D.OverlappedGradSync(params, bucket_bytes=4096) if mode == "overlapped" else None
    step = TR.GraphedXEStep(m, crit, opt, *batch, 10.0, warmup=1, grad_sync=gs)
    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    out[mode] = (float(loss), {k: v.detach().clone() for k, v in m.state_dict().items()}, gs.buckets_launched if gs else 0)
assert out["overlapped"][2] >= 2, "the hooks must have launched several buckets during the capture"
assert abs(out["plain"][0] - out["overlapped"][0]) <= 1e-5 * max(1.0, abs(out["plain"][0]))
worst = max(float((out["plain"][1][k] - out["overlapped"][1][k]).abs().max()) for k in out["plain"][1]
            if not k.endswith("att_h_2_out.bias"))
assert worst <= 2e-5, worst
dist.destroy_process_group()
print("OVERLAP_OK buckets=%d worst=%.2e" % (out["overlapped"][2], worst))
