"""Worker for tests/test_gpu_ddp_nccl.py: rank r of a 2-process NCCL group computes the XE gradients of ITS half of the
config-2 batch (40 of the 80 rows) with the CUDA path, the ranks average them with dist.average_gradients, and rank 0 writes
the averaged gradients to a file.  Every loss term of the reference is normalised by the LOCAL row count
(misc/utils.py:72,177,184,188), so the mean of the rank gradients must equal the single-process gradient on all 80 rows."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rfnet_oracle as O  # noqa: E402
from recurrent_fusion_network_b200 import dist as D  # noqa: E402
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion  # noqa: E402
from tests._gpu_util import build_model  # noqa: E402
from tests.test_gpu_training import _config2_batch  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
cfg, fc, att, labels, masks, top = _config2_batch()
rows = labels.shape[0]
per = rows // world
sl = slice(rank * per, (rank + 1) * per)
sd = O.make_state_dict(cfg, seed=1234)
m = build_model(cfg, sd).train()
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
lp, rp = m([f[sl].cuda() for f in fc], [a[sl].cuda() for a in att], labels[sl].cuda())
loss = crit(lp, labels[sl, 1:].cuda(), masks[sl, 1:].cuda(), rp, top[sl].cuda(), 10.0)
loss.backward()
D.average_gradients(m.parameters())   # a generator on purpose (it used to be exhausted after the first pass)
loss_sum = loss.detach().clone()
dist.all_reduce(loss_sum)
torch.cuda.synchronize()
# the same step with the all-reduce overlapped inside the backward pass (dist.OverlappedGradSync: post-accumulate hooks plus
# the early hand-over of each fusion step's weight gradients from tape.Stage1Fn): SUM of the rank gradients = world x the mean
avg = {k: p.grad.clone() for k, p in m.named_parameters()}
m.zero_grad(set_to_none=True)
params = list(m.parameters())
gs = D.OverlappedGradSync(params, bucket_bytes=8 << 20)
gs.install()
lp, rp = m([f[sl].cuda() for f in fc], [a[sl].cuda() for a in att], labels[sl].cuda())
loss2 = crit(lp, labels[sl, 1:].cuda(), masks[sl, 1:].cuda(), rp, top[sl].cuda(), 10.0)
loss2.backward()
gs.finish()
gs.remove()
torch.cuda.synchronize()
worst = 0.0
for k, p in m.named_parameters():
    scale = float(avg[k].abs().max()) + 1e-6
    err = float((p.grad / world - avg[k]).abs().max())
    worst = max(worst, err / scale)
    assert err / scale <= 2e-5 or err <= 1e-6, f"overlapped sync, {k}: rel err {err / scale:.3g}"
assert gs.buckets_launched >= 8, gs.buckets_launched
print(f"OVERLAPPED_SYNC_OK rank {rank} buckets {gs.buckets_launched} worst rel err {worst:.2e}")
for k, p in m.named_parameters():
    p.grad = avg[k]
if rank == 0:
    torch.save({"loss_mean": float(loss_sum) / world, "grads": {k: p.grad.cpu() for k, p in m.named_parameters()}}, sys.argv[1])
dist.barrier()
dist.destroy_process_group()
print("DDP_OK rank", rank)
