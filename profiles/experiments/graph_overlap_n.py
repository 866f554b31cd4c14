"""torchrun --nproc-per-node N: GraphedXEStep with the gradient all-reduce (a) between the two graphs and (b) captured
inside the backward graph (OverlappedGradSync); parameters after 3 steps must agree between the two, step time printed."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from types import SimpleNamespace
from recurrent_fusion_network_b200 import dist as D, training as TR
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
from recurrent_fusion_network_b200.optim import FusedAdam
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
imgs, spi, L = 16, 5, 16
rows = imgs * spi
fc = [torch.randn(imgs, f, device=dev, generator=g).repeat_interleave(spi, 0) for (_, _, f) in bench.ENC]
att = [torch.randn(imgs, n, d, device=dev, generator=g).repeat_interleave(spi, 0) for (n, d, _) in bench.ENC]
cg = torch.Generator().manual_seed(200 + rank)
labels = torch.zeros(rows, L + 2, dtype=torch.int64); masks = torch.zeros(rows, L + 2)
for b in range(rows):
    n = int(torch.randint(5, L + 1, (1,), generator=cg)); labels[b, 1:n + 1] = torch.randint(1, 9488, (n,), generator=cg); masks[b, :n + 2] = 1.0
top = torch.full((rows, 1000), -1, dtype=torch.int64)
for b in range(rows):
    n = int(torch.randint(2, 30, (1,), generator=cg)); top[b, :n] = torch.randperm(1000, generator=cg)[:n]
labels, masks, top = labels.to(dev), masks.to(dev), top.to(dev)
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
t = torch.ones(1, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
res = {}
for mode in ("between", "overlapped"):
    model = bench.build_model(dev); model.train()      # same seed -> same weights on every rank and in both modes; dropout 0
    params = list(model.parameters())
    opt = FusedAdam(params, lr=5e-4, weight_decay=1e-5, grad_clip=1.0, capturable=True, grad_scale=1.0 / world)
    if mode == "between":
        step = TR.GraphedXEStep(model, crit, opt, fc, att, labels, masks, top, 10.0, warmup=2,
                                between=lambda: D.average_gradients(params, divide=False))
    else:
        gs = D.OverlappedGradSync(params)
        step = TR.GraphedXEStep(model, crit, opt, fc, att, labels, masks, top, 10.0, warmup=2, grad_sync=gs)
        if rank == 0: print("buckets", gs.buckets_launched, flush=True)
    for _ in range(3): loss = step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    snap = [p.detach().clone() for p in params[:40]] + [params[-1].detach().clone()]
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    res[mode] = (e0.elapsed_time(e1) / 5, float(loss), snap)
    del step, opt, model, params
    torch.cuda.empty_cache()
worst = max(float((a - b).abs().max()) for a, b in zip(res["between"][2], res["overlapped"][2]))
if rank == 0:
    print(f"world {world}: between {res['between'][0]:.2f} ms  overlapped {res['overlapped'][0]:.2f} ms  loss {res['between'][1]:.4f} / {res['overlapped'][1]:.4f}  max param diff after 3 steps {worst:.2e}")
dist.destroy_process_group()
