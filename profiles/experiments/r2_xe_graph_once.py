"""One replay of the graphed XE step (as written, 80 rows) for an ncu launch list: python r2_xe_graph_once.py [dedup] [replays]"""
import sys
import torch
sys.path.insert(0, '.')
import bench  # noqa: E402
from types import SimpleNamespace  # noqa: E402
from recurrent_fusion_network_b200 import training as TR  # noqa: E402
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion  # noqa: E402
from recurrent_fusion_network_b200.optim import FusedAdam  # noqa: E402

dd = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
g = torch.Generator(device=dev).manual_seed(100)
imgs, spi, L = 16, 5, 16
rows = imgs * spi
fc = [torch.randn(imgs, f, device=dev, generator=g).repeat_interleave(spi, 0) for (_, _, f) in bench.ENC]
att = [torch.randn(imgs, n, d, device=dev, generator=g).repeat_interleave(spi, 0) for (n, d, _) in bench.ENC]
cg = torch.Generator().manual_seed(200)
labels = torch.zeros(rows, L + 2, dtype=torch.int64); masks = torch.zeros(rows, L + 2)
for b in range(rows):
    n = int(torch.randint(5, L + 1, (1,), generator=cg)); labels[b, 1:n + 1] = torch.randint(1, 9488, (n,), generator=cg); masks[b, :n + 2] = 1.0
top = torch.full((rows, 1000), -1, dtype=torch.int64)
for b in range(rows):
    n = int(torch.randint(2, 30, (1,), generator=cg)); top[b, :n] = torch.randperm(1000, generator=cg)[:n]
labels, masks, top = labels.to(dev), masks.to(dev), top.to(dev)
model.train(); model.dedup_rows = dd
model.drop_prob_lm = model.decoder.drop_prob_lm = 0.3
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=1e-5, grad_clip=1.0, capturable=True)
gs = TR.GraphedXEStep(model, crit, opt, fc, att, labels, masks, top, 10.0, warmup=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    gs()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ms per replay", e0.elapsed_time(e1) / reps, "loss", float(gs.loss))
