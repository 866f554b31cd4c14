import sys, torch
sys.path.insert(0, '.')  # run from the repo root
import bench
from types import SimpleNamespace
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion
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
model.train(); model.dedup_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
from recurrent_fusion_network_b200.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=1e-5, grad_clip=1.0)
for it in range(2):
    opt.zero_grad(set_to_none=True)
    lp, rp = model(fc, att, labels)
    loss = crit(lp, labels[:, 1:], masks[:, 1:], rp, top, 10.0)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print(float(loss))
