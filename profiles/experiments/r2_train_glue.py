"""Where do the at:: fill / add / copy kernels of one eager XE training step come from?  torch.profiler, every aten op that
launches a kernel attributed to the outermost enclosing op (an autograd node in backward, the python frame in forward)."""
import collections
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from types import SimpleNamespace  # noqa: E402
from recurrent_fusion_network_b200.criteria import ReviewNetEnsembleCriterion  # noqa: E402
from recurrent_fusion_network_b200.optim import FusedAdam  # noqa: E402

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
model.train(); model.dedup_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1
crit = ReviewNetEnsembleCriterion(SimpleNamespace(use_label_smoothing=1, label_smoothing_epsilon=0.1, use_cuda=1))
opt = FusedAdam(model.parameters(), lr=5e-4, weight_decay=1e-5, grad_clip=1.0)


def step():
    opt.zero_grad(set_to_none=True)
    lp, rp = model(fc, att, labels)
    loss = crit(lp, labels[:, 1:], masks[:, 1:], rp, top, 10.0)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()

ev = prof.events()
by = collections.Counter()
kern = collections.Counter()
for e in ev:
    if not e.kernels:
        continue
    if any(c.kernels for c in e.cpu_children):     # count the innermost op that owns the kernel
        continue
    top_ = e
    chain = [e.name]
    while top_.cpu_parent is not None:
        top_ = top_.cpu_parent
        chain.append(top_.name)
    where = ''
    if e.stack:
        fr = [s for s in e.stack if 'recurrent_fusion_network_b200' in s or 'bench' in s or 'r2_train_glue' in s]
        where = fr[0].split('/')[-1] if fr else ''
    # keep: innermost op, nearest autograd node (or python frame)
    node = next((c for c in chain if 'evaluate_function' in c or 'Backward' in c or 'Fn' in c), chain[-1])
    by[(e.name, node[:70], where[:60])] += len(e.kernels)
    for k in e.kernels:
        kern[k.name[:60]] += 1
print('kernels by name:')
for k, v in kern.most_common(25):
    print(f'{v:6d}  {k}')
print('\nkernel-launching aten ops by (op, enclosing node, python frame):')
for k, v in by.most_common(70):
    print(f'{v:6d}  {k[0]:28s} {k[1]:70s} {k[2]}')
