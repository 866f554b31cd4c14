"""One replay of the graphed RL step (as written, 250 rows) for an ncu launch list: python r2_rl_graph_once.py [dedup] [replays]"""
import sys
import torch
sys.path.insert(0, '.')
import bench  # noqa: E402
from types import SimpleNamespace  # noqa: E402
from recurrent_fusion_network_b200 import reward as RW, training as TR  # noqa: E402
from recurrent_fusion_network_b200.criteria import ReviewNetRewardCriterion  # noqa: E402
from recurrent_fusion_network_b200.optim import FusedAdam  # noqa: E402

dd = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
device = torch.device('cuda', 0)
model = bench.build_model(device)
g = torch.Generator(device=device).manual_seed(300)
imgs, spi, L = 50, 5, model.seq_length
rows = imgs * spi
fc = [torch.randn(imgs, f, device=device, generator=g).repeat_interleave(spi, 0) for (_, _, f) in bench.ENC]
att = [torch.randn(imgs, n, d, device=device, generator=g).repeat_interleave(spi, 0) for (n, d, _) in bench.ENC]
cg = torch.Generator().manual_seed(400)
gts, df = [], {}
for _ in range(imgs):
    refs, seen = [], set()
    for _ in range(5):
        n = int(torch.randint(5, L + 1, (1,), generator=cg))
        r = torch.randint(1, 9488, (n,), generator=cg).tolist() + [0]
        refs.append(r)
        for k in range(1, 5):
            seen.update(tuple(r[j:j + k]) for j in range(len(r) - k + 1))
    gts.append(refs)
    for ng in seen:
        df[ng] = df.get(ng, 0.0) + 1.0
table = RW.DocumentFrequency(df, imgs, device)
top = torch.full((rows, 1000), -1, dtype=torch.int64)
for b in range(rows):
    n = int(torch.randint(2, 30, (1,), generator=cg))
    top[b, :n] = torch.randperm(1000, generator=cg)[:n]
top = top.to(device)
ropt = SimpleNamespace(cider_weight=1.0, bleu4_weight=0, spice_weight=0, use_baseline=1, use_ppo=0)
crit = ReviewNetRewardCriterion(SimpleNamespace(use_label_smoothing=0, label_smoothing_epsilon=0.1))
params = [p for p in model.parameters()]
u = torch.rand(rows, L, device=device)
model.train(); model.dedup_rows = dd
opt_g = FusedAdam(params, lr=5e-5, weight_decay=1e-5, grad_clip=1.0, capturable=True)
gs = TR.GraphedRLStep(model, crit, opt_g, fc, att, u, top, gts, table, ropt, spi, 10.0, entropy_reg=0.0, warmup=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    gs(uniforms=u)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ms per replay", e0.elapsed_time(e1) / reps, "loss", float(gs.loss))
