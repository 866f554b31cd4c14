"""torchrun --nproc-per-node N: dist.average_gradients on the full model's parameter shapes -- correctness of the coalesced
in-place NCCL path (sum / mean / clamp) and its time against the flat-bucket path."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
from recurrent_fusion_network_b200 import dist as D, make_opt, setup
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = setup(make_opt()).cuda()
params = list(model.parameters())
def fill():
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (1 + i % 3))
fill(); D.average_gradients(params, divide=False)
ok = all(float(p.grad.flatten()[0]) == sum(range(1, world + 1)) * (1 + i % 3) for i, p in enumerate(params))
fill(); D.average_gradients(params, grad_clip=1.2)
want = lambda i: min(1.2, sum(range(1, world + 1)) * (1 + i % 3) / world)
ok &= all(abs(float(p.grad.flatten()[-1]) - want(i)) < 1e-6 for i, p in enumerate(params))
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_new = timed(lambda: D.average_gradients(params, divide=False))
os.environ["RFN_FLAT_ALLREDUCE"] = "1"
t_old = timed(lambda: D.average_gradients(params))
nbytes = sum(p.numel() for p in params) * 4
if rank == 0:
    print(f"world {world}: correct={ok}  coalesced in-place sum {t_new:.2f} ms ({nbytes / 1e9:.2f} GB; bus {2 * (world - 1) / world * nbytes / t_new / 1e6:.0f} GB/s)  flat buckets + divide {t_old:.2f} ms")
dist.destroy_process_group()
