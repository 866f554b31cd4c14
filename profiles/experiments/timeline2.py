import sys, torch
EPI = int(sys.argv[1]) if len(sys.argv) > 1 else -1
sys.path.insert(0, '.')  # run from the repo root
import bench
from recurrent_fusion_network_b200 import _capi
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
n = 1500
fc, att = bench.make_features(n, dev, seed=7)
buf = torch.zeros(40000, 8, dtype=torch.int64, device='cuda')
with torch.no_grad():
    model.beam_search(fc, att, 3, want_reason=False)
    torch.cuda.synchronize()
    _capi.lib().rfn_debug_set_timeline(buf.data_ptr(), EPI)
    model.beam_search(fc, att, 3, want_reason=False)
    torch.cuda.synchronize()
    _capi.lib().rfn_debug_set_timeline(None, -1)
t = buf.cpu().double()
nct = 2 * 38 * ((n * 3 + 255) // 256)
lead = t[0:nct:2]
d = lambda a, b_: float((lead[:, b_] - lead[:, a]).median())
print(f"logits EPI2 (grid {nct}): init {d(0,1):.0f} fill {d(1,2):.0f} mainloop {d(2,3):.0f} lastMMA->drained {d(3,4):.0f} epilogue {d(4,5):.0f} teardown {d(5,6):.0f} total {d(0,6):.0f}")
