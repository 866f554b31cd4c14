"""Near-tie analysis of engine mode 3 (TF32 + BF16 cross terms) on the 5000-image bench workload: captions against mode 1
(3xTF32) and mode 0 (fp32 SIMT), and against the oracle on a subset."""
import sys, json, torch
sys.path.insert(0, '.')
import bench
from recurrent_fusion_network_b200 import _capi
from oracle import rfnet_oracle as O
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
fc, att = bench.make_features(n, dev, seed=7)
res = {}
for mode in (3, 1, 0):
    _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
    with torch.no_grad():
        seq, slp, ds, dl, dp, nd, _ = model.beam_search(fc, att, 3, want_reason=False)
    res[mode] = (seq.cpu(), slp.cpu())
_capi.check(_capi.lib().rfn_set_gemm_mode(1))
out = dict(images=n)
for a, b in ((3, 1), (3, 0), (1, 0)):
    d = (res[a][0] != res[b][0]).any(dim=1)
    out[f'captions_differ_mode{a}_vs_mode{b}'] = int(d.sum())
    out[f'max_abs_seq_logprob_diff_on_equal_mode{a}_vs_mode{b}'] = float((res[a][1][~d] - res[b][1][~d]).abs().max())
m = int(sys.argv[2]) if len(sys.argv) > 2 else 48
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
torch.set_num_threads(16)
with torch.no_grad():
    oseq, oslp, ots, otp, _ = O.sample_beam(sd, O.RFNConfig(), [f[:m].cpu() for f in fc], [t[:m].cpu() for t in att], beam_size=3)
for mode in (3, 1, 0):
    d = (res[mode][0][:m] != oseq).any(dim=1)
    out[f'oracle_subset_{m}_mode{mode}_captions_differ'] = int(d.sum())
    out[f'oracle_subset_{m}_mode{mode}_max_lp_diff_on_equal'] = float((res[mode][1][:m][~d] - oslp[~d]).abs().max())
print(json.dumps(out))
