"""Full-size near-tie analysis: captions of the tensor engine (mode 1) vs the fp32 SIMT engine (mode 0)
on the 5000-image bench workload, and both against the oracle on a subset, classifying every divergence by
the decision margin."""
import sys, json, torch
sys.path.insert(0, '.')  # run from the repo root
import bench
from recurrent_fusion_network_b200 import _capi
from oracle import rfnet_oracle as O
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
fc, att = bench.make_features(n, dev, seed=7)
res = {}
for mode in (1, 0):
    _capi.check(_capi.lib().rfn_set_gemm_mode(mode))
    with torch.no_grad():
        seq, slp, ds, dl, dp, nd, _ = model.beam_search(fc, att, 3, want_reason=False)
    res[mode] = (seq.cpu(), slp.cpu(), dp.cpu(), nd.cpu())
_capi.check(_capi.lib().rfn_set_gemm_mode(1))
a, b = res[1], res[0]
diff = (a[0] != b[0]).any(dim=1)
out = dict(images=n, captions_differ_mode1_vs_mode0=int(diff.sum()))
# for differing images: gap between best and second-best finished beam score (the final decision margin)
def top2gap(dp, nd):
    g = []
    for k in range(dp.shape[0]):
        p = dp[k, :nd[k]]
        g.append(float(p[0] - p[1]) if nd[k] > 1 else float('inf'))
    return torch.tensor(g)
g1 = top2gap(a[2], a[3])
out['final_margin_of_differing_images_mode1'] = sorted(g1[diff].tolist())[:20]
out['max_abs_seq_logprob_diff_on_equal_captions'] = float((a[1][~diff] - b[1][~diff]).abs().max())
out['sum_logprob_diff_stats'] = dict(mean=float((a[1].sum(1) - b[1].sum(1))[~diff].abs().mean()))
# oracle on a subset
m = int(sys.argv[2]) if len(sys.argv) > 2 else 48
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
cfg = O.RFNConfig()
torch.set_num_threads(16)
with torch.no_grad():
    oseq, oslp, ots, otp, _ = O.sample_beam(sd, cfg, [f[:m].cpu() for f in fc], [t[:m].cpu() for t in att], beam_size=3)
for mode in (1, 0):
    d = (res[mode][0][:m] != oseq).any(dim=1)
    out[f'oracle_subset_{m}_mode{mode}_captions_differ'] = int(d.sum())
    out[f'oracle_subset_{m}_mode{mode}_max_lp_diff_on_equal'] = float((res[mode][1][:m][~d] - oslp[~d]).abs().max())
    margins = []
    for k in torch.nonzero(d).flatten().tolist():
        p = otp[k]
        margins.append(p[0] - p[1] if len(p) > 1 else None)
    out[f'oracle_subset_{m}_mode{mode}_oracle_final_margins_of_differing'] = margins
print(json.dumps(out))
