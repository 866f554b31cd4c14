"""Per-phase timeline of the persistent cooperative decoder (CTA 0, globaltimer) and the batch-16 greedy latency with the
persistent kernel on / off, eager and replayed from a CUDA graph."""
import json, sys, torch
sys.path.insert(0, '.')
from oracle import rfnet_oracle as O
from recurrent_fusion_network_b200 import _capi
from recurrent_fusion_network_b200.graphs import GraphedSample
from tests._gpu_util import build_model
lib = _capi.lib()
cfg = O.config1(49)
sd = O.make_state_dict(cfg, seed=1234, sharpen=True)
m = build_model(cfg, sd)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fc, att = O.make_inputs(cfg, rows, seed=7)
fcg, attg = [t.cuda() for t in fc], [t.cuda() for t in att]
opt = {"sample_max": 1, "return_logprobs_all": False}
L = cfg.seq_length
buf = torch.zeros((L + 1) * 8, dtype=torch.int64, device='cuda')
with torch.no_grad():
    m.sample(fcg, attg, opt)
    _capi.check(lib.rfn_debug_set_pd_timeline(buf.data_ptr()))
    m.sample(fcg, attg, opt)
    torch.cuda.synchronize()
    _capi.check(lib.rfn_debug_set_pd_timeline(None))
st = buf.cpu().view(L + 1, 8)[:L].double()
names = ["C gates+cell", "barrier", "D logits+g_next", "barrier", "E merge+select+attention", "barrier", "beam attention (+barrier)"]
d = (st[:, 1:] - st[:, :-1])[1:]          # skip step 0
out = {n: round(float(d[:, i].mean()) / 1e3, 2) for i, n in enumerate(names)}
out["step_us"] = round(float((st[2:, 0] - st[1:-1, 0]).mean()) / 1e3, 2)
print(json.dumps(out))

def lat(fn, reps=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {}
for on in (1, 0):
    _capi.check(lib.rfn_set_persistent_decoder(on))
    with torch.no_grad():
        res[f"eager_ms_pd{on}"] = round(lat(lambda: m.sample(fcg, attg, opt), 50), 4)
    g = GraphedSample(m, fcg, attg, opt)
    res[f"graph_ms_pd{on}"] = round(lat(lambda: g(), 200), 4)
    res[f"kernels_pd{on}"] = g.kernels_per_replay
    del g
_capi.check(lib.rfn_set_persistent_decoder(1))
print(json.dumps(res))
json.dump(dict(rows=rows, phases_us=out, latency=res), open(f'gpurun_out/r2_pd_timeline_{rows}.json', 'w'), indent=1)
