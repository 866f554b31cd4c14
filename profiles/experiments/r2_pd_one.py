import sys, torch
sys.path.insert(0, '.')
from oracle import rfnet_oracle as O
from tests._gpu_util import build_model
cfg = O.config1(49)
sd = O.make_state_dict(cfg, seed=1234, sharpen=True)
m = build_model(cfg, sd)
fc, att = O.make_inputs(cfg, 16, seed=7)
with torch.no_grad():
    for _ in range(2):
        m.sample([t.cuda() for t in fc], [t.cuda() for t in att], {"sample_max": 1, "return_logprobs_all": False})
torch.cuda.synchronize()
