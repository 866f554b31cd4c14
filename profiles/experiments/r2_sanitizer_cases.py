"""Small cases that drive every round-2 kernel (split-fp16 GEMM with store / score / vocabulary epilogues incl. a 3-source
GEMM, the operand split, the persistent cooperative decoder in greedy and beam mode, the bf16 engine) -- sized for
compute-sanitizer (memcheck / racecheck / synccheck), results checked against the oracle."""
import sys, torch
sys.path.insert(0, '.')
from oracle import rfnet_oracle as O
from recurrent_fusion_network_b200 import _capi
from tests._gpu_util import build_model, cuda_list, assert_beam_match_with_tie_policy
lib = _capi.lib()
cfg = O.RFNConfig(encoders=(O.Encoder(8, 256, 64), O.Encoder(4, 320, 32)), rnn_size=256, att_hid_size=256, input_encoding_size=256,
                  vocab_size=299, seq_length=4, num_review_steps_0=2, num_review_steps=2, top_words_count=24)
sd = O.make_state_dict(cfg, seed=3, init_range=0.3, logit_scale=2.0, eos_bias=0.5)
m = build_model(cfg, sd)
for mode in (4, 5):
    _capi.check(lib.rfn_set_gemm_mode(mode))
    fc, att = O.make_inputs(cfg, 90, seed=5)          # 270 beam rows: split engine with all three epilogues
    before = _capi.engine_launch_counts()
    with torch.no_grad():
        seq, slp, *_ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        torch.cuda.synchronize()
    ran = {k: v - before[k] for k, v in _capi.engine_launch_counts().items() if v - before[k]}
    print("mode", mode, ran)
    if mode == 4:
        margins = []
        with torch.no_grad():
            o = O.sample_beam(sd, cfg, fc, att, beam_size=3, margins_out=margins)
        print("ties", assert_beam_match_with_tie_policy(seq, slp, o[0], o[1], margins, "sanitizer case"))
_capi.check(lib.rfn_set_gemm_mode(4))
fc, att = O.make_inputs(cfg, 5, seed=6)               # 5 greedy rows / 15 beam rows: persistent decoder
with torch.no_grad():
    s, sl, la, _ = m.sample(cuda_list(fc), cuda_list(att), {"sample_max": 1})
    so, slo, lao, _ = O.sample(sd, cfg, fc, att)
    assert torch.equal(s.cpu(), so) and float((la.cpu() - lao).abs().max()) < 2e-4
    b = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
    o = O.sample_beam(sd, cfg, fc, att, beam_size=3)
    assert torch.equal(b[0].cpu(), o[0])
torch.cuda.synchronize()
print("SANITIZER_CASES_OK", _capi.engine_launch_counts())
