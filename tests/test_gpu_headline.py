"""GPU: the configuration bench.py's headline number is measured in, gated against the ORACLE.

bench.py decodes with engine mode 3 (TF32 hi.hi + BF16 cross terms) and thousands of decoder rows, where the
logits GEMM runs as the persistent 2-CTA kernel with the fused vocabulary epilogue (`gemm_tc2p_kernel<2>`,
needs >= 256 decoder rows and >= 256 vocabulary columns) and stage 1 as `gemm_tc2p_kernel<1>`.  The fixtures
of tests/golden stop at 48 decoder rows, so these tests run beam 3 at >= 256 decoder rows with the 9488-way
vocabulary, check WHICH kernel families launched (rfn_engine_launch_counts) and compare captions with the
oracle under the near-tie policy of SURVEY.md section 4.3 (misc/RecurrentFusionModel.py:352-543)."""
import pytest
import torch

from oracle import rfnet_oracle as O
from tests._gpu_util import LP_TOL, assert_beam_match_with_tie_policy, build_model, cuda_list, maxdiff

pytestmark = pytest.mark.gpu


@pytest.fixture()
def mode3():
    from recurrent_fusion_network_b200 import _capi
    prev_mode, prev_cl = _capi.lib().rfn_get_gemm_mode(), _capi.lib().rfn_get_tc_cluster()
    _capi.check(_capi.lib().rfn_set_gemm_mode(3))
    _capi.check(_capi.lib().rfn_set_tc_cluster(2))
    torch.set_num_threads(max(16, torch.get_num_threads()))
    yield
    _capi.check(_capi.lib().rfn_set_gemm_mode(prev_mode))
    _capi.check(_capi.lib().rfn_set_tc_cluster(prev_cl))


CASES = [
    # name, config, images (x beam 3 = decoder rows), sharpened weights (EOS fires -> mixed lengths, done-beam overflow)
    ("config1_sharp_288rows", lambda: O.config1(49), 96, True),
    ("full_j5_bench_init_264rows", lambda: O.RFNConfig(), 88, False),     # bench.py's weights: reference init, seed 1234
    ("full_j5_sharp_264rows", lambda: O.RFNConfig(), 88, True),
]


@pytest.mark.parametrize("name,make_cfg,images,sharpen", CASES, ids=[c[0] for c in CASES])
def test_headline_engine_beam3_matches_oracle(mode3, name, make_cfg, images, sharpen):
    from recurrent_fusion_network_b200 import _capi
    cfg = make_cfg()
    sd = O.make_state_dict(cfg, seed=1234, sharpen=sharpen)
    fc, att = O.make_inputs(cfg, images, seed=31)
    m = build_model(cfg, sd)
    before = _capi.engine_launch_counts()
    with torch.no_grad():
        seq, slp, top_seq, top_prob, _ = m.sample_beam(cuda_list(fc), cuda_list(att), {"beam_size": 3})
        torch.cuda.synchronize()
    after = _capi.engine_launch_counts()
    ran = {k: after[k] - before[k] for k in after}
    # the instantiations the bench runs: fused vocabulary epilogue on the persistent 2-CTA kernel, one launch per decoder
    # step, and the fused attention-score epilogue for every (stage-1 step, encoder)
    assert ran["tcgen05_2cta_persistent_vocab"] == cfg.seq_length, ran
    assert ran["tcgen05_2cta_persistent_score"] >= cfg.num_review_steps_0 * cfg.J, ran
    margins = []
    with torch.no_grad():
        o_seq, o_slp, o_top_seq, o_top_prob, _ = O.sample_beam(sd, cfg, fc, att, beam_size=3, margins_out=margins)
    ties = assert_beam_match_with_tie_policy(seq, slp, o_seq, o_slp, margins, name)
    assert ties <= 2, f"{name}: {ties} tie-broken captions out of {images}"
    # the finished-beam lists (top_seq / top_prob) of every image that was not tie-broken
    for k in range(images):
        if not torch.equal(seq[k].cpu(), o_seq[k]):
            continue
        if top_seq[k].shape != o_top_seq[k].shape or not torch.equal(top_seq[k], o_top_seq[k]):
            mk = min(min(margins[k]["steps"], default=float("inf")), margins[k]["final"])
            assert mk < 1e-5, f"{name}: image {k} finished-beam list differs (oracle margin {mk:.3g})"
            continue
        assert maxdiff(torch.tensor(top_prob[k]), torch.tensor(o_top_prob[k])) <= 2 * LP_TOL


def test_headline_engine_greedy_256_rows_matches_oracle(mode3):
    """Greedy decode of 256 rows (config 1 shapes): the tensor-engine gates / logits GEMMs of the sample path."""
    cfg = O.config1(49)
    sd = O.make_state_dict(cfg, seed=1234, sharpen=True)
    fc, att = O.make_inputs(cfg, 256, seed=32)
    m = build_model(cfg, sd)
    with torch.no_grad():
        s, sl, la, _ = m.sample(cuda_list(fc), cuda_list(att), {"sample_max": 1})
        so, slo, lao, _ = O.sample(sd, cfg, fc, att)
    from tests._gpu_util import assert_tokens_match_with_tie_policy
    T = min(s.shape[1], so.shape[1])
    ties = assert_tokens_match_with_tie_policy(s[:, :T], so[:, :T], lao, "greedy 256 rows")
    assert ties <= 2
    same = (s[:, :T].cpu() == so[:, :T]).all(dim=1)
    assert maxdiff(sl[:, :T][same.cuda()], slo[:, :T][same]) <= LP_TOL
    assert maxdiff(la[:, :T + 1][same.cuda()], lao[:, :T + 1][same]) <= LP_TOL
