"""CPU oracle for RFNet's recurrent fusion + decode path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 restatement of the reference algorithm
(cswhjiang/Recurrent_Fusion_Network).  It exists to CHECK the CUDA path; it is never
the thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is the reference's own Python modules imported
read-only from /root/reference by ``oracle/gen_golden.py`` in the build container:
that script checks every function below against the reference module on identical
weights and inputs and writes the small fixtures under ``tests/golden/`` which the
CPU test-suite re-checks on every run.

Every function cites the reference file:line it restates (paths relative to the
reference repo root).  Weights are addressed by the reference's ``state_dict`` key
names so that a reference checkpoint drives the oracle unchanged.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class Encoder:
    att_num: int        # N_j
    att_feat_size: int  # D_j
    fc_feat_size: int   # F_j


# feat_array.py:6-9 (resnet), :240-244 (order of the five encoders)
FULL_ENCODERS = (
    Encoder(196, 2048, 2048),  # resnet-101 14x14
    Encoder(64, 1536, 1536),   # inception_v4
    Encoder(64, 1280, 2048),   # inception_v3
    Encoder(49, 2208, 2208),   # densenet-161
    Encoder(64, 1536, 1536),   # inception_resnet_v2
)


@dataclasses.dataclass(frozen=True)
class RFNConfig:
    """The fields of ``opt`` that RecurrentFusionModel reads (misc/RecurrentFusionModel.py:120-151)."""
    encoders: Tuple[Encoder, ...] = FULL_ENCODERS
    rnn_size: int = 512              # opts.py:55
    att_hid_size: int = 512          # opts.py:63
    input_encoding_size: int = 512   # opts.py:61
    vocab_size: int = 9487           # set from the loader, main.py:37
    seq_length: int = 16             # main.py:38
    num_review_steps_0: int = 8      # opts.py:207-210
    num_review_steps: int = 8
    top_words_count: int = 1000      # opts.py:23
    review_maxout: int = 0           # opts.py:182 (stage-2 cells 5R wide, misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:25)
    decoder_maxout: int = 0          # opts.py:180 `--maxout` (misc/LSTMSoftAttentionCore.py:25)

    @property
    def J(self) -> int:
        return len(self.encoders)

    @property
    def V1(self) -> int:
        return self.vocab_size + 1


def config1(att_num: int = 49) -> RFNConfig:
    """BASELINE.json configs[0]: single encoder, 7x7x2048 (or 14x14 with att_num=196)."""
    return RFNConfig(encoders=(Encoder(att_num, 2048, 2048),))


def tiny_config(J: int = 2) -> RFNConfig:
    """Small shapes for golden fixtures and fast CPU tests (all dims multiples of 4)."""
    encs = (Encoder(12, 40, 24), Encoder(8, 24, 32), Encoder(5, 16, 16))[:J]
    return RFNConfig(encoders=encs, rnn_size=32, att_hid_size=16, input_encoding_size=24,
                     vocab_size=59, seq_length=6, num_review_steps_0=3, num_review_steps=2,
                     top_words_count=20)


# --------------------------------------------------------------------------------------
# state_dict layout (773 tensors for the full model) and deterministic synthetic weights
# --------------------------------------------------------------------------------------
def state_dict_shapes(cfg: RFNConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape, in the reference's registration order (misc/RecurrentFusionModel.py:153-184)."""
    R, A, E, K, V1, J = (cfg.rnn_size, cfg.att_hid_size, cfg.input_encoding_size,
                         cfg.top_words_count, cfg.V1, cfg.J)
    out: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        out[name + ".weight"] = (o, i)
        out[name + ".bias"] = (o,)

    def att(prefix, D):  # misc/AttentionModelCore.py:16-18
        lin(prefix + ".att_2_att_h", A, D)
        lin(prefix + ".h_2_att_h", A, R)
        lin(prefix + ".att_h_2_out", 1, A)

    for j, e in enumerate(cfg.encoders):
        lin(f"fc2h.{j}", R, e.fc_feat_size)
    out["embed.weight"] = (V1, E)
    lin("logit", V1, R)
    for s in range(cfg.num_review_steps_0):
        for j, e in enumerate(cfg.encoders):
            p = f"review_steps_individual.{s}.lstm.{j}"
            att(p + ".att_model", e.att_feat_size)
            lin(p + ".H2h", 4 * R, J * R)
            lin(p + ".z2h", 4 * R, e.att_feat_size)
    for j in range(J):
        lin(f"reason_linear_individual.{j}", K, R)
    g2 = (5 if cfg.review_maxout else 4) * R
    gd = (5 if cfg.decoder_maxout else 4) * R
    for s in range(cfg.num_review_steps):
        p = f"review_steps.{s}"
        lin(p + ".h2h", g2, R)
        for j in range(J):
            lin(p + f".z_2_h.{j}", g2, R)
        for j in range(J):
            att(p + f".att_model.{j}", R)
    lin("reason_linear", K, R)
    lin("decoder.i2h", gd, E)
    lin("decoder.h2h", gd, R)
    lin("decoder.z2h", gd, R)
    lin("decoder.att_2_att_h", A, R)
    lin("decoder.h_2_att_h", A, R)
    lin("decoder.att_h_2_out", 1, A)
    return out


def make_state_dict(cfg: RFNConfig, seed: int = 1234, sharpen: bool = False,
                    init_range: float = 0.1, logit_scale: Optional[float] = None,
                    eos_bias: Optional[float] = None) -> StateDict:
    """Deterministic synthetic weights, U(-init_range, init_range) like the reference's init
    (misc/RecurrentFusionModel.py:188-196, misc/AttentionModelCore.py:21-29), drawn from one
    seeded CPU generator in ``state_dict_shapes`` order so every machine with the same torch
    regenerates identical values.  ``sharpen`` applies SURVEY.md section 4.2's recipe that makes
    EOS fire at full size (logit.weight *= 30, logit.bias[0] = 7); ``logit_scale`` / ``eos_bias``
    set those two knobs individually (used by the tiny fixtures)."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for k, shp in state_dict_shapes(cfg).items():
        sd[k] = (torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * init_range
    sd["logit.bias"].zero_()  # RecurrentFusionModel.py:192
    if sharpen:
        logit_scale = 30.0 if logit_scale is None else logit_scale
        eos_bias = 7.0 if eos_bias is None else eos_bias
    if logit_scale is not None:
        sd["logit.weight"] *= logit_scale
    if eos_bias is not None:
        sd["logit.bias"][0] = eos_bias
    return sd


def make_inputs(cfg: RFNConfig, rows: int, seed: int = 7):
    """fc_j ~ N(0,1) (rows,F_j), att_j ~ N(0,1) (rows,N_j,D_j): the distribution of the reference's
    own smoke block (misc/RecurrentFusionModel.py:685-693)."""
    g = torch.Generator().manual_seed(seed)
    fc = [torch.randn(rows, e.fc_feat_size, generator=g) for e in cfg.encoders]
    att = [torch.randn(rows, e.att_num, e.att_feat_size, generator=g) for e in cfg.encoders]
    return fc, att


def make_labels(cfg: RFNConfig, rows: int, seed: int = 11, min_len: int = 2):
    """labels (rows, L+2) int64 with column 0 = BOS 0 and zero padding; masks = 1 for the first
    len+2 positions (dataloader.py:312-314); top_words (rows, K) int64, -1 terminated."""
    g = torch.Generator().manual_seed(seed)
    L = cfg.seq_length
    labels = torch.zeros(rows, L + 2, dtype=torch.int64)
    masks = torch.zeros(rows, L + 2, dtype=torch.float32)
    lens = torch.randint(min(min_len, L), L + 1, (rows,), generator=g)
    for b in range(rows):
        n = int(lens[b])
        labels[b, 1:n + 1] = torch.randint(1, cfg.V1, (n,), generator=g)
        masks[b, :n + 2] = 1.0
    K = cfg.top_words_count
    top = torch.full((rows, K), -1, dtype=torch.int64)
    for b in range(rows):
        n = int(torch.randint(2, max(3, K // 4), (1,), generator=g))
        top[b, :n] = torch.randperm(K, generator=g)[:n]
    return labels, masks, top


# --------------------------------------------------------------------------------------
# A.1  additive soft attention            misc/AttentionModelCore.py:31-48
# --------------------------------------------------------------------------------------
def attention(sd: StateDict, prefix: str, h: Tensor, A: Tensor) -> Tensor:
    """z = softmax_n(w . tanh(U A_n + W h + b)) . A   -- no mask.
    ``prefix`` names the module holding att_2_att_h / h_2_att_h / att_h_2_out."""
    rows, N, D = A.shape
    P = F.linear(A.reshape(-1, D), sd[prefix + ".att_2_att_h.weight"],
                 sd[prefix + ".att_2_att_h.bias"]).view(rows, N, -1)          # :32-34
    g = F.linear(h, sd[prefix + ".h_2_att_h.weight"], sd[prefix + ".h_2_att_h.bias"])  # :36
    t = torch.tanh(g.unsqueeze(1) + P)                                         # :37-39
    e = F.linear(t.reshape(rows * N, -1), sd[prefix + ".att_h_2_out.weight"],
                 sd[prefix + ".att_h_2_out.bias"]).view(rows, N)               # :41-43
    a = torch.softmax(e, dim=1)                                                # :44
    return torch.bmm(A.transpose(1, 2), a.unsqueeze(2)).squeeze(2)             # :45-47


def attention_weights(sd: StateDict, prefix: str, h: Tensor, A: Tensor) -> Tuple[Tensor, Tensor]:
    """Same as ``attention`` but also returns the pre-softmax scores (for margin analysis)."""
    rows, N, D = A.shape
    P = F.linear(A.reshape(-1, D), sd[prefix + ".att_2_att_h.weight"],
                 sd[prefix + ".att_2_att_h.bias"]).view(rows, N, -1)
    g = F.linear(h, sd[prefix + ".h_2_att_h.weight"], sd[prefix + ".h_2_att_h.bias"])
    e = F.linear(torch.tanh(g.unsqueeze(1) + P).reshape(rows * N, -1),
                 sd[prefix + ".att_h_2_out.weight"], sd[prefix + ".att_h_2_out.bias"]).view(rows, N)
    return e, torch.softmax(e, dim=1)


# --------------------------------------------------------------------------------------
# A.2  LSTM update, gate order [i | f | o | g]     misc/RecurrentFusionModel.py:55-73
# --------------------------------------------------------------------------------------
def lstm_cell(G: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
    R = c.shape[1]
    sig = torch.sigmoid(G[:, :3 * R])
    i, f, o = sig[:, :R], sig[:, R:2 * R], sig[:, 2 * R:3 * R]
    if G.shape[1] == 5 * R:      # maxout variant: no tanh (misc/LSTMSoftAttentionCore.py:88-91, ...FeatArrayNoInputCore.py:59-62)
        g = torch.max(G[:, 3 * R:4 * R], G[:, 4 * R:5 * R])
    else:
        g = torch.tanh(G[:, 3 * R:4 * R])
    c2 = f * c + i * g
    h2 = o * torch.tanh(c2)
    return h2, c2  # dropout is identity in eval mode / p = 0


# --------------------------------------------------------------------------------------
# A.0  init state                         misc/RecurrentFusionModel.py:202-208, :333-343
# --------------------------------------------------------------------------------------
def get_init_state(sd: StateDict, cfg: RFNConfig, fc: Sequence[Tensor]):
    st = []
    for j in range(cfg.J):
        h0 = F.linear(fc[j], sd[f"fc2h.{j}.weight"], sd[f"fc2h.{j}.bias"])
        st.append((h0, h0.clone()))
    return st


# --------------------------------------------------------------------------------------
# A.3  stage-1 fusion step                misc/RecurrentFusionModel.py:47-74, :101-114
# --------------------------------------------------------------------------------------
def stage1_step(sd: StateDict, cfg: RFNConfig, s: int, att: Sequence[Tensor], state):
    H = torch.cat([h for (h, _) in state], dim=1)       # :102-107, all PREVIOUS hidden states
    new_state = []
    for j in range(cfg.J):
        p = f"review_steps_individual.{s}.lstm.{j}"
        h, c = state[j]
        z = attention(sd, p + ".att_model", h, att[j])                         # :51
        G = F.linear(H, sd[p + ".H2h.weight"], sd[p + ".H2h.bias"]) + \
            F.linear(z, sd[p + ".z2h.weight"], sd[p + ".z2h.bias"])            # :53
        new_state.append(lstm_cell(G, c))
    return new_state


# --------------------------------------------------------------------------------------
# A.5  stage-2 review step        misc/LSTMSoftMultiAttentionFeatArrayNoInputCore.py:41-73
# --------------------------------------------------------------------------------------
def stage2_step(sd: StateDict, cfg: RFNConfig, s: int, TV: Sequence[Tensor], state):
    h, c = state
    p = f"review_steps.{s}"
    zs = [attention(sd, p + f".att_model.{j}", h, TV[j]) for j in range(cfg.J)]   # :46-48
    G = F.linear(h, sd[p + ".h2h.weight"], sd[p + ".h2h.bias"])                    # :50
    for j in range(cfg.J):
        G = G + F.linear(zs[j], sd[p + f".z_2_h.{j}.weight"], sd[p + f".z_2_h.{j}.bias"])  # :51-52
    return lstm_cell(G, c)


# --------------------------------------------------------------------------------------
# A.3-A.5 orchestration: thought vectors    misc/RecurrentFusionModel.py:283-331
# --------------------------------------------------------------------------------------
def get_thought_vectors(sd: StateDict, cfg: RFNConfig, att: Sequence[Tensor], state,
                        return_individual: bool = False):
    J = cfg.J
    tv = [[] for _ in range(J)]
    rm = [[] for _ in range(J)]
    for s in range(cfg.num_review_steps_0):
        state = stage1_step(sd, cfg, s, att, state)
        for j in range(J):
            tv[j].append(state[j][0])
            rm[j].append(F.linear(state[j][0], sd[f"reason_linear_individual.{j}.weight"],
                                  sd[f"reason_linear_individual.{j}.bias"]))
    TV = [torch.stack(tv[j], dim=1).contiguous() for j in range(J)]               # (rows,S0,R)
    reason_pred = [torch.stack(rm[j], dim=1).max(dim=1)[0] for j in range(J)]     # :303
    h = sum(st[0] for st in state) / J                                            # :307-309
    c = sum(st[1] for st in state) / J
    st2 = (h, c)
    tvc, rmc = [], []
    for s in range(cfg.num_review_steps):
        st2 = stage2_step(sd, cfg, s, TV, st2)
        tvc.append(st2[0])
        rmc.append(F.linear(st2[0], sd["reason_linear.weight"], sd["reason_linear.bias"]))
    TVc = torch.stack(tvc, dim=1).contiguous()                                    # (rows,S1,R)
    reason_pred.append(torch.stack(rmc, dim=1).max(dim=1)[0])
    if return_individual:
        return TVc, reason_pred, st2, TV
    return TVc, reason_pred, st2


# --------------------------------------------------------------------------------------
# A.6  decoder step                        misc/LSTMSoftAttentionCore.py:60-102
# --------------------------------------------------------------------------------------
def decoder_step(sd: StateDict, x: Tensor, TVc: Tensor, state):
    h, c = state
    z = attention(sd, "decoder", h, TVc)                                          # :64-79
    G = F.linear(x, sd["decoder.i2h.weight"], sd["decoder.i2h.bias"]) + \
        F.linear(h, sd["decoder.h2h.weight"], sd["decoder.h2h.bias"]) + \
        F.linear(z, sd["decoder.z2h.weight"], sd["decoder.z2h.bias"])             # :81
    return lstm_cell(G, c)


def one_time_step(sd: StateDict, x: Tensor, TVc: Tensor, state):
    """misc/RecurrentFusionModel.py:345-350 -- returns LOGITS, not log-probs."""
    h, c = decoder_step(sd, x, TVc, state)
    return F.linear(h, sd["logit.weight"], sd["logit.bias"]), (h, c)


def _step_logprobs(sd, tok, TVc, state):
    x = sd["embed.weight"][tok]
    logits, state = one_time_step(sd, x, TVc, state)
    return torch.log_softmax(logits, dim=1), state


# --------------------------------------------------------------------------------------
# A.7  XE teacher-forced forward           misc/RecurrentFusionModel.py:198-281
# --------------------------------------------------------------------------------------
def forward_xe(sd: StateDict, cfg: RFNConfig, fc, att, seq: Tensor):
    """Returns (logprobs (rows, T', V1), reason_pred list[J+1]).  Scheduled sampling
    (ss_prob > 0, :260-270) draws from torch's RNG and is not restated."""
    state = get_init_state(sd, cfg, fc)
    TVc, reason_pred, st = get_thought_vectors(sd, cfg, att, state)
    outs = []
    for i in range(seq.shape[1]):
        if i >= 1 and int(seq[:, i].sum()) == 0:                                  # :274-275
            break
        lp, st = _step_logprobs(sd, seq[:, i], TVc, st)
        outs.append(lp)
    return torch.stack(outs, dim=1).contiguous(), reason_pred


# --------------------------------------------------------------------------------------
# Appendix B  greedy / multinomial sample  misc/RecurrentFusionModel.py:545-658
# --------------------------------------------------------------------------------------
def inverse_cdf_sample(p: Tensor, u: Tensor) -> Tensor:
    """Token = first index whose inclusive prefix sum of p (fp32, index order) exceeds u * total.
    The reference draws with torch.multinomial on the CPU RNG (:624-631), which no GPU kernel can
    replay; oracle and kernel instead share externally supplied uniforms u in [0,1)."""
    cdf = torch.cumsum(p.double(), dim=1)
    thr = (u.double() * cdf[:, -1]).unsqueeze(1)
    idx = (cdf <= thr).sum(dim=1)
    return idx.clamp_(max=p.shape[1] - 1)


def sample(sd: StateDict, cfg: RFNConfig, fc, att, sample_max: int = 1, temperature: float = 1.0,
           uniforms: Optional[Tensor] = None, forced_tokens: Optional[Tensor] = None):
    """Returns (seq (rows,T) i64, seqLogprobs (rows,T), logprobs_all (rows,T+1,V1), reason_pred).
    ``uniforms`` (rows, L) feeds the multinomial path; ``forced_tokens`` (rows, L) replays a given
    token sequence (used to pin log-probs given the kernel's own samples)."""
    rows = fc[0].shape[0]
    state = get_init_state(sd, cfg, fc)
    TVc, reason_pred, st = get_thought_vectors(sd, cfg, att, state)
    seq, slp, lp_all = [], [], []
    lp = None
    unfinished = None
    for t in range(cfg.seq_length + 1):
        if t == 0:
            it = torch.zeros(rows, dtype=torch.int64)                             # :617-618
        elif forced_tokens is not None:
            it = forced_tokens[:, t - 1].clone()
            s_lp = lp.gather(1, it.unsqueeze(1)).squeeze(1)
        elif sample_max:
            s_lp, it = torch.max(lp, dim=1)                                       # :620
        else:
            p = torch.exp(lp) if temperature == 1.0 else torch.exp(lp / temperature)  # :623-627
            it = inverse_cdf_sample(p, uniforms[:, t - 1])
            s_lp = lp.gather(1, it.unsqueeze(1)).squeeze(1)                       # :632
        x_tok = it                                                                # :637 unmasked token
        if t >= 1:
            unfinished = (it > 0) if t == 1 else unfinished & (it > 0)            # :641-644
            if int(unfinished.sum()) == 0:                                        # :645
                break
            it = it * unfinished.to(it.dtype)                                     # :647
            seq.append(it)
            slp.append(s_lp)
        lp, st = _step_logprobs(sd, x_tok, TVc, st)                               # :651-653
        lp_all.append(lp)
    return (torch.stack(seq, 1), torch.stack(slp, 1), torch.stack(lp_all, 1).contiguous(),
            reason_pred)


# --------------------------------------------------------------------------------------
# Appendix C  beam search                   misc/RecurrentFusionModel.py:352-543
# --------------------------------------------------------------------------------------
def _topk_desc(lp: Tensor, k: int):
    """Top-k per row, descending, ties -> lower index first.  The reference does a full
    torch.sort(descending) (:463) whose tie order is unspecified; only columns < beam are read."""
    ys, ix = torch.sort(lp, dim=1, descending=True, stable=True)
    return ys[:, :k], ix[:, :k]


def beam_merge(beam: int, t: int, L: int, ys: Tensor, ix: Tensor, beam_seq: Tensor,
               beam_lp: Tensor, beam_sum: Tensor, done: list, margins: Optional[list] = None):
    """One beam merge step (:465-514) on CPU tensors.  ys/ix: (beam, >=beam) sorted top logprobs.
    Mutates beam_seq (L,beam) i64, beam_lp (L,beam) f32, beam_sum (beam,) f32, appends to done.
    Returns the list of source beams q per new slot, or None when no candidate is live (:480).
    ``margins`` (near-tie policy): receives this step's decision margin = the smallest
    gap between neighbours among the first beam+1 sorted candidates, i.e. by how much a candidate
    sum would have to move to change which beams survive or in which slot order."""
    cand = []
    rows = 1 if t == 1 else beam                                                  # :468-469
    for c in range(min(beam, ys.shape[1])):                                       # c OUTER :470
        for q in range(rows):                                                     # q INNER :471
            if t > 1 and int(beam_seq[t - 2, q]) == 0:                            # :475
                continue
            local = ys[q, c]
            p = (beam_sum[q] + local)                                             # fp32 add :474
            cand.append((int(ix[q, c]), q, float(p), float(local)))
    if not cand:
        return None
    cand.sort(key=lambda v: -v[2])                                                # stable :482
    if margins is not None:
        top = [v[2] for v in cand[:beam + 1]]
        margins.append(min([a - b for a, b in zip(top, top[1:])], default=float("inf")))
    prev_seq = beam_seq[:t - 1].clone()
    prev_lp = beam_lp[:t - 1].clone()
    src = []
    for v in range(min(beam, len(cand))):                                         # :491
        c, q, p, r = cand[v]
        if t > 1:
            beam_seq[:t - 1, v] = prev_seq[:, q]
            beam_lp[:t - 1, v] = prev_lp[:, q]
        src.append(q)
        beam_seq[t - 1, v] = c
        beam_lp[t - 1, v] = r
        beam_sum[v] = p
        if c == 0 or t == L:                                                      # :508
            done.append({"seq": beam_seq[:, v].clone(), "logps": beam_lp[:, v].clone(),
                         "p": float(beam_sum[v])})                                # VALUE, SURVEY D10
    return src


def sample_beam(sd: StateDict, cfg: RFNConfig, fc, att, beam_size: int = 3,
                logit_fn=None, margins_out: Optional[list] = None):
    """Per image, serial, ``beam_size`` identical rows -- the reference's own batching, kept
    because the reference is not batch-invariant.
    Returns (seq (B,L) i64, seqLogprobs (B,L) f32, top_seq list[(n_done,L)], top_prob list[list],
    reason_pred_batch).  ``margins_out`` receives one dict per image: ``steps`` = the merge decision
    margin of every step (see beam_merge) and ``final`` = best minus second-best finished beam."""
    B = fc[0].shape[0]
    L = cfg.seq_length
    assert beam_size <= cfg.V1                                                    # :360
    seq = torch.zeros(L, B, dtype=torch.int64)
    seq_lp = torch.zeros(L, B, dtype=torch.float32)
    top_seq, top_prob, reason_batch, all_done = [], [], [], []
    for k in range(B):
        fck = [f[k:k + 1].expand(beam_size, -1).contiguous() for f in fc]         # :376-386
        attk = [a[k:k + 1].expand(beam_size, -1, -1).contiguous() for a in att]
        state = get_init_state(sd, cfg, fck)
        TVc, reason_pred, st = get_thought_vectors(sd, cfg, attk, state)
        reason_batch.append(reason_pred)
        beam_seq = torch.zeros(L, beam_size, dtype=torch.int64)
        beam_lp = torch.zeros(L, beam_size, dtype=torch.float32)
        beam_sum = torch.zeros(beam_size, dtype=torch.float32)
        done: list = []
        step_margins: list = []
        lp = None
        for t in range(L + 1):
            if t == 0:
                it = torch.zeros(beam_size, dtype=torch.int64)                    # :453
            else:
                ys, ix = _topk_desc(lp, beam_size)
                src = beam_merge(beam_size, t, L, ys, ix, beam_seq, beam_lp, beam_sum, done, step_margins)
                if src is None:
                    break
                idx = torch.tensor(src, dtype=torch.int64)
                h, c = st
                h2, c2 = h.clone(), c.clone()
                h2[:len(src)] = h[idx]                                            # :499-501
                c2[:len(src)] = c[idx]
                st = (h2, c2)
                it = beam_seq[t - 1].clone()                                      # :517
            if t == L:
                break  # the reference runs one more (unused) decoder step here (:526)
            lp, st = _step_logprobs(sd, it, TVc, st)                              # :526-527
        done.sort(key=lambda d: -d["p"])                                          # stable :529
        seq[:, k] = done[0]["seq"]
        seq_lp[:, k] = done[0]["logps"]
        top_seq.append(torch.stack([d["seq"] for d in done], 0))
        top_prob.append([d["p"] for d in done])
        all_done.append(done)
        if margins_out is not None:
            margins_out.append({"steps": step_margins,
                                "final": done[0]["p"] - done[1]["p"] if len(done) > 1 else float("inf")})
    return seq.t().contiguous(), seq_lp.t().contiguous(), top_seq, top_prob, reason_batch


# --------------------------------------------------------------------------------------
# ensemble beam search     eval_utils.py:268-290 (logit mean -> log_softmax), :482-658 (beam)
# restated with the model's REAL signatures
# --------------------------------------------------------------------------------------
def ensemble_sample_beam(sds: Sequence[StateDict], cfg: RFNConfig, fc, att, beam_size: int = 3):
    B = fc[0].shape[0]
    L = cfg.seq_length
    M = len(sds)
    seq = torch.zeros(L, B, dtype=torch.int64)
    seq_lp = torch.zeros(L, B, dtype=torch.float32)
    top_seq, top_prob = [], []
    for k in range(B):
        fck = [f[k:k + 1].expand(beam_size, -1).contiguous() for f in fc]
        attk = [a[k:k + 1].expand(beam_size, -1, -1).contiguous() for a in att]
        TVcs, sts = [], []
        for sd in sds:                                                            # :514-543
            TVc, _, st = get_thought_vectors(sd, cfg, attk, get_init_state(sd, cfg, fck))
            TVcs.append(TVc)
            sts.append(st)
        beam_seq = torch.zeros(L, beam_size, dtype=torch.int64)
        beam_lp = torch.zeros(L, beam_size, dtype=torch.float32)
        beam_sum = torch.zeros(beam_size, dtype=torch.float32)
        done: list = []
        lp = None
        for t in range(L + 1):
            if t == 0:
                it = torch.zeros(beam_size, dtype=torch.int64)
            else:
                ys, ix = _topk_desc(lp, beam_size)
                src = beam_merge(beam_size, t, L, ys, ix, beam_seq, beam_lp, beam_sum, done)
                if src is None:
                    break
                idx = torch.tensor(src, dtype=torch.int64)
                for m in range(M):                                                # :604-611
                    h, c = sts[m]
                    h2, c2 = h.clone(), c.clone()
                    h2[:len(src)] = h[idx]
                    c2[:len(src)] = c[idx]
                    sts[m] = (h2, c2)
                it = beam_seq[t - 1].clone()
            if t == L:
                break
            logits = []
            for m, sd in enumerate(sds):                                          # :268-290
                lg, sts[m] = one_time_step(sd, sd["embed.weight"][it], TVcs[m], sts[m])
                logits.append(lg)
            lp = torch.log_softmax(sum(logits) / M, dim=1)                        # :282-288
        done.sort(key=lambda d: -d["p"])
        seq[:, k] = done[0]["seq"]
        seq_lp[:, k] = done[0]["logps"]
        top_seq.append(torch.stack([d["seq"] for d in done], 0))
        top_prob.append([d["p"] for d in done])
    return seq.t().contiguous(), seq_lp.t().contiguous(), top_seq, top_prob


# --------------------------------------------------------------------------------------
# A.8 / A.9  criteria                       misc/utils.py:161-192, :50-84, :292-296
# --------------------------------------------------------------------------------------
def multilabel_margin(pred: Tensor, target: Tensor) -> Tensor:
    """nn.MultiLabelMarginLoss with default mean reduction (misc/utils.py:188)."""
    return F.multilabel_margin_loss(pred, target)


def xe_loss(log_prob: Tensor, target: Tensor, mask: Tensor, top_pred: Sequence[Tensor],
            top_true: Tensor, reason_weight: float, label_smoothing: float = 0.0) -> Tensor:
    """ReviewNetEnsembleCriterion.forward (misc/utils.py:161-192); callers pass labels[:,1:],
    masks[:,1:] (train.py:155)."""
    rows, T, K = log_prob.shape
    target = target[:, :T]
    mask = mask[:, :T]
    picked = log_prob.gather(2, target.unsqueeze(2)).squeeze(2)
    if label_smoothing > 0:
        eps = label_smoothing
        per = -((1.0 - eps) * picked + (eps / K) * log_prob.sum(dim=2)) * mask    # :170-178
    else:
        per = -picked * mask                                                      # :180-184
    out = per.sum() / rows
    disc = sum(multilabel_margin(p, top_true) for p in top_pred)
    return out + disc * reason_weight / len(top_pred)                             # :186-190


def rl_loss(sample_logprobs: Tensor, seq: Tensor, reward: Tensor, logprobs_all: Tensor,
            entropy_reg: float, top_pred: Sequence[Tensor], top_true: Tensor,
            reason_weight: float) -> Tensor:
    """ReviewNetRewardCriterion.forward, non-PPO branch (misc/utils.py:50-84)."""
    rows, T = sample_logprobs.shape
    mask0 = (seq > 0).float()
    mask = torch.cat([torch.ones(rows, 1), mask0[:, :-1]], dim=1)                 # :56
    lp = logprobs_all[:, :T, :]                                                   # :59
    ent_minus = (lp * torch.exp(lp)).sum(dim=2) * mask0                           # :60-61
    out = (-sample_logprobs * reward * mask).sum() / rows + \
        entropy_reg * ent_minus.sum() / rows                                      # :70-72
    disc = sum(multilabel_margin(p, top_true) for p in top_pred)
    return out + disc * reason_weight / len(top_pred)                             # :78-82


def clip_gradient_(grads: Sequence[Tensor], grad_clip: float) -> None:
    """Element-wise clamp to +-grad_clip (misc/utils.py:292-296)."""
    for g in grads:
        g.clamp_(-grad_clip, grad_clip)


# --------------------------------------------------------------------------------------
# near-tie policy helper
# --------------------------------------------------------------------------------------
def top2_margin(lp: Tensor) -> Tensor:
    v = torch.topk(lp, 2, dim=-1)[0]
    return v[..., 0] - v[..., 1]
