"""Self-critical reward on the device: the CIDEr-D scorer of
cider/pyciderevalcap/ciderD/ciderD.py + ciderD_scorer.py and the reward assembly of get_rewards.py:39-129, with the
same call signatures.  BLEU / SPICE weights must stay 0 (their scorers are outside the path: SPICE is a Java HTTP
service).  Hypotheses, references and document frequencies live on the GPU; scores are fp64."""
from __future__ import annotations

import ctypes as C
import math
from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _capi
from ._capi import check, lib, ptr, stream

MAXLEN = 32
_EMPTY = -2 ** 31


def _tokens(s: str, vocab: Dict[str, int]) -> List[int]:
    out = []
    for w in s.split():
        if w not in vocab:
            # integer tokens keep their value; other words get ids above every int32 token id a caption can carry, so the two
            # namespaces cannot collide in mixed input
            vocab[w] = (1 << 30) + len(vocab) if not w.lstrip("-").isdigit() else int(w)
        out.append(vocab[w])
    return out


class DocumentFrequency:
    """Device hash table n-gram -> document frequency (ciderD_scorer.py:69-70 loads it from data/<df>.p)."""

    def __init__(self, df: Optional[Dict[tuple, float]], n_documents: float, device):
        self.ref_len = math.log(float(n_documents))
        items = list(df.items()) if df else []
        cap = 1
        while cap < 2 * max(1, len(items)):
            cap *= 2
        self.cap = cap if items else 0
        keys = np.full((max(cap, 1), 4), -1, dtype=np.int32)
        keys[:, 0] = _EMPTY
        vals = np.zeros(max(cap, 1), dtype=np.float64)
        kbuf = (C.c_int32 * 4)()
        for ng, v in items:
            k = list(ng) + [-1] * (4 - len(ng))
            for i in range(4):
                kbuf[i] = int(k[i])
            s = lib().rfn_ciderd_hash(kbuf) & (cap - 1)
            while keys[s, 0] != _EMPTY:
                s = (s + 1) & (cap - 1)
            keys[s] = k
            vals[s] = float(v)
        self.keys = torch.from_numpy(keys).to(device)
        self.vals = torch.from_numpy(vals).to(device)


def ciderd_scores(hyp: torch.Tensor, hyp_img: torch.Tensor, refs: torch.Tensor, n_refs: torch.Tensor, df: DocumentFrequency,
                  sigma: float = 6.0) -> torch.Tensor:
    """hyp (n_hyp, Lh) int32, hyp_img (n_hyp) int32, refs (n_img, R, Lr) int32, n_refs (n_img) int32 -> (n_hyp) float64."""
    n_hyp, ld_h = hyp.shape
    n_img, R, Lr = refs.shape
    if ld_h > MAXLEN or Lr > MAXLEN:
        raise _capi.RfnError(f"captions longer than {MAXLEN} tokens are not supported")
    scratch = torch.empty(n_hyp * R * 4, dtype=torch.float64, device=hyp.device)
    scores = torch.empty(n_hyp, dtype=torch.float64, device=hyp.device)
    check(lib().rfn_ciderd_scores_f64(ptr(hyp), ld_h, n_hyp, ptr(hyp_img), ptr(refs), ptr(n_refs), R, Lr,
                                      ptr(df.keys) if df.cap else None, ptr(df.vals) if df.cap else None, df.cap,
                                      df.ref_len, float(sigma), ptr(scratch), ptr(scores), stream()), "rfn_ciderd_scores_f64")
    return scores


def pack_references(gts: Sequence[Sequence[Sequence[int]]], device) -> tuple:
    """list over images of lists of token sequences -> (refs (n_img,R,Lr) int32 zero padded, n_refs (n_img) int32)."""
    R = max(len(g) for g in gts)
    Lr = max(max(len(r) for r in g) for g in gts)
    Lr = min(max(Lr + 1, 2), MAXLEN)   # room for the terminating 0
    refs = np.zeros((len(gts), R, Lr), dtype=np.int32)
    n_refs = np.zeros(len(gts), dtype=np.int32)
    for i, g in enumerate(gts):
        n_refs[i] = len(g)
        for j, r in enumerate(g):
            r = list(r)[:Lr]
            refs[i, j, :len(r)] = r
    return torch.from_numpy(refs).to(device), torch.from_numpy(n_refs).to(device)


class CiderD:
    """cider/pyciderevalcap/ciderD/ciderD.py:14-54.  df='corpus' derives document frequencies from the references
    of the call (ciderD_scorer.py:101-112, 201-213); otherwise pass `document_frequency` (n-gram tuple of token ids ->
    count) and `n_documents` (113287 for 'coco-train', ciderD_scorer.py:172-177)."""

    def __init__(self, n=4, sigma=6.0, df="corpus", document_frequency=None, n_documents=None, device="cuda"):
        if n != 4:
            raise _capi.RfnError("CIDEr-D is built for n = 4")
        self._n, self._sigma, self._df = n, sigma, df
        self.device = torch.device(device)
        self._table = None
        if df != "corpus":
            if document_frequency is None or n_documents is None:
                raise _capi.RfnError("df != 'corpus' needs document_frequency and n_documents (the reference's data/*.p is an LFS stub)")
            self._table = DocumentFrequency(document_frequency, n_documents, self.device)

    def compute_score(self, gts, res):
        """gts: {image_id: [ref caption str, ...]}, res: [{'image_id': id, 'caption': [str]}] -> (mean, np.array)."""
        vocab: Dict[str, int] = {}
        hyps, refs_per_hyp = [], []
        for r in res:
            hypo, ref = r["caption"], gts[r["image_id"]]
            assert type(hypo) is list and len(hypo) == 1 and type(ref) is list and len(ref) > 0
            hyps.append(_tokens(hypo[0], vocab))
            refs_per_hyp.append([_tokens(x, vocab) for x in ref])
        table = self._table
        if table is None:  # corpus mode: one document per scored item (ciderD_scorer.py:101-112)
            df = defaultdict(float)
            for refs in refs_per_hyp:
                seen = set()
                for ref in refs:
                    for k in range(1, 5):
                        for i in range(len(ref) - k + 1):
                            seen.add(tuple(ref[i:i + k]))
                for ng in seen:
                    df[ng] += 1
            table = DocumentFrequency(df, len(refs_per_hyp), self.device)
        # the kernel delimits a caption by its first 0, which get_rewards.py:20-27 always writes into the string
        Lh = min(max(len(h) for h in hyps) + 1, MAXLEN)
        hyp = np.zeros((len(hyps), Lh), dtype=np.int32)
        for i, h in enumerate(hyps):
            hyp[i, :len(h)] = h[:Lh]
        refs_t, n_refs = pack_references(refs_per_hyp, self.device)
        self._check_zero_convention(hyps, refs_per_hyp)
        hyp_img = torch.arange(len(hyps), dtype=torch.int32, device=self.device)
        sc = ciderd_scores(torch.from_numpy(hyp).to(self.device), hyp_img, refs_t, n_refs, table, self._sigma).cpu().numpy()
        return float(np.mean(sc)), sc

    @staticmethod
    def _check_zero_convention(hyps, refs_per_hyp):
        """The device kernel delimits a caption by its first 0 (get_rewards.py:20-27 always writes that 0)."""
        for seq in list(hyps) + [r for refs in refs_per_hyp for r in refs]:
            if len(seq) == 0 or seq[-1] != 0 or 0 in seq[:-1]:
                raise _capi.RfnError("captions must end with the token 0 and contain no other 0 (array_to_str convention)")

    def method(self):
        return "CIDEr-D"


def compute_reward(gen_result: torch.Tensor, greedy_res: torch.Tensor, gts, table: DocumentFrequency, opt, seq_per_img=None):
    """get_rewards.py:39-112 on tensors.  gen_result / greedy_res (rows, T) token tensors on the device; gts: list over
    images of lists of token arrays (data['gts']); returns rewards (rows, T) float32 ON THE DEVICE and the fp64 scores."""
    if getattr(opt, "bleu4_weight", 0) > 0 or getattr(opt, "spice_weight", 0) > 0:
        raise NotImplementedError("BLEU-4 / SPICE rewards are outside this path (SURVEY.md section 2)")
    rows, T = gen_result.shape
    dev = gen_result.device
    seq_per_img = seq_per_img or rows // len(gts)
    # array_to_str stops after the first 0 and otherwise uses all T tokens: no terminator is appended (ld_h = T)
    hyp = torch.cat([gen_result, greedy_res], 0).to(torch.int32).contiguous()
    hyp_img = ((torch.arange(2 * rows, device=dev) % rows) // seq_per_img).to(torch.int32)
    refs, n_refs = pack_references([[list(map(int, r)) for r in g] for g in gts], dev)
    if table is None:   # CiderD(df='corpus'): documents = the 2*rows scored items (an image counted once per item)
        df = defaultdict(float)
        for i in range(2 * rows):
            seen = set()
            for r in gts[(i % rows) // seq_per_img]:
                r = _first_zero(list(map(int, r)))
                for k in range(1, 5):
                    for j in range(len(r) - k + 1):
                        seen.add(tuple(r[j:j + k]))
            for ng in seen:
                df[ng] += 1
        table = DocumentFrequency(df, 2 * rows, dev)
    return compute_reward_packed(gen_result, greedy_res, refs, n_refs, table, opt, seq_per_img)


def pack_references_static(gts, n_images: int, max_refs: int, max_len: int):
    """Host side of compute_reward for a CUDA-graph replay: references packed into FIXED-shape numpy arrays
    (n_images, max_refs, max_len) / (n_images,), to be copied into the graph's static device buffers."""
    if len(gts) != n_images:
        raise ValueError(f"expected references for {n_images} images, got {len(gts)}")
    refs = np.zeros((n_images, max_refs, max_len), dtype=np.int32)
    n_refs = np.zeros(n_images, dtype=np.int32)
    for i, g in enumerate(gts):
        if len(g) > max_refs:
            raise ValueError(f"image {i} has {len(g)} references, the graph was captured for at most {max_refs}")
        n_refs[i] = len(g)
        for j, r in enumerate(g):
            r = _first_zero(list(map(int, r)))
            if len(r) >= max_len and r[-1] != 0:
                raise ValueError(f"reference longer than {max_len - 1} tokens")
            refs[i, j, :min(len(r), max_len)] = r[:max_len]
    return refs, n_refs


def compute_reward_packed(gen_result, greedy_res, refs, n_refs, table: DocumentFrequency, opt, seq_per_img):
    """The device part of compute_reward (no host work, no synchronisation: capturable in a CUDA graph): references already
    packed on the device (pack_references), a fixed document-frequency table."""
    rows, T = gen_result.shape
    dev = gen_result.device
    hyp = torch.cat([gen_result, greedy_res], 0).to(torch.int32).contiguous()
    hyp_img = ((torch.arange(2 * rows, device=dev) % rows) // seq_per_img).to(torch.int32)
    scores = ciderd_scores(hyp, hyp_img, refs, n_refs, table)
    reward = torch.empty(rows, T, dtype=torch.float32, device=dev)
    check(lib().rfn_ciderd_reward_f32(ptr(scores), rows, T, float(getattr(opt, "cider_weight", 1.0)),
                                      1 if getattr(opt, "use_baseline", 1) else 0, ptr(reward), stream()), "rfn_ciderd_reward_f32")
    return reward, scores


def _first_zero(seq):
    out = []
    for t in seq:
        out.append(t)
        if t == 0:
            break
    return out


def get_self_critical_reward_feat_array(idx_to_word, model, fc_feat_array, att_feat_array, data, gen_result, opt, table=None,
                                        baseline_in_train_mode=False):
    """get_rewards.py:115-129: greedy baseline decode (no grad) + CIDEr-D(sample) - CIDEr-D(greedy), broadcast over T.
    Returns a numpy array like the reference (use compute_reward() to keep the rewards on the device).

    The reference decodes the baseline in whatever mode train_rl.py left the model in -- train mode, so with
    drop_prob_lm > 0 its greedy baseline is drawn through dropout.  The shipped RL scripts leave every dropout at 0
    (opts.py defaults), where the two modes are the same computation; here the baseline is decoded in eval mode (the
    fused no-tape device loop) unless baseline_in_train_mode=True asks for the reference's behaviour with dropout."""
    with torch.no_grad():
        was_training = model.training
        if not baseline_in_train_mode:
            model.eval()
        greedy_res = model.sample([f.detach() for f in fc_feat_array], [a.detach() for a in att_feat_array], {})[0]
        model.train(was_training)
    T = gen_result.shape[1]
    if greedy_res.shape[1] < T:
        greedy_res = torch.nn.functional.pad(greedy_res, (0, T - greedy_res.shape[1]))
    reward, _ = compute_reward(gen_result, greedy_res[:, :T].contiguous(), data["gts"], table, opt)
    return reward.cpu().numpy()
