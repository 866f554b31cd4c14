"""CPU oracle for the CIDEr-D reward scorer.  TEST INFRASTRUCTURE ONLY.

Plain-Python restatement of cider/pyciderevalcap/ciderD/ciderD_scorer.py:114-199 (compute_cider) and of the
string conventions of get_rewards.py:20-27, 39-67 on integer token sequences: a caption is the token ids up to
AND INCLUDING the first 0 (array_to_str).  Pinned against the imported reference scorer by
oracle/gen_golden.py (fixture tests/golden/ciderd.npz)."""
import math
from collections import defaultdict

import numpy as np


def caption_tokens(arr):
    """get_rewards.py:20-27 array_to_str: tokens up to and including the first 0."""
    out = []
    for t in arr:
        out.append(int(t))
        if t == 0:
            break
    return tuple(out)


def precook(words, n=4):
    """ciderD_scorer.py:13-29: n-gram counts, insertion order = (k, first position)."""
    counts = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            counts[tuple(words[i:i + k])] += 1
    return counts


def corpus_document_frequency(refs_per_image, n=4):
    """ciderD_scorer.py:101-112 (df_mode == 'corpus'): one count per image in which the n-gram occurs.
    refs_per_image: list over scored items of lists of token tuples (as the scorer sees them: one entry
    per hypothesis, so an image scored twice counts twice, exactly as the reference does)."""
    df = defaultdict(float)
    for refs in refs_per_image:
        for ngram in set(ng for ref in refs for ng in precook(ref, n)):
            df[ngram] += 1
    return df


def ciderd_scores(hyps, refs_per_hyp, df, ref_len, n=4, sigma=6.0):
    """ciderD_scorer.py:114-199.  hyps: list of token tuples; refs_per_hyp: list (same length) of lists of token
    tuples; df: mapping n-gram tuple -> document frequency (missing = 0); ref_len = log(#documents).
    Returns np.float64 scores, one per hypothesis."""
    def counts2vec(cnts):
        vec = [dict() for _ in range(n)]
        length = 0
        norm = [0.0 for _ in range(n)]
        for ngram, tf in cnts.items():
            d = np.log(max(1.0, df.get(ngram, 0.0)))
            k = len(ngram) - 1
            vec[k][ngram] = float(tf) * (ref_len - d)
            norm[k] += pow(vec[k][ngram], 2)
            if k == 1:
                length += tf
        return vec, [np.sqrt(x) for x in norm], length

    def sim(vh, vr, nh, nr, lh, lr):
        delta = float(lh - lr)
        val = np.array([0.0 for _ in range(n)])
        for k in range(n):
            for ngram in vh[k]:
                val[k] += min(vh[k][ngram], vr[k].get(ngram, 0.0)) * vr[k].get(ngram, 0.0)
            if nh[k] != 0 and nr[k] != 0:
                val[k] /= (nh[k] * nr[k])
            val[k] *= np.e ** (-(delta ** 2) / (2 * sigma ** 2))
        return val

    scores = []
    for hyp, refs in zip(hyps, refs_per_hyp):
        vec, norm, length = counts2vec(precook(hyp, n))
        score = np.array([0.0 for _ in range(n)])
        for ref in refs:
            vr, nr, lr = counts2vec(precook(ref, n))
            score += sim(vec, vr, norm, nr, length, lr)
        s = np.mean(score)
        s /= len(refs)
        s *= 10.0
        scores.append(s)
    return np.array(scores)


def self_critical_reward(gen_result, greedy_res, gts, df, ref_len, seq_per_img, cider_weight=1.0, use_baseline=True):
    """get_rewards.py:39-112 with bleu4_weight = spice_weight = 0: reward[b, :] = CIDEr-D(sample_b) - CIDEr-D(greedy_b).
    gen_result / greedy_res: (rows, T) int arrays; gts: list over images of lists of int arrays."""
    rows = gen_result.shape[0]
    hyps = [caption_tokens(gen_result[i]) for i in range(rows)] + [caption_tokens(greedy_res[i]) for i in range(rows)]
    gt_tok = [[caption_tokens(g) for g in gts[i]] for i in range(len(gts))]
    refs = [gt_tok[(i % rows) // seq_per_img] for i in range(2 * rows)]
    if df is None:   # 'corpus' mode
        df = corpus_document_frequency(refs)
        ref_len = np.log(float(len(refs)))
    sc = ciderd_scores(hyps, refs, df, ref_len)
    out = sc[:rows] - sc[rows:] if use_baseline else sc[:rows]
    out = out * cider_weight
    return np.repeat(out[:, np.newaxis], gen_result.shape[1], 1), sc
