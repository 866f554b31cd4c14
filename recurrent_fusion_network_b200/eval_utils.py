"""Evaluation drivers over the device path: the reference's eval_utils.eval_split (eval_utils.py:66-265) and
eval_ensemble (eval_utils.py:387-719) for the feat_array / recurrent_fusion_model configuration, same keyword
dictionary, same loader protocol (reset_iterator, get_batch, batch_size, seq_per_img, get_vocab) and the same return
value (mean loss, predictions, lang_stats).

What differs from the reference:
* a whole loader batch is decoded by ONE device beam search (the reference loops over images and beam steps in Python);
* the ensemble driver calls the real signatures (the reference's own are stale) and ignores the
  obj / mil / matching feature keys that its stale code still reads;
* language_eval needs the Java coco-caption scorers, which are outside this path: pass a callable as
  eval_kwargs['language_eval_fn'](dataset, predictions, model_id, split), or language_eval=0;
* eval_kwargs['compute_loss']=0 skips the teacher-forced loss pass of eval_split (default 1 = as the reference).
Reference quirks kept on purpose: eval_split stops after the first batch when val_images_use is left at -1
(`if n >= val_images_use: break`, eval_utils.py:249), predictions beyond the split end are popped."""
from __future__ import annotations

import numpy as np
import torch

from .ensemble import ensemble_sample_beam, ensemble_sample_greedy


def decode_sequence(ix_to_word, seq):
    """misc/utils.py:19-33: words up to the first 0 joined by blanks; one host copy for the whole batch."""
    rows = seq.detach().cpu().tolist() if torch.is_tensor(seq) else np.asarray(seq).tolist()
    out = []
    for r in rows:
        words = []
        for ix in r:
            if ix <= 0:
                break
            words.append(ix_to_word[str(int(ix))])
        out.append(" ".join(words))
    return out


def _dev(model):
    return next(model.parameters()).device


def _to(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)


def _language_eval(eval_kwargs, dataset, predictions, model_id, split):
    fn = eval_kwargs.get("language_eval_fn")
    if fn is None:
        raise RuntimeError("language_eval=1 needs the coco-caption scorers: pass eval_kwargs['language_eval_fn'] or language_eval=0")
    return fn(dataset, predictions, model_id, split)


def _pop_overrun(predictions, n, data, limit):
    ix1 = data["bounds"]["it_max"]
    if limit != -1:
        ix1 = min(ix1, limit)
    for _ in range(n - ix1):
        if predictions:
            predictions.pop()


def eval_split(model, crit, loader, eval_kwargs={}):
    verbose = eval_kwargs.get("verbose", True)
    val_images_use = eval_kwargs.get("val_images_use", -1)
    split = eval_kwargs.get("eval_split", "val")
    lang_eval = eval_kwargs.get("language_eval", 1)
    dataset = eval_kwargs.get("dataset", "coco")
    beam_size = eval_kwargs.get("beam_size", 1)
    reason_weight = eval_kwargs.get("reason_weight", 10)
    rank = eval_kwargs.get("rank", 0)
    sample_max = int(eval_kwargs.get("sample_max", 1))
    compute_loss = eval_kwargs.get("compute_loss", 1)
    if eval_kwargs.get("feature_type", "feat_array") != "feat_array":
        raise NotImplementedError("only feature_type='feat_array' (recurrent_fusion_model) is on this path")
    dev = _dev(model)
    model.eval()
    loader.reset_iterator(split)
    n, loss_sum, loss_evals, lp_sentence_sum = 0, 0.0, 0, 0.0
    predictions = []
    sampling = split in ("val", "test")
    with torch.no_grad():
        while True:
            data = loader.get_batch(split)
            n += loader.batch_size
            loss = float("nan")
            if compute_loss:
                fc = [_to(a, dev) for a in data["fc_feats_array"]]
                att = [_to(a, dev) for a in data["att_feats_array"]]
                labels, masks, top_words = (_to(data[k], dev) for k in ("labels", "masks", "top_words"))
                log_prob, top_pred = model(fc, att, labels)
                loss = float(crit(log_prob, labels[:, 1:], masks[:, 1:], top_pred, top_words, reason_weight))
                loss_sum += loss
            loss_evals += 1
            if sampling:
                pick = np.arange(loader.batch_size) * loader.seq_per_img      # one row per image
                fc = [_to(a[pick], dev) for a in data["fc_feats_array"]]
                att = [_to(a[pick], dev) for a in data["att_feats_array"]]
                res = model.sample(fc, att, {"beam_size": beam_size, "sample_max": sample_max})
                seq, seq_lp = res[0], res[1]
                lp_sentence_sum += float((seq_lp * (seq > 0).float()).sum(1).mean())
                for k, sent in enumerate(decode_sequence(loader.get_vocab(), seq)):
                    predictions.append({"image_id": data["infos"][k]["id"], "caption": sent})
                _pop_overrun(predictions, n, data, val_images_use)
            if verbose:
                print("evaluating validation performance ... %d/%d (%f)" % (data["bounds"]["it_pos_now"] - 1,
                                                                           data["bounds"]["it_max"], loss))
            if data["bounds"]["wrapped"] or n >= val_images_use:
                break
    lang_stats = None
    if lang_eval == 1:
        lang_stats = _language_eval(eval_kwargs, dataset, predictions, "eval_split_" + str(eval_kwargs.get("id", "")) + "_" + str(rank), split)
    model.train()
    if verbose:
        print("log_probs_sentence_mean:" + str(lp_sentence_sum / max(loss_evals, 1)))
    return loss_sum / max(loss_evals, 1), predictions, lang_stats


def eval_ensemble_greedy(model_list, loader, eval_kwargs={}):
    """eval_utils.py:729-975 -- what the reference's shipped eval_ensemble.sh runs (--beam_size 1)."""
    return eval_ensemble(model_list, loader, dict(eval_kwargs, beam_size=1))


def eval_ensemble(model_list, loader, eval_kwargs={}):
    """Beam search (eval_utils.py:387-719) or, with beam_size 1, greedy search (:729-975; the reference dispatches on
    beam_size in eval_ensemble.py:179-186) over the logit-mean ensemble, one device call per loader batch; predictions
    carry 'log_prob'."""
    num_images = eval_kwargs.get("num_images", -1)
    split = eval_kwargs.get("eval_split", "test")
    lang_eval = eval_kwargs.get("language_eval", 0)
    dataset = eval_kwargs.get("dataset", "coco")
    beam_size = eval_kwargs.get("beam_size", 3)
    batch_size = eval_kwargs.get("batch_size", 1)
    verbose = eval_kwargs.get("verbose", True)
    if beam_size < 1:
        raise AssertionError("beam_size not correct")
    for m in model_list:
        m.eval()
    dev = _dev(model_list[0])
    loader.reset_iterator(split)
    n = 0
    predictions = []
    with torch.no_grad():
        while True:
            data = loader.get_batch(split, batch_size)
            n += batch_size
            pick = np.arange(loader.batch_size) * loader.seq_per_img
            fc = [_to(a[pick], dev) for a in data["fc_feats_array"]]
            att = [_to(a[pick], dev) for a in data["att_feats_array"]]
            if beam_size == 1:
                seq, seq_lp = ensemble_sample_greedy(model_list, fc, att)
            else:
                seq, seq_lp = ensemble_sample_beam(model_list, fc, att, {"beam_size": beam_size})[:2]
            log_probs = (seq_lp * (seq > 0).float()).sum(1).cpu().tolist()
            for k, sent in enumerate(decode_sequence(loader.get_vocab(), seq)):
                entry = {"image_id": data["infos"][k]["id"], "caption": sent, "log_prob": log_probs[k]}
                predictions.append(entry)
                if verbose:
                    print("%s\t%s\t%s" % (entry["image_id"], entry["log_prob"], entry["caption"]))
            _pop_overrun(predictions, n, data, num_images)
            if data["bounds"]["wrapped"] or n >= num_images >= 0:
                break
    lang_stats = None
    if lang_eval == 1:
        lang_stats = _language_eval(eval_kwargs, dataset, predictions, "ensemble_" + str(eval_kwargs.get("caption_model", "")), split)
    return 0.0, predictions, lang_stats
