"""CPU oracle of the reference's evaluation drivers.  TEST INFRASTRUCTURE ONLY (see oracle/rfnet_oracle.py).

Restates the control flow of eval_utils.py for caption_model = recurrent_fusion_model, feature_type = feat_array, on top of
the model oracle (oracle/rfnet_oracle.py, itself pinned bit-for-bit against the imported reference):

  eval_split            eval_utils.py:66-265   loss over the replicated batch, one row per image decoded, predictions popped
                                                beyond the split end, the two break conditions, mean loss
  eval_ensemble         eval_utils.py:387-719  beam search over the logit-mean ensemble, 'log_prob' = masked sum of seqLogprobs
  eval_ensemble_greedy  eval_utils.py:729-975  greedy search over the logit-mean ensemble

Pin: eval_split is checked against the REFERENCE's own eval_split, executed from its source text on the reference model
(oracle/gen_golden_eval.py; fixture tests/golden/eval_split_cases.json: prediction lists equal, loss difference 0).  The two
ensemble drivers cannot be executed: they call get_thought_vectors / one_time_step with stale signatures and read loader keys
that do not exist, and language_eval needs the Java scorers; their restatement is pinned through the MODEL
oracle only (every number it produces comes from functions that are) and by review of the cited lines -- DESIGN.md lists the
ensemble drivers as 'restated, control flow unpinned'."""
import numpy as np
import torch

from oracle import rfnet_oracle as O


def decode_sequence(ix_to_word, seq):                                             # misc/utils.py:19-33
    out = []
    for i in range(seq.shape[0]):
        txt = ""
        for j in range(seq.shape[1]):
            ix = int(seq[i, j])
            if ix > 0:
                if j >= 1:
                    txt = txt + " "
                txt = txt + ix_to_word[str(ix)]
            else:
                break
        out.append(txt)
    return out


def _pick(data, loader):
    idx = np.arange(loader.batch_size) * loader.seq_per_img                       # :169-170, :428-429, :776-777
    fc = [torch.from_numpy(a[idx]) for a in data["fc_feats_array"]]
    att = [torch.from_numpy(a[idx]) for a in data["att_feats_array"]]
    return fc, att


def eval_split(sd, cfg, loader, eval_kwargs, label_smoothing=0.0):
    val_images_use = eval_kwargs.get("val_images_use", -1)
    split = eval_kwargs.get("eval_split", "val")
    beam_size = eval_kwargs.get("beam_size", 1)
    reason_weight = eval_kwargs.get("reason_weight", 10)
    loader.reset_iterator(split)
    n, loss_sum, loss_evals = 0, 0.0, 0
    predictions = []
    with torch.no_grad():
        while True:
            data = loader.get_batch(split)
            n = n + loader.batch_size
            fc = [torch.from_numpy(a) for a in data["fc_feats_array"]]
            att = [torch.from_numpy(a) for a in data["att_feats_array"]]
            labels, masks, top_words = (torch.from_numpy(data[k]) for k in ("labels", "masks", "top_words"))
            log_prob, top_pred = O.forward_xe(sd, cfg, fc, att, labels)           # :150
            loss = float(O.xe_loss(log_prob, labels[:, 1:], masks[:, 1:], top_pred, top_words, reason_weight, label_smoothing))
            loss_sum += loss
            loss_evals += 1
            if split in ("val", "test"):
                fc1, att1 = _pick(data, loader)
                if beam_size > 1:
                    seq = O.sample_beam(sd, cfg, fc1, att1, beam_size=beam_size)[0]   # :195 -> sample_beam
                else:
                    seq = O.sample(sd, cfg, fc1, att1, sample_max=1)[0]
                for k, sent in enumerate(decode_sequence(loader.get_vocab(), seq)):   # :220-224
                    predictions.append({"image_id": data["infos"][k]["id"], "caption": sent})
            ix1 = data["bounds"]["it_max"]                                        # :239-245
            if val_images_use != -1:
                ix1 = min(ix1, val_images_use)
            for _ in range(n - ix1):
                if split in ("val", "test"):
                    predictions.pop()
            if data["bounds"]["wrapped"]:                                         # :249-252
                break
            if n >= val_images_use:
                break
    return loss_sum / loss_evals, predictions


def ensemble_sample_greedy(sds, cfg, fc, att):
    """The decode loop of eval_ensemble_greedy (:834-946) with the models' real signatures."""
    rows = fc[0].shape[0]
    M = len(sds)
    TVcs, sts = [], []
    for sd in sds:
        TVc, _, st = O.get_thought_vectors(sd, cfg, att, O.get_init_state(sd, cfg, fc))
        TVcs.append(TVc)
        sts.append(st)
    seq, slp = [], []
    lp, unfinished = None, None
    for t in range(cfg.seq_length + 1):
        if t == 0:
            it = torch.zeros(rows, dtype=torch.int64)                             # :871-875
        else:
            s_lp, it = torch.max(lp, 1)                                           # :877
        x_tok = it                                                                # :883-886: xt from the unmasked token
        if t >= 1:
            unfinished = (it > 0) if t == 1 else unfinished & (it > 0)            # :889-894
            if int(unfinished.sum()) == 0:
                break
            it = it * unfinished.to(it.dtype)
            seq.append(it)
            slp.append(s_lp)
        logits = []
        for m, sd in enumerate(sds):                                              # :268-290
            lg, sts[m] = O.one_time_step(sd, sd["embed.weight"][x_tok], TVcs[m], sts[m])
            logits.append(lg)
        lp = torch.log_softmax(sum(logits) / M, dim=1)
    return torch.stack(seq, 1), torch.stack(slp, 1)


def eval_ensemble(sds, cfg, loader, eval_kwargs):
    num_images = eval_kwargs.get("num_images", -1)
    split = eval_kwargs.get("eval_split", "test")
    beam_size = eval_kwargs.get("beam_size", 3)
    batch_size = eval_kwargs.get("batch_size", 1)
    loader.reset_iterator(split)
    n = 0
    predictions = []
    with torch.no_grad():
        while True:
            data = loader.get_batch(split, batch_size)
            n = n + batch_size
            fc, att = _pick(data, loader)
            if beam_size == 1:
                seq, seq_lp = ensemble_sample_greedy(sds, cfg, fc, att)
            else:
                seq, seq_lp = O.ensemble_sample_beam(sds, cfg, fc, att, beam_size=beam_size)[:2]
            log_probs = torch.sum(seq_lp * (seq > 0).float(), 1)                  # :660, :948
            for k, sent in enumerate(decode_sequence(loader.get_vocab(), seq)):
                predictions.append({"image_id": data["infos"][k]["id"], "caption": sent, "log_prob": float(log_probs[k])})
            ix1 = data["bounds"]["it_max"]                                        # :700-705, :958-963
            if num_images != -1:
                ix1 = min(ix1, num_images)
            for _ in range(n - ix1):
                predictions.pop()
            if data["bounds"]["wrapped"]:
                break
            if n >= num_images >= 0:
                break
    return predictions
