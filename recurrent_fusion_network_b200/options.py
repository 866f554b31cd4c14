"""A minimal ``opt`` namespace holding the ~20 fields the model reads from the reference's
opts.parse_opt() (opts.py; misc/RecurrentFusionModel.py:120-151).  The reference's own ``opt``
object works unchanged; this helper exists for programmatic use and tests."""
from types import SimpleNamespace

# feat_array.py:6-9, :240-244
FULL_FEAT_ARRAY_INFO = [
    dict(fc_feat_size=2048, att_feat_size=2048, att_num=196),   # resnet
    dict(fc_feat_size=1536, att_feat_size=1536, att_num=64),    # inception_v4
    dict(fc_feat_size=2048, att_feat_size=1280, att_num=64),    # inception_v3
    dict(fc_feat_size=2208, att_feat_size=2208, att_num=49),    # densenet
    dict(fc_feat_size=1536, att_feat_size=1536, att_num=64),    # inception_resnet_v2
]


def make_opt(feat_array_info=None, **kw):
    opt = SimpleNamespace(
        caption_model="recurrent_fusion_model", vocab_size=9487, input_encoding_size=512, rnn_type="lstm",
        rnn_size=512, num_layers=1, drop_prob_lm=0.0, drop_prob_reason=0.0, drop_prob_fusion=0.0, seq_length=16,
        num_review_steps=8, num_review_steps_0=8, top_words_count=1000, att_hid_size=512, review_maxout=0,
        maxout=0, fusion_maxout=0, use_cuda=1, start_from=None, load_model_id="",
        feat_array_info=[dict(f) for f in (feat_array_info or FULL_FEAT_ARRAY_INFO)])
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt
