"""``models.setup(opt)`` -- same factory contract as the reference's models.py:14-38."""
import os

import torch

from .model import RecurrentFusionModel


def setup(opt):
    if opt.caption_model == "recurrent_fusion_model":
        model = RecurrentFusionModel(opt)
    elif opt.caption_model in ("show_tell", "review_net"):
        # other model families of the reference are outside this path
        raise Exception("Caption model not built in recurrent_fusion_network_b200: {}".format(opt.caption_model))
    else:
        raise Exception("Caption model not supported: {}".format(opt.caption_model))

    # check compatibility if training is continued from previously saved model (models.py:25-36)
    if vars(opt).get("start_from", None) is not None:
        assert os.path.isdir(opt.start_from), " %s must be a a path" % opt.start_from
        assert os.path.isfile(os.path.join(opt.start_from, "infos_" + opt.load_model_id + ".pkl")), \
            "infos.pkl file does not exist in path %s" % opt.start_from
        model.load_state_dict(torch.load(os.path.join(opt.start_from, "model_" + opt.load_model_id + ".pth"),
                                         map_location="cpu"))
    return model
