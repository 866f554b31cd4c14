"""CPU oracle of the reference's feature ingest.  TEST INFRASTRUCTURE ONLY (see oracle/rfnet_oracle.py).

Restates, in numpy, what `dataloader.py` hands to train.py for the feat_array feature type:
  get_npy_feat_array           dataloader.py:21-29    np.load(fc .npy), np.load(att .npz)['feat'] per encoder
  DataLoader.get_batch         dataloader.py:241-247  flatten a 3-D attention map, replicate each image seq_per_img times
                               dataloader.py:344-349  np.stack per encoder
The reference module itself cannot be imported here (h5py is absent and it uses np.int, removed in numpy 2),
so this restatement is pinned by the fixture tests/golden/ingest_case.npz written by oracle/gen_golden_ingest.py, which runs
the reference's own `get_npy_feat_array` SOURCE TEXT (extracted from /root/reference/dataloader.py) on the same files."""
import os

import numpy as np


def get_npy_feat_array(image_id, fc_file_list, att_file_list):                    # dataloader.py:21-29
    fc_feat, att_feat = [], []
    for i in range(len(fc_file_list)):
        fc_feat.append(np.load(fc_file_list[i]))
        att_feat.append(np.load(att_file_list[i])["feat"])
    return fc_feat, att_feat, image_id


def get_batch_features(image_ids, fc_dirs, att_dirs, seq_per_img):
    """-> (fc_feats_array, att_feats_array): per encoder (batch * seq_per_img, F) and (batch * seq_per_img, N, D)."""
    J = len(fc_dirs)
    fc_batch = [[] for _ in range(J)]
    att_batch = [[] for _ in range(J)]
    for image_id in image_ids:
        fc_files = [os.path.join(a, str(image_id) + ".npy") for a in fc_dirs]      # :442-446
        att_files = [os.path.join(a, str(image_id) + ".npz") for a in att_dirs]
        tmp_fc, tmp_att, _ = get_npy_feat_array(image_id, fc_files, att_files)
        for feat_id in range(J):
            a = tmp_att[feat_id]
            if len(a.shape) == 3:                                                  # :243-245
                a = a.reshape(-1, a.shape[2])
            att_batch[feat_id] += [a] * seq_per_img                                # :246
            fc_batch[feat_id] += [tmp_fc[feat_id]] * seq_per_img                   # :247
    return [np.stack(f) for f in fc_batch], [np.stack(a) for a in att_batch]       # :344-349
