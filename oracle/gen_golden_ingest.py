"""Writes tests/golden/ingest_case.npz: the reference's OWN feature-ingest code run on small synthetic feature files.

dataloader.py cannot be imported in this container (h5py absent, np.int removed), so the two pieces of it that touch the
features are executed from their SOURCE TEXT, unmodified: `get_npy_feat_array` (dataloader.py:21-29) and the per-encoder
flatten + replicate loop / np.stack of DataLoader.get_batch (:247-254 and :336-340 of the file as numbered here), driven with
the local variables those lines use.  The fixture stores the input files' contents and the reference's output arrays; the
oracle restatement (oracle/ingest_oracle.py) and the product ingest (recurrent_fusion_network_b200/ingest.py) are both checked
against it.  Run from the repo root:  python oracle/gen_golden_ingest.py"""
import os
import sys
import tempfile
import textwrap

import numpy as np

REF = "/root/reference/dataloader.py"
lines = open(REF).read().split("\n")


def grab(first, last):
    return "\n".join(lines[first - 1:last])


def find(text, start=0):
    for i in range(start, len(lines)):
        if text in lines[i]:
            return i + 1
    raise SystemExit("reference line not found: " + text)


ns = {"np": np}
a = find("def get_npy_feat_array")
exec(grab(a, a + 8), ns)                                   # the reference's loader function, verbatim
loop_a = find("for feat_id in range(self.num_feat_array):")
loop_src = textwrap.dedent(grab(loop_a, loop_a + 5)).replace("self.num_feat_array", "num_feat_array")
stack_a = find("fc_data_all = []")
stack_src = textwrap.dedent(grab(stack_a, stack_a + 4)).replace("self.num_feat_array", "num_feat_array")
print(loop_src)
print(stack_src)

rng = np.random.RandomState(5)
J, seq_per_img = 3, 5
shapes = [((7, 7, 16), 24), ((36, 12), 12), ((2, 3, 8), 8)]     # (att map shape: 3-D or already 2-D, fc size)
ids = [391895, 522418, 184613, 318219]
tmp = tempfile.mkdtemp()
fc_dirs = [os.path.join(tmp, f"fc{j}") for j in range(J)]
att_dirs = [os.path.join(tmp, f"att{j}") for j in range(J)]
files = {}
for j in range(J):
    os.makedirs(fc_dirs[j]); os.makedirs(att_dirs[j])
    for i in ids:
        f = rng.randn(shapes[j][1]).astype(np.float32)
        t = rng.randn(*shapes[j][0]).astype(np.float32)
        np.save(os.path.join(fc_dirs[j], f"{i}.npy"), f)
        np.savez_compressed(os.path.join(att_dirs[j], f"{i}.npz"), feat=t)
        files[f"fc_{j}_{i}"] = f
        files[f"att_{j}_{i}"] = t

batch = [ids[2], ids[0], ids[3]]
num_feat_array = J
fc_batch = [[] for _ in range(J)]
att_batch = [[] for _ in range(J)]
for image_id in batch:
    fcs = [os.path.join(d, str(image_id) + ".npy") for d in fc_dirs]
    atts = [os.path.join(d, str(image_id) + ".npz") for d in att_dirs]
    tmp_fc_feat_array, tmp_att_feat_array, _ = ns["get_npy_feat_array"](image_id, fcs, atts)
    exec(loop_src)                                          # the reference's flatten + replicate lines, verbatim
data = {}
exec(stack_src)                                             # the reference's np.stack lines, verbatim
out = dict(files)
out["batch_ids"] = np.array(batch)
out["all_ids"] = np.array(ids)
out["seq_per_img"] = np.array(seq_per_img)
out["J"] = np.array(J)
for j in range(J):
    out[f"ref_fc_{j}"] = fc_data_all[j]
    out[f"ref_att_{j}"] = att_data_all[j]
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ingest_case.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items() if k.startswith("ref_")})
