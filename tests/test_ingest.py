"""Feature ingest (dataloader.py:15-29, 221-356): the oracle restatement and the product loader against
the fixture written from the reference's own source text (oracle/gen_golden_ingest.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_case.npz")


@pytest.fixture()
def case(tmp_path):
    d = np.load(GOLD)
    J, ids = int(d["J"]), [int(i) for i in d["all_ids"]]
    fc_dirs = [str(tmp_path / f"fc{j}") for j in range(J)]
    att_dirs = [str(tmp_path / f"att{j}") for j in range(J)]
    for j in range(J):
        os.makedirs(fc_dirs[j]); os.makedirs(att_dirs[j])
        for i in ids:
            np.save(os.path.join(fc_dirs[j], f"{i}.npy"), d[f"fc_{j}_{i}"])
            np.savez_compressed(os.path.join(att_dirs[j], f"{i}.npz"), feat=d[f"att_{j}_{i}"])
    return d, J, fc_dirs, att_dirs, [int(i) for i in d["batch_ids"]], int(d["seq_per_img"])


def test_oracle_matches_reference_fixture(case):
    d, J, fc_dirs, att_dirs, batch, spi = case
    fc, att = IO.get_batch_features(batch, fc_dirs, att_dirs, spi)
    for j in range(J):
        assert np.array_equal(fc[j], d[f"ref_fc_{j}"]) and np.array_equal(att[j], d[f"ref_att_{j}"])


def test_host_ingest_unique_rows_plus_index_reproduce_the_replicated_batch(case):
    """No GPU: unique rows + index == the reference's replicated, stacked batch (bit-exact), incl. repeated images, a ragged
    batch, seq_per_img = 1, prefetch ahead of use and the 3-D -> 2-D flattening."""
    from recurrent_fusion_network_b200.ingest import FeatureIngest
    d, J, fc_dirs, att_dirs, batch, spi = case
    ing = FeatureIngest(fc_dirs, att_dirs, device=None, workers=2)
    ing.prefetch(batch)
    fb = ing.get_batch(batch, spi)
    assert fb.rows == len(batch) * spi and fb.index.tolist() == [k for k in range(len(batch)) for _ in range(spi)]
    fce, atte = fb.expanded()
    for j in range(J):
        assert fb.fc[j].shape[0] == len(batch)                                   # only unique rows are held / shipped
        assert np.array_equal(fb.fc[j][fb.index].numpy(), d[f"ref_fc_{j}"])
        assert np.array_equal(fb.att[j][fb.index].numpy(), d[f"ref_att_{j}"])
        assert np.array_equal(fce[j].numpy(), d[f"ref_fc_{j}"]) and np.array_equal(atte[j].numpy(), d[f"ref_att_{j}"])
    ids = [int(i) for i in d["all_ids"]]
    for b, g in (([ids[1], ids[1], ids[0]], 2), ([ids[3]], 1), (ids, 3)):
        fb = ing.get_batch(b, g)
        want_fc, want_att = IO.get_batch_features(b, fc_dirs, att_dirs, g)
        fce, atte = fb.expanded()
        for j in range(J):
            assert np.array_equal(fce[j].numpy(), want_fc[j]) and np.array_equal(atte[j].numpy(), want_att[j])
    with pytest.raises(FileNotFoundError):
        ing.get_batch([12345], 1)
    ing.close()


@pytest.mark.gpu
def test_device_ingest_expand_matches_reference_fixture(case):
    """GPU: unique rows over PCIe from pinned staging, replication on the device == the reference's uploaded batch."""
    from recurrent_fusion_network_b200.ingest import FeatureIngest
    d, J, fc_dirs, att_dirs, batch, spi = case
    ing = FeatureIngest(fc_dirs, att_dirs, device=torch.device("cuda", 0))
    fb = ing.get_batch(batch, spi)
    assert all(t.is_cuda for t in fb.fc + fb.att)
    fce, atte = fb.expanded()
    torch.cuda.synchronize()
    for j in range(J):
        assert np.array_equal(fce[j].cpu().numpy(), d[f"ref_fc_{j}"]) and np.array_equal(atte[j].cpu().numpy(), d[f"ref_att_{j}"])
    fb2 = ing.get_batch(batch[::-1], spi)          # staging buffers are reused: the first batch must be intact
    for j in range(J):
        assert np.array_equal(fb.fc[j].cpu().numpy(), d[f"ref_fc_{j}"][::spi])
    ing.close()
