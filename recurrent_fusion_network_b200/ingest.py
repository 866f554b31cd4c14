"""Feature ingest for the path: the loader side of `dataloader.py:15-29,221-356` that feeds
`model(fc_feats, att_feats, labels)`.

The reference reads, per image and per CNN encoder, one `.npy` (fc feature) and one compressed `.npz` (attention map under
the key 'feat', dataloader.py:21-29), flattens a (H, W, D) map to (H*W, D) (:243-245), REPLICATES every image's features
`seq_per_img` times in host memory (:246-247), stacks them (:344-349) and copies the replicated batch to the GPU
(train.py:116-133): 5 x 3.15 MB per image over PCIe.  Here the files are read by a thread pool that runs ahead of the
consumer, only the UNIQUE rows travel (pinned staging buffers, one non-blocking copy per encoder tensor) together with the
row -> image index, and the replication -- when a caller wants the reference's replicated layout -- happens on the device
(`rfn_expand_rows_f32`).  `FeatureBatch.expanded()` is bit-identical to what the reference would have uploaded.

Label / mask / top-word assembly (h5py, vocabulary json) stays with the caller's loader: it is out of scope."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np
import torch


def load_feat_array(image_id, fc_dirs: Sequence[str], att_dirs: Sequence[str]):
    """get_npy_feat_array (dataloader.py:21-29) + the 3-D -> 2-D flattening of get_batch (:243-245)."""
    fc, att = [], []
    for fd, ad in zip(fc_dirs, att_dirs):
        f = np.load(os.path.join(fd, str(image_id) + ".npy"))
        a = np.load(os.path.join(ad, str(image_id) + ".npz"))["feat"]
        if a.ndim == 3:
            a = a.reshape(-1, a.shape[2])
        fc.append(np.ascontiguousarray(f, dtype=np.float32))
        att.append(np.ascontiguousarray(a, dtype=np.float32))
    return fc, att, image_id


class FeatureBatch:
    """Unique feature rows of a batch + how the loader's rows map onto them."""

    def __init__(self, fc: List[torch.Tensor], att: List[torch.Tensor], image_ids, seq_per_img: int):
        self.fc, self.att = fc, att                  # J x (images, F_j), J x (images, N_j, D_j)
        self.image_ids = list(image_ids)
        self.seq_per_img = int(seq_per_img)
        self.index = torch.arange(len(self.image_ids)).repeat_interleave(self.seq_per_img)   # loader row -> unique row

    @property
    def rows(self) -> int:
        return len(self.image_ids) * self.seq_per_img

    def expanded(self):
        """(fc_feats_array, att_feats_array) exactly as the reference's get_batch stacks them (:246-247, :344-349): every
        image's row repeated seq_per_img times, consecutively."""
        g = self.seq_per_img
        if g == 1:
            return self.fc, self.att
        if self.fc[0].is_cuda:
            from ._capi import check, lib, ptr, stream
            out_fc, out_att = [], []
            for t in self.fc + self.att:
                n = t.shape[0]
                row = t[0].numel()
                o = torch.empty((n * g,) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
                check(lib().rfn_expand_rows_f32(ptr(t), g, ptr(o), n * g, row, stream()), "rfn_expand_rows_f32")
                (out_fc if len(out_fc) < len(self.fc) else out_att).append(o)
            return out_fc, out_att
        return [f.repeat_interleave(g, 0) for f in self.fc], [a.repeat_interleave(g, 0) for a in self.att]


class FeatureIngest:
    """Prefetching reader of per-image feature files.

    fc_dirs / att_dirs: one directory per encoder (the reference's fc_feat_file_list[flip_type] /
    att_feat_file_list[flip_type], dataloader.py:439-446).  `device=None` keeps the batch on the host (pinned when CUDA is
    available) -- the host logic is testable without a GPU; with a device the unique rows are staged in pinned memory and
    copied with one non-blocking transfer per tensor on `copy_stream` (default: the current stream)."""

    def __init__(self, fc_dirs: Sequence[str], att_dirs: Sequence[str], device: Optional[torch.device] = None, workers: int = 8):
        assert len(fc_dirs) == len(att_dirs) and len(fc_dirs) >= 1
        self.fc_dirs, self.att_dirs = list(fc_dirs), list(att_dirs)
        self.device = torch.device(device) if device is not None else None
        self.pool = ThreadPoolExecutor(max_workers=workers)   # np.load / zlib release the GIL
        self._pending = {}
        self._staging = {}

    def prefetch(self, image_ids):
        """Starts reading these images in the background (the reference keeps a 512-deep fifo, dataloader.py:421)."""
        for i in image_ids:
            if i not in self._pending:
                self._pending[i] = self.pool.submit(load_feat_array, i, self.fc_dirs, self.att_dirs)

    def _stage(self, key, shape):
        buf = self._staging.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
            self._staging[key] = buf
        return buf

    def get_batch(self, image_ids, seq_per_img: int = 1, copy_stream=None) -> FeatureBatch:
        """Features of `image_ids` (one unique row each); the loader's batch has seq_per_img consecutive rows per image."""
        image_ids = list(image_ids)
        self.prefetch(image_ids)
        loaded = [self._pending.pop(i).result() for i in dict.fromkeys(image_ids)]
        by_id = {r[2]: r for r in loaded}
        J = len(self.fc_dirs)
        n = len(image_ids)
        fc_out, att_out = [], []
        for j in range(J):
            f0, a0 = by_id[image_ids[0]][0][j], by_id[image_ids[0]][1][j]
            sf = self._stage(("fc", j), (n,) + f0.shape)
            sa = self._stage(("att", j), (n,) + a0.shape)
            for k, i in enumerate(image_ids):
                f, a = by_id[i][0][j], by_id[i][1][j]
                if f.shape != f0.shape or a.shape != a0.shape:
                    raise ValueError(f"image {i}, encoder {j}: feature shapes {f.shape} / {a.shape} differ from {f0.shape} / {a0.shape}")
                sf[k].copy_(torch.from_numpy(f))
                sa[k].copy_(torch.from_numpy(a))
            if self.device is None:
                fc_out.append(sf.clone())
                att_out.append(sa.clone())
            else:
                st = copy_stream or torch.cuda.current_stream(self.device)
                with torch.cuda.stream(st):
                    fc_out.append(sf.to(self.device, non_blocking=True))
                    att_out.append(sa.to(self.device, non_blocking=True))
        if self.device is not None:
            # the pinned staging buffers are reused by the next call: the copies must have left them
            (copy_stream or torch.cuda.current_stream(self.device)).synchronize()
        return FeatureBatch(fc_out, att_out, image_ids, seq_per_img)

    def close(self):
        self.pool.shutdown(wait=False, cancel_futures=True)
