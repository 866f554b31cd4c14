"""Multi-GPU plumbing of the path: one process per GPU, images sharded contiguously, weights replicated.

Inference needs no collective besides the final caption gather.  Training (when the
gradient path is used) is plain data parallelism: every loss term of the reference is normalised by the
LOCAL row count (misc/utils.py:72,177,184,188), so with equal rows per rank the MEAN of the rank gradients
equals the single-process gradient on the concatenated batch; the element-wise clamp of
misc/utils.py:292-296 must be applied AFTER the average."""
import os
from typing import Callable, Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [k0, k1) of n_total images for `rank` (last ranks may be short or empty)."""
    per = (n_total + world - 1) // world
    k0 = min(n_total, rank * per)
    return k0, min(n_total, k0 + per)


def gather_captions(seq: torch.Tensor, seq_logprobs: torch.Tensor, n_total: int, group=None):
    """All-gather the per-shard captions into (n_total, L) tensors on every rank.

    seq (n_local, L) int64 and seq_logprobs (n_local, L) float32 live on the backend's device (CUDA for
    NCCL, CPU for gloo).  Ragged shards are padded to the common per-rank size and trimmed again."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return seq, seq_logprobs
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    L = seq.shape[1]
    n_local = seq.shape[0]
    pad_s = torch.zeros(per, L, dtype=seq.dtype, device=seq.device)
    pad_l = torch.zeros(per, L, dtype=seq_logprobs.dtype, device=seq.device)
    pad_s[:n_local] = seq
    pad_l[:n_local] = seq_logprobs
    out_s = torch.empty(world * per, L, dtype=seq.dtype, device=seq.device)
    out_l = torch.empty(world * per, L, dtype=seq_logprobs.dtype, device=seq.device)
    dist.all_gather_into_tensor(out_s, pad_s, group=group)
    dist.all_gather_into_tensor(out_l, pad_l, group=group)
    return out_s[:n_total], out_l[:n_total]


def sharded_decode(decode: Callable, fc_feats: Sequence[torch.Tensor], att_feats: Sequence[torch.Tensor],
                   group=None, device: Optional[torch.device] = None):
    """Decode this rank's contiguous shard of the images with `decode(fc_shard, att_shard) -> (seq, seq_logprobs)`
    and gather every rank's captions.  `fc_feats` / `att_feats` hold ALL images (e.g. host arrays); only the
    local shard is touched."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_total = fc_feats[0].shape[0]
    k0, k1 = shard_range(n_total, world, rank)
    if k1 > k0:
        seq, slp = decode([f[k0:k1] for f in fc_feats], [a[k0:k1] for a in att_feats])
    else:
        seq = slp = None
    # greedy / multinomial decodes return seq[:, :T] with a rank-dependent T (the early break of
    # misc/RecurrentFusionModel.py:645): every rank pads to the widest shard before the gather
    L = _max_width(0 if seq is None else seq.shape[1], group)
    if seq is None:  # empty shard: contribute zero rows of the right width
        dev = device or torch.device("cpu")
        seq = torch.zeros(0, L, dtype=torch.int64, device=dev)
        slp = torch.zeros(0, L, dtype=torch.float32, device=dev)
    elif seq.shape[1] < L:
        seq = torch.nn.functional.pad(seq, (0, L - seq.shape[1]))
        slp = torch.nn.functional.pad(slp, (0, L - slp.shape[1]))
    return gather_captions(seq, slp, n_total, group)


def _max_width(L, group):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return L
    objs = [None] * dist.get_world_size(group)
    dist.all_gather_object(objs, int(L), group=group)
    return max(objs)


def average_gradients(params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 256 << 20,
                      grad_clip: Optional[float] = None, divide: bool = True) -> None:
    """All-reduce of the gradients (sum, then / world unless divide=False), then the reference's element-wise clamp.
    1.958 GB of fp32 gradients per step for the full model.
    NCCL: the gradient tensors are reduced IN PLACE as coalesced groups (one ncclGroup per <= 256 tensors): no
    flatten / copy-back passes over HBM.  divide=False leaves the SUM for FusedAdam(grad_scale=1/world), which folds the
    division (and the clamp) into the optimizer pass.  Other backends (gloo in the CPU tests): flat buckets."""
    params = list(params)   # iterated more than once below: a generator (model.parameters()) must not be exhausted
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if grad_clip is not None:
            for p in params:
                if p.grad is not None:
                    p.grad.clamp_(-grad_clip, grad_clip)
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if grads and grads[0].is_cuda and dist.get_backend(group) == "nccl" and not os.environ.get("RFN_FLAT_ALLREDUCE"):
        from torch.distributed.distributed_c10d import _coalescing_manager
        for i in range(0, len(grads), 256):
            with _coalescing_manager(group=group, device=grads[0].device):
                for g in grads[i:i + 256]:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        if divide:
            torch._foreach_div_(grads, float(world))
        if grad_clip is not None:
            torch._foreach_clamp_min_(grads, -grad_clip)
            torch._foreach_clamp_max_(grads, grad_clip)
        return
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if divide:
            flat.div_(world)
        if grad_clip is not None:
            flat.clamp_(-grad_clip, grad_clip)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for p in params:
        if p.grad is None:
            continue
        bucket.append(p.grad)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


class OverlappedGradSync:
    """Gradient all-reduce that overlaps the backward pass: post-accumulate-grad hooks collect the gradients as autograd
    finishes them (decoder first, stage-1 steps last) and, every `bucket_bytes`, launch one coalesced in-place NCCL
    all-reduce (SUM) on a side stream that waits only for the work issued so far.  finish() flushes the rest and joins
    the side stream.  Usable eagerly or inside a CUDA-graph capture (training.GraphedXEStep), where the fork / join
    become graph dependencies and the collectives are replayed with the graph.  The mean is left to
    FusedAdam(grad_scale=1/world)."""

    def __init__(self, params, group=None, bucket_bytes: int = 96 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.bucket_bytes = group, int(bucket_bytes)
        self.side = torch.cuda.Stream()
        self.pending, self.pending_bytes, self.handles = [], 0, []
        self.buckets_launched = 0
        self._early_done = set()

    def install(self):
        self.remove()
        for p in self.params:
            self.handles.append(p.register_post_accumulate_grad_hook(self._hook))
        from . import tape
        self._early_done = set()
        tape.GRAD_SINK = self

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []
        from . import tape
        if tape.GRAD_SINK is self:
            tape.GRAD_SINK = None

    # tape.GRAD_SINK protocol: the hand-scheduled stage-1 backward (tape.Stage1Fn) hands over each fusion step's weight
    # gradients as soon as they are issued, long before autograd sees them (it returns all 400 tensors at the very end)
    def early(self, pairs, streams):
        from torch.distributed.distributed_c10d import _coalescing_manager
        grads = [g for _, g in pairs]
        if not grads:
            return
        for st in streams:
            self.side.wait_stream(st)
        with torch.cuda.stream(self.side):
            for i in range(0, len(grads), 256):
                with _coalescing_manager(group=self.group, device=grads[0].device):
                    for g in grads[i:i + 256]:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        for p, _ in pairs:
            self._early_done.add(p.data_ptr())
        self.buckets_launched += 1

    def join(self):
        torch.cuda.current_stream().wait_stream(self.side)

    def _hook(self, p):
        if p.data_ptr() in self._early_done:      # reduced in place by early(); p.grad holds (or is a copy of) the sum
            self._early_done.discard(p.data_ptr())
            return
        self.pending.append(p.grad)
        self.pending_bytes += p.grad.numel() * p.grad.element_size()
        if self.pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self.pending:
            return
        from torch.distributed.distributed_c10d import _coalescing_manager
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            for i in range(0, len(self.pending), 256):
                with _coalescing_manager(group=self.group, device=self.pending[0].device):
                    for g in self.pending[i:i + 256]:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        self.pending, self.pending_bytes = [], 0
        self.buckets_launched += 1

    def finish(self):
        self._flush()
        torch.cuda.current_stream().wait_stream(self.side)
