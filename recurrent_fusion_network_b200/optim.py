"""FusedAdam: torch.optim.Adam's update (train.py:56: Adam(lr, weight_decay), no amsgrad) with the reference's
clip_gradient (misc/utils.py:292-296) folded in, as one HBM pass per step through rfn_adam_step_f32.
Same constructor arguments / param_groups / state conventions as torch.optim.Adam, so misc/utils.set_lr and the
reference's checkpoint code keep working.

grad_scale (1 / world_size) turns the SUM of dist.average_gradients(..., divide=False) into the mean inside the same
pass, before the clamp, exactly where the reference's averaged gradient would be clamped.

capturable=True keeps {step, lr} of each group in a device tensor that the kernel reads, so that step() can be
captured in a CUDA graph (training.GraphedXEStep): the update count advances on the device and `sync_lr()` pushes a
changed param_group['lr'] before the next replay."""
import ctypes as C

import torch

from . import _capi
from ._capi import check, lib, ptr, ptr_array, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clip=0.0, capturable=False,
                 grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_clip=grad_clip, capturable=capturable,
                        grad_scale=grad_scale)
        super().__init__(params, defaults)
        self._hyper_t = {}     # group index -> (device {step, lr}, device {1, 0}); kept out of param_groups / state_dict
        self._begun = False    # begin_step() already advanced the update count of the step in progress
        self._done = set()     # id(p) of the parameters apply_to() has already updated in the step in progress
        self._group_of = None

    def _hyper(self, gi, group):
        if gi not in self._hyper_t:
            dev = group["params"][0].device
            steps = [int(self.state[p]["step"]) for p in group["params"] if self.state.get(p)]
            self._hyper_t[gi] = (torch.tensor([float(max(steps, default=0)), float(group["lr"])], dtype=torch.float32, device=dev),
                                 torch.tensor([1.0, 0.0], dtype=torch.float32, device=dev))
        return self._hyper_t[gi]

    def init_state(self):
        """Allocates exp_avg / exp_avg_sq (and the device {step, lr}) now instead of at the first step(): a CUDA-graph
        capture of step() must not contain the zero-fills."""
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                st = self.state[p]
                if p.requires_grad and not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if group.get("capturable"):
                self._hyper(gi, group)

    def state_dict(self):
        """capturable mode: graph replays advance the update count on the device only; bring the host mirror up to date
        before the state is saved (a checkpoint then resumes with the right bias correction)."""
        for gi, group in enumerate(self.param_groups):
            if gi in self._hyper_t:
                step = int(self._hyper_t[gi][0][0].item())
                for p in group["params"]:
                    if self.state.get(p):
                        self.state[p]["step"] = step
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._hyper_t = {}     # re-created from the loaded update count at the next step

    def sync_lr(self):
        """capturable mode: copy each group's lr to the device (call after misc/utils.set_lr)."""
        for gi, group in enumerate(self.param_groups):
            if group.get("capturable"):
                self._hyper(gi, group)[0][1] = float(group["lr"])

    @torch.no_grad()
    def begin_step(self):
        """Advances the update count (bias correction) of the step that starts now.  step() does this itself; a caller that
        applies part of the update early -- apply_to(), from inside the backward pass -- calls it first."""
        if self._begun:
            return
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                st = self.state[p]
                if p.requires_grad and not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if group.get("capturable"):
                hyper, one = self._hyper(gi, group)   # created from the update count BEFORE this step
                hyper.add_(one)                        # step += 1 on the device (captured with the graph)
            for p in group["params"]:
                if self.state.get(p):
                    self.state[p]["step"] += 1         # host mirror; not advanced by graph replays
        self._begun = True

    def _launch(self, gi, group, ps, grads):
        for p, g in zip(ps, grads):
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous() and g.dtype == torch.float32):
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
        hyper = self._hyper(gi, group)[0] if group.get("capturable") else None
        step = self.state[ps[0]]["step"]
        n = len(ps)
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        b1, b2 = group["betas"]
        check(lib().rfn_adam_step_f32(n, ptr_array(ps), ptr_array(grads),
                                      ptr_array([self.state[p]["exp_avg"] for p in ps]),
                                      ptr_array([self.state[p]["exp_avg_sq"] for p in ps]), numel, float(group["lr"]),
                                      float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                      float(group.get("grad_clip", 0.0)), float(group.get("grad_scale", 1.0)), int(step),
                                      ptr(hyper), stream()),
              "rfn_adam_step_f32")

    @torch.no_grad()
    def apply_to(self, pairs):
        """The update of some parameters from explicitly given, final gradients, on the CURRENT stream, before step(): the
        hand-scheduled stage-1 backward (tape.Stage1Fn) finishes the weight gradients of a fusion step long before the
        backward pass ends, and those 95 % of the parameters are touched by nothing after it, so their optimizer pass
        overlaps the rest of the backward pass.  begin_step() must have been called; step() then skips them."""
        if not self._begun:
            raise RuntimeError("FusedAdam.apply_to: call begin_step() first")
        if self._group_of is None:
            self._group_of = {id(p): gi for gi, group in enumerate(self.param_groups) for p in group["params"]}
        by_group = {}
        for p, g in pairs:
            gi = self._group_of.get(id(p))
            if gi is None:
                # the tape hands over the tensors autograd saved: the same storage as the registered Parameter
                gi = self._by_ptr().get(p.data_ptr())
                if gi is None:
                    continue
                p = self._ptr_param[p.data_ptr()]
            by_group.setdefault(gi, []).append((p, g))
        for gi, items in by_group.items():
            self._launch(gi, self.param_groups[gi], [p for p, _ in items], [g for _, g in items])
            for p, _ in items:
                self._done.add(id(p))

    def _by_ptr(self):
        if getattr(self, "_ptr_group", None) is None:
            self._ptr_group, self._ptr_param = {}, {}
            for gi, group in enumerate(self.param_groups):
                for p in group["params"]:
                    self._ptr_group[p.data_ptr()] = gi
                    self._ptr_param[p.data_ptr()] = p
        return self._ptr_group

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.begin_step()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None and id(p) not in self._done]
            if not ps:
                continue
            self._launch(gi, group, ps, [p.grad for p in ps])
        self._begun = False
        self._done = set()
        _capi.WEIGHTS_EPOCH[0] += 1   # the kernel wrote the parameters through raw pointers: invalidate derived weight caches
        return loss
