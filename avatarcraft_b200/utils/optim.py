"""Flat-buffer Adam for the stylisation trainer (reference: torch.optim.Adam over
net_style.parameters(), stylize.py:355-363, stepped at :199).

The reference's step touches the 12.25 M-float parameter set through ~10 foreach launches and, in a multi-GPU step,
would need the gradients gathered into a flat buffer and scattered back.  Here parameters, gradients and both Adam
moments each live in ONE contiguous fp32 buffer (the nn.Parameters are re-pointed at views of it, the state-dict is
unchanged):

  zero_grad   one memset of the flat gradient
  backward    autograd accumulates straight into the views
  all-reduce  ONE NCCL all-reduce on the flat gradient, in place (no cat / copy-back)
  step        ONE launch of ac_adam_step over the flat buffers; slots that never received a gradient (most of the
              fine hash levels) are read and skipped

The update rule is torch.optim.Adam's (no weight decay, no amsgrad); tests/test_gpu_frame_ops.py compares it with
torch.optim.Adam step by step.
"""
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from .. import _lib


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam got no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters (no CPU path)")
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise RuntimeError("FlatAdam needs fp32 parameters on one device")
        # every parameter starts on a 16 B boundary so views stay vector-aligned
        # ... except tables of float2 entries (the hash grid): they start 8 bytes past a 16-byte boundary, where neighbouring
        # corner entries of the reference's offsets table pair up into aligned 16-byte blocks, so the training scatter can
        # merge two reductions into one red.global.add.v4.f32 (csrc/nsr_train_tc.cu: red_corner_pair)
        self._spans, at = [], 0
        for p in self.params:
            at = (at + 3) // 4 * 4
            if p.dim() == 2 and p.shape[1] == 2 and p.numel() >= (1 << 16):
                at += 2
            self._spans.append((at, p.numel()))
            at += p.numel()
        at = (at + 3) // 4 * 4
        self.numel = at
        self.flat_param = torch.zeros(at, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(at, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(at, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(at, device=dev, dtype=torch.float32)
        for p, (a, n) in zip(self.params, self._spans):
            self.flat_param[a:a + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_param[a:a + n].view_as(p)
            p.grad = self.flat_grad[a:a + n].view_as(p)
        self.lr, self.betas, self.eps, self.group = float(lr), (float(betas[0]), float(betas[1])), float(eps), group
        self.step_count = 0
        self.param_groups = [{"params": self.params, "lr": self.lr}]       # torch.optim-style view (lr schedulers)

    def zero_grad(self, set_to_none: bool = False):
        with torch.cuda.device(self.flat_grad.device):
            _lib.check(_lib.lib().ac_zero(_lib.ptr(self.flat_grad), self.numel * 4, _lib.stream_ptr()), "ac_zero")
        for p, (a, n) in zip(self.params, self._spans):                   # re-attach if autograd replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * a:
                p.grad = self.flat_grad[a:a + n].view_as(p)

    def _gather_stray_grads(self):
        for p, (a, n) in zip(self.params, self._spans):
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * a:
                self.flat_grad[a:a + n].copy_(p.grad.reshape(-1))
                p.grad = self.flat_grad[a:a + n].view_as(p)

    def all_reduce(self) -> int:
        """ONE all-reduce(SUM) of the flat gradient, in place.  Returns the number of floats reduced."""
        self._gather_stray_grads()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.numel

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        self._gather_stray_grads()
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        with torch.cuda.device(self.flat_param.device):
            _lib.check(_lib.lib().ac_adam_step(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                               _lib.ptr(self.exp_avg_sq), self.numel, lr, self.betas[0], self.betas[1], self.eps,
                                               self.step_count, float(grad_scale), _lib.stream_ptr()), "ac_adam_step")
        # the kernel wrote through raw pointers: bump the version counters the packed-MLP cache keys on
        torch.autograd.graph.increment_version(self.params)

    # ---- checkpointing (utils/checkpoint.py) ----
    def state_dict(self):
        return {"step": self.step_count, "lr": self.param_groups[0]["lr"], "betas": self.betas, "eps": self.eps,
                "exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu()}

    def load_state_dict(self, sd):
        if sd["exp_avg"].numel() != self.numel:
            raise RuntimeError("optimizer state does not match the parameter layout")
        self.step_count = int(sd["step"])
        self.param_groups[0]["lr"] = float(sd["lr"])
        self.betas, self.eps = tuple(sd["betas"]), float(sd["eps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
