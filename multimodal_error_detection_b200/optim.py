"""Fused Adam over flat parameter / gradient buffers (torch.optim.Adam semantics).

Replaces ``torch.optim.Adam(FE.params + model.params, lr, weight_decay)`` of the reference
(MED/modeling/modeling_utils.py:221-222): L2-coupled weight decay, betas (0.9, 0.999), eps 1e-8.
All parameters are re-homed as views into ONE contiguous fp32 buffer (and their ``.grad`` into one
gradient buffer), so a step is one kernel launch and the data-parallel gradient exchange is one
all-reduce of one buffer.  It subclasses ``torch.optim.Optimizer`` so the stock
``CosineAnnealingLR`` the reference uses (modeling_utils.py:257-258) drives ``param_groups[0]['lr']``.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat_ready = False
        self.grad_scale = 1.0          # set to 1/world_size when gradients were SUM-all-reduced
        self._lr_on_device = None

    def _flatten(self):
        ps = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("b200med FusedAdam needs CUDA parameters (no CPU fallback)")
        # 16-byte aligned segments so that the kernel's float4 path and tensor views both work
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(ps, offs):
            n = p.numel()
            self.flat_param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + n].view_as(p.data)
            gview = self.flat_grad[o:o + n].view_as(p.data)
            if p.grad is not None:
                gview.copy_(p.grad)
            p.grad = gview
        self.state_dev = torch.zeros(4, dtype=torch.float32, device=dev)   # {step, lr, bc1, sqrt(bc2)}
        self.n_params = sum(p.numel() for p in ps)
        self._flat_ready = True

    def prepare(self):
        if not self._flat_ready:
            self._flatten()
        return self

    def zero_grad(self, set_to_none: bool = False):
        # gradients live in the flat buffer; keep the views, zero the storage (one memset)
        self.prepare()
        self.flat_grad.zero_()

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.state_dev[1:2].fill_(lr)
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self, closure=None):
        self.prepare()
        g = self.param_groups[0]
        self.sync_lr()
        b1, b2 = g["betas"]
        ops.adam_advance(self.state_dev, b1, b2)
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.state_dev, b1, b2, g["eps"],
                      g["weight_decay"], self.grad_scale)
