"""Fused Adam over flat parameter / gradient buffers (torch.optim.Adam semantics).

Replaces ``torch.optim.Adam(FE.params + model.params, lr, weight_decay)`` of the reference
(MED/modeling/modeling_utils.py:221-222): L2-coupled weight decay, betas (0.9, 0.999), eps 1e-8.
Parameters are re-homed as views into a few contiguous fp32 CHUNKS (and their ``.grad`` into matching
gradient chunks), so a step is one kernel launch per chunk and the data-parallel gradient exchange is
one all-reduce per chunk:

* chunk 0 packs every ordinary parameter (16-byte aligned segments);
* a module that already keeps its weights in ONE storage of its own -- ``nn.LSTM`` after
  ``flatten_parameters()``, whose cuDNN kernels need that exact layout -- is adopted in place as a
  further chunk, so cuDNN never has to re-compact the weights.

Like ``torch.optim.Adam``, a parameter that has never received a gradient is skipped entirely (no
decay, no moments): the reference registers an unused FeatureExtractor when ``video_dims == 2048``
(SURVEY Appendix A-11).  Two documented differences from ``torch.optim.Adam``: there is ONE step counter / bias correction
for all parameters (torch keeps one per parameter; identical whenever every trained parameter receives a gradient from the
first step on, which holds for every reference configuration), and a parameter that has been active once keeps decaying
with a zero gradient in a step where its ``.grad`` is None (torch would skip it; the reference's loops never produce that).
``state_dict()`` / ``load_state_dict()`` carry the flat moments and the step scalars.  The class subclasses ``torch.optim.Optimizer`` so the stock
``CosineAnnealingLR`` of the reference (modeling_utils.py:257-258) drives ``param_groups[0]['lr']``.
"""
from __future__ import annotations

from typing import List

import torch

from . import ops


class _Chunk:
    """One contiguous fp32 parameter buffer with its gradient / moment twins."""

    def __init__(self, param_buf: torch.Tensor):
        self.param = param_buf
        self.grad = torch.zeros_like(param_buf)
        self.exp_avg = torch.zeros_like(param_buf)
        self.exp_avg_sq = torch.zeros_like(param_buf)
        self.members = []        # (param, offset, numel)
        self.runs = []           # active [begin, end) ranges, multiples of 4 elements


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat_ready = False
        self._active = None            # per-parameter "has received a gradient" flags, known after the first step
        self.grad_scale = 1.0          # set to 1/world_size when gradients were SUM-all-reduced
        self._lr_on_device = None
        self.chunks: List[_Chunk] = []

    # ------------------------------------------------------------------------------------ layout
    def _flatten(self):
        ps = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("b200med FusedAdam needs CUDA parameters (no CPU fallback)")
        # parameters that already share one storage (cuDNN-flattened LSTM weights) are adopted in place
        by_storage = {}
        for p in ps:
            by_storage.setdefault(p.data.untyped_storage().data_ptr(), []).append(p)
        packed, self.chunks = [], []
        main = None
        for p in ps:
            group = by_storage[p.data.untyped_storage().data_ptr()]
            if len(group) == 1 or p.dtype != torch.float32:
                packed.append(p)
        offs, total = [], 0
        for p in packed:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        if packed:
            main = _Chunk(torch.zeros(total, dtype=torch.float32, device=dev))
            for p, o in zip(packed, offs):
                n = p.numel()
                main.param[o:o + n].copy_(p.data.reshape(-1))
                p.data = main.param[o:o + n].view_as(p.data)
                main.members.append((p, o, n))
            self.chunks.append(main)
        seen = set()
        for p in ps:
            key = p.data.untyped_storage().data_ptr()
            group = by_storage.get(key, [])
            if len(group) > 1 and p.dtype == torch.float32 and key not in seen:
                seen.add(key)
                n_store = p.data.untyped_storage().nbytes() // 4
                buf = torch.empty(0, dtype=torch.float32, device=dev).set_(p.data.untyped_storage(), 0, (n_store,), (1,))
                ch = _Chunk(buf)
                for q in group:
                    ch.members.append((q, q.data.storage_offset(), q.numel()))
                self.chunks.append(ch)
        self.state_dev = torch.zeros(4, dtype=torch.float32, device=dev)   # {step, lr, bc1, sqrt(bc2)}
        self.n_params = sum(p.numel() for p in ps)
        self._flat_ready = True

    def prepare(self):
        if not self._flat_ready:
            self._flatten()
        return self

    # flat views of chunk 0, kept for callers that want "the" buffers (tests, DESIGN.md examples)
    @property
    def flat_param(self):
        return self.prepare().chunks[0].param

    @property
    def flat_grad(self):
        return self.prepare().chunks[0].grad

    def grad_buffers(self):
        return [c.grad for c in self.prepare().chunks]

    def enable_peer_exchange(self) -> bool:
        """Move the flat gradient buffer into peer-mapped memory and exchange it with parallel.PeerAllReduce (one kernel over
        NVLink peer memory) instead of the NCCL all-reduce.  COLLECTIVE: every rank must call it at the same point.  Returns
        False (and leaves everything as it was) when the mappings cannot be made on every rank."""
        import os
        from . import parallel
        self.prepare()
        if getattr(self, "_peer", None) is not None:
            return True
        if getattr(self, "_peer_failed", False) or os.environ.get("B200MED_PEER_EXCHANGE", "1") == "0":
            return False
        # ONE peer-mapped allocation holds the gradient twins of all chunks back to back (16-byte aligned): one kernel per step
        offs, total = [], 0
        for c in self.chunks:
            offs.append(total)
            total += (c.grad.numel() + 3) // 4 * 4
        try:
            peer = parallel.PeerAllReduce(total, self.chunks[0].grad.device)
        except Exception as e:      # noqa: BLE001 -- collective failure: every rank lands here together
            import warnings
            warnings.warn(f"b200med: gradient exchange stays on NCCL ({e})", RuntimeWarning)
            self._peer_failed = True
            return False
        for c, off in zip(self.chunks, offs):
            new = peer.buffer[off:off + c.grad.numel()]
            with torch.no_grad():
                new.copy_(c.grad)
            for p, o, n in c.members:       # gradient views that pointed into the old buffer
                if p.grad is not None and p.grad.data_ptr() == c.grad[o:o + n].data_ptr():
                    p.grad = new[o:o + n].view_as(p.data)
            c.grad = new
        self._peer = peer
        return True

    def peer_all_reduce(self) -> bool:
        """Sum the flat gradient buffer over the ranks with the peer-memory kernel; False when it is not enabled."""
        peer = getattr(self, "_peer", None)
        if peer is None:
            return False
        peer.all_reduce()
        return True

    # ------------------------------------------------------------------------------------ step protocol
    def zero_grad(self, set_to_none: bool = False):
        self.prepare()
        if self._active is None:
            # before the first step nothing is known about which parameters take part: use None like torch
            for c in self.chunks:
                for p, _, _ in c.members:
                    p.grad = None
            return
        # Detach the views: autograd then ASSIGNS each parameter's incoming gradient (no `grad += g` kernel per tensor);
        # _refresh_active() gathers them into the chunk with one multi-tensor copy before the all-reduce / Adam step.
        for c in self.chunks:
            for p, _, _ in c.members:
                p.grad = None

    def _refresh_active(self, advance=None):
        """Adopt the gradients of parameters seen for the first time and rebuild the active ranges.  advance = (beta1, beta2):
        the launch that gathers the gradients also advances the Adam step scalars (returns True when it did)."""
        from . import lstm_stack
        lstm_stack.join_pending()      # gradient GEMMs still running on the side stream (lstm_stack.DEFER_JOIN)
        changed = self._active is None
        if self._active is None:
            self._active = {}
        for c in self.chunks:
            for p, o, n in c.members:
                if not self._active.get(id(p), False) and p.grad is not None:
                    view = c.grad[o:o + n].view_as(p.data)
                    if p.grad.data_ptr() != view.data_ptr():
                        view.copy_(p.grad)
                        p.grad = view
                    self._active[id(p)] = True
                    changed = True
        # gather this step's gradients into the chunks (one multi-tensor copy; untouched active parameters read as zero)
        dst, src, zero = [], [], []
        for c in self.chunks:
            for p, o, n in c.members:
                if not self._active.get(id(p), False):
                    continue
                view = c.grad[o:o + n].view_as(p.data)
                if p.grad is None:
                    zero.append(view)
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    dst.append(view)
                    src.append(p.grad.detach())
                    p.grad = view
        advanced = False
        if dst:      # one kernel for all tensors (was torch._foreach_copy_: 18 us for 28 tensors on 44 CTAs)
            ops.multi_copy([d.reshape(-1) for d in dst], [s_.contiguous().reshape(-1) for s_ in src],
                           adam=None if advance is None else (self.state_dev, *advance))
            advanced = advance is not None
        if zero:
            torch._foreach_zero_(zero)
        if changed:
            for c in self.chunks:
                runs = []
                for p, o, n in sorted(c.members, key=lambda m: m[1]):
                    if not self._active.get(id(p), False):
                        continue
                    b, e = o // 4 * 4, (o + n + 3) // 4 * 4
                    if runs and b <= runs[-1][1]:
                        runs[-1][1] = max(runs[-1][1], e)
                    else:
                        runs.append([b, e])
                c.runs = [(b, min(e, c.param.numel())) for b, e in runs]
        return advanced

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.state_dev[1:2].fill_(lr)
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self, closure=None):
        self.prepare()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        advanced = self._refresh_active(advance=(b1, b2))
        self.sync_lr()
        if not advanced:
            ops.adam_advance(self.state_dev, b1, b2)
        for c in self.chunks:
            for b, e in c.runs:
                ops.adam_step(c.param[b:e], c.grad[b:e], c.exp_avg[b:e], c.exp_avg_sq[b:e], self.state_dev, b1, b2,
                              g["eps"], g["weight_decay"], self.grad_scale)

    # ------------------------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """torch's param_groups plus the flat Adam state: per chunk exp_avg / exp_avg_sq, the device step scalars
        {step, lr, bc1, sqrt(bc2)} and which parameters are active (by position)."""
        self.prepare()
        sd = super().state_dict()
        order = [p for g in self.param_groups for p in g["params"]]
        sd["b200med"] = {
            "chunks": [{"exp_avg": c.exp_avg.detach().cpu().clone(), "exp_avg_sq": c.exp_avg_sq.detach().cpu().clone()} for c in self.chunks],
            "state_dev": self.state_dev.detach().cpu().clone(),
            "active": None if self._active is None else [bool(self._active.get(id(p), False)) for p in order],
            "grad_scale": self.grad_scale,
        }
        return sd

    def load_state_dict(self, state_dict):
        extra = state_dict.get("b200med")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "b200med"})
        if extra is None:
            return
        self.prepare()
        if len(extra["chunks"]) != len(self.chunks):
            raise ValueError("FusedAdam.load_state_dict: the checkpoint has a different chunk layout")
        with torch.no_grad():
            for c, rec in zip(self.chunks, extra["chunks"]):
                c.exp_avg.copy_(rec["exp_avg"]); c.exp_avg_sq.copy_(rec["exp_avg_sq"])
            self.state_dev.copy_(extra["state_dev"])
        self._lr_on_device = None
        self.grad_scale = extra.get("grad_scale", 1.0)
        if extra["active"] is not None:
            order = [p for g in self.param_groups for p in g["params"]]
            self._active = {id(p): a for p, a in zip(order, extra["active"])}
            for c in self.chunks:          # rebuild the active ranges
                runs = []
                for p, o, n in sorted(c.members, key=lambda m: m[1]):
                    if not self._active.get(id(p), False):
                        continue
                    b, e = o // 4 * 4, (o + n + 3) // 4 * 4
                    if runs and b <= runs[-1][1]:
                        runs[-1][1] = max(runs[-1][1], e)
                    else:
                        runs.append([b, e])
                c.runs = [(b, min(e, c.param.numel())) for b, e in runs]
