"""TeCNo stage on the fused sm_100a kernels of ``csrc/tcn.cu`` (C ABI: ``b200med_tcn_*``).

One autograd node per SingleStageModel (reference MED/modeling/models_TCN.py:76-100): 1x1 input convolution
(fp32 GEMM on the [T, F] frame rows) -> L fused DilatedResidualLayer launches -> 1x1 class convolution, with the
inter-stage ``softmax(dim=1)`` (models_TCN.py:48) folded into the consuming stage.  Activations are time-major
[T, 64]; the stage returns logits [C, T].  Backward = 2 launches per layer + one reduction of the weight-gradient
partials per stage (deterministic).  There is no torch / cuDNN convolution on this path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops

MAPS = ops.TCN_MAPS
_WD, _W1 = 3 * MAPS * MAPS, MAPS * MAPS


class StageConfig:
    """Static description of one stage + the cached device table of its layer-parameter addresses."""

    def __init__(self, n_layers: int, causal: bool, softmax_in: bool, layer_base: int = 0):
        self.n_layers, self.causal, self.softmax_in, self.layer_base = n_layers, bool(causal), bool(softmax_in), layer_base
        self.drop_p: List[float] = [0.0] * n_layers     # per layer, 0 in eval mode
        self.seed = 0
        self.tloc: Optional[torch.Tensor] = None         # ragged batch geometry (None = one video)
        self.trem: Optional[torch.Tensor] = None
        self._ptr_key, self._ptr_table = None, None

    def ptr_table(self, layer_params: Sequence[torch.Tensor]) -> torch.Tensor:
        key = tuple(p.data_ptr() for p in layer_params)
        if key != self._ptr_key:
            self._ptr_table = torch.tensor(key, dtype=torch.int64).to(layer_params[0].device)
            self._ptr_key = key
        return self._ptr_table


def supported(n_maps: int, kernel_size: int, n_classes: int, in_dim: int) -> bool:
    return n_maps == MAPS and kernel_size == 3 and 1 <= n_classes <= 8 and in_dim >= 1


def require_cuda(x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError("b200med: MultiStageModel runs on CUDA devices only (no CPU fallback)")


def _check(x: torch.Tensor, params):
    require_cuda(x)
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise TypeError("b200med: TeCNo parameters must be contiguous fp32 tensors")


class TcnStageFunction(torch.autograd.Function):
    """forward(x, cfg, in_w, in_b, (wd, bd, w1, b1) * L, out_w, out_b) -> logits [C, T].

    x: [T, F] frame rows (stage 1) or the previous stage's logits [C, T] when ``cfg.softmax_in``."""

    @staticmethod
    def forward(ctx, x, cfg: StageConfig, *params):
        L = cfg.n_layers
        _check(x, params)
        in_w, in_b, out_w, out_b = params[0], params[1], params[-2], params[-1]
        layer_params = params[2:2 + 4 * L]
        x = x.detach().contiguous().float()
        p_in = ops.tcn_softmax_fwd(x) if cfg.softmax_in else None
        xin = p_in if cfg.softmax_in else x
        T = xin.shape[0]
        dev = x.device
        keep = any(ctx.needs_input_grad)
        acts = torch.empty((L + 1) if keep else 2, T, MAPS, dtype=torch.float32, device=dev)
        ys = torch.empty(L, T, MAPS, dtype=torch.float32, device=dev) if keep else None
        ops.linear_fwd_f32(xin, in_w.detach().view(MAPS, -1), in_b.detach(), relu=False, out=acts[0])
        pack = ops.tcn_pack(cfg.ptr_table(layer_params), L)
        for l in range(L):
            src, dst = (acts[l], acts[l + 1]) if keep else (acts[l & 1], acts[(l + 1) & 1])
            ops.tcn_layer_fwd(src, pack[l], dst, None if ys is None else ys[l], 2 ** l, cfg.causal, cfg.drop_p[l], cfg.seed,
                              (cfg.layer_base + l) << 40, cfg.tloc, cfg.trem)
        last = acts[L] if keep else acts[L & 1]
        n_cls = out_w.shape[0]
        logits = ops.tcn_out_fwd(last, out_w.detach().view(n_cls, MAPS), out_b.detach())
        if keep:
            ctx.cfg, ctx.geom = cfg, (list(cfg.drop_p), cfg.seed, cfg.tloc, cfg.trem)
            ctx.save_for_backward(xin, acts, ys, pack, in_w, out_w)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        cfg = ctx.cfg
        L = cfg.n_layers
        drop_p, seed, tloc, trem = ctx.geom
        xin, acts, ys, pack, in_w, out_w = ctx.saved_tensors
        T, dev = xin.shape[0], xin.device
        n_cls = out_w.shape[0]
        dA, dl_t = ops.tcn_out_bwd(dlogits.contiguous().float(), out_w.detach().view(n_cls, MAPS))
        d_out_w, d_out_b = ops.linear_bwd_weight_f32(dl_t, acts[L])
        n_slots = ops.tcn_slots(T)
        partials = torch.empty(L, n_slots, ops.TCN_GRAD_FLOATS, dtype=torch.float32, device=dev)
        dpre = torch.empty(T, MAPS, dtype=torch.float32, device=dev)
        spare = torch.empty(T, MAPS, dtype=torch.float32, device=dev)
        for l in reversed(range(L)):
            ops.tcn_layer_bwd_hidden(dA, acts[l], ys[l], pack[l], dpre, partials[l], n_slots, 2 ** l, cfg.causal, drop_p[l],
                                     seed, (cfg.layer_base + l) << 40, tloc, trem)
            ops.tcn_layer_bwd_input(dpre, dA, pack[l], spare, 2 ** l, cfg.causal, tloc, trem)
            dA, spare = spare, dA
        grads = ops.tcn_reduce_grads(partials, L, n_slots)
        d_in_w, d_in_b = ops.linear_bwd_weight_f32(dA, xin)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_bwd_data_f32(dA, in_w.detach().view(MAPS, -1))
            if cfg.softmax_in:
                dx = ops.tcn_softmax_bwd(xin, dx)
        out = [dx, None, d_in_w.view_as(in_w), d_in_b]
        for l in range(L):
            g = grads[l]
            out += [g[:_WD].view(MAPS, MAPS, 3), g[_WD + _W1:_WD + _W1 + MAPS], g[_WD:_WD + _W1].view(MAPS, MAPS, 1),
                    g[_WD + _W1 + MAPS:]]
        out += [d_out_w.view_as(out_w), d_out_b]
        return tuple(out)


def ragged_geometry(lengths: Sequence[int], device) -> tuple:
    """(tloc, trem) int32 [sum(lengths)] for videos concatenated along T: index inside the video, frames left after it."""
    lens = torch.as_tensor(list(lengths), dtype=torch.int64)
    starts = torch.cumsum(lens, 0) - lens
    tloc = torch.arange(int(lens.sum()), dtype=torch.int64) - torch.repeat_interleave(starts, lens)
    trem = torch.repeat_interleave(lens, lens) - 1 - tloc
    return tloc.to(torch.int32).to(device), trem.to(torch.int32).to(device)
