"""TeCNo stage on the fused sm_100a kernels of ``csrc/tcn.cu`` (C ABI: ``b200med_tcn_*``).

One autograd node per SingleStageModel (reference MED/modeling/models_TCN.py:76-100): 1x1 input convolution
(fp32 GEMM on the [T, F] frame rows) -> L fused DilatedResidualLayer launches -> 1x1 class convolution, with the
inter-stage ``softmax(dim=1)`` (models_TCN.py:48) folded into the consuming stage.  Activations are time-major
[T, 64]; the stage returns logits [C, T].  Backward = 2 launches per layer + one reduction of the weight-gradient
partials per stage (deterministic).  There is no torch / cuDNN convolution on this path.  A stage forward / backward is
ONE C call each (``b200med_tcn_stage_fwd`` / ``_bwd`` sequence the kernels in C: driven launch by launch from Python
the step was host-bound, 2.2 ms for ~0.3 ms of kernels).
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
        self.grad_enabled = True                          # torch.is_grad_enabled() at call time (set by the caller)
        self.precision = "fp32"                          # "bf16": inference runs the layers on the tcgen05 kernel
        self.seed_dev: Optional[torch.Tensor] = None     # int64 device scalar added to the seed on the device (CUDA graphs)
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
        # needs_input_grad ignores torch.no_grad() (it mirrors requires_grad), and inside forward() grad mode is always off:
        # the caller records the grad mode in the config
        keep = cfg.grad_enabled and any(ctx.needs_input_grad)
        if cfg.precision == "bf16" and not keep and max(cfg.drop_p) == 0.0 and ops.has_tcgen05():
            # inference in the bf16 mode: layers as tcgen05 MMAs over 128-frame tiles, fp32 residual stream
            return ops.tcn_stage_fwd_bf16(x.detach().contiguous().float(), cfg.softmax_in, in_w.detach().view(MAPS, -1), in_b.detach(),
                                          cfg.ptr_table(params[2:2 + 4 * L]), L, out_w.detach().view(out_w.shape[0], MAPS),
                                          out_b.detach(), cfg.causal, cfg.tloc, cfg.trem)
        r = ops.tcn_stage_fwd(x.detach().contiguous().float(), cfg.softmax_in, in_w.detach().view(MAPS, -1), in_b.detach(),
                              cfg.ptr_table(params[2:2 + 4 * L]), L, out_w.detach().view(out_w.shape[0], MAPS), out_b.detach(),
                              cfg.causal, cfg.drop_p, cfg.seed, cfg.layer_base, keep, cfg.tloc, cfg.trem, cfg.seed_dev)
        if keep:
            ctx.cfg, ctx.geom = cfg, (list(cfg.drop_p), cfg.seed, cfg.layer_base, cfg.tloc, cfg.trem, cfg.seed_dev)
            ctx.save_for_backward(r["xin"], r["acts"], r["ys"], r["pack"], in_w, out_w)
        return r["logits"]

    @staticmethod
    def backward(ctx, dlogits):
        cfg = ctx.cfg
        L = cfg.n_layers
        drop_p, seed, layer_base, tloc, trem, seed_dev = ctx.geom
        xin, acts, ys, pack, in_w, out_w = ctx.saved_tensors
        r = ops.tcn_stage_bwd(dlogits.contiguous().float(), xin, cfg.softmax_in, in_w.detach().view(MAPS, -1),
                              out_w.detach().view(out_w.shape[0], MAPS), L, cfg.causal, acts, ys, pack,
                              ctx.needs_input_grad[0], drop_p, seed, layer_base, tloc, trem, seed_dev)
        out = [r["dx"], None, r["d_in_w"].view_as(in_w), r["d_in_b"]]
        for l in range(L):
            g = r["layer_grads"][l]
            out += [g[:_WD].view(MAPS, MAPS, 3), g[_WD + _W1:_WD + _W1 + MAPS], g[_WD:_WD + _W1].view(MAPS, MAPS, 1),
                    g[_WD + _W1 + MAPS:]]
        out += [r["d_out_w"].view_as(out_w), r["d_out_b"]]
        return tuple(out)


def ragged_geometry(lengths: Sequence[int], device) -> tuple:
    """(tloc, trem) int32 [sum(lengths)] for videos concatenated along T: index inside the video, frames left after it."""
    lens = torch.as_tensor(list(lengths), dtype=torch.int64)
    starts = torch.cumsum(lens, 0) - lens
    tloc = torch.arange(int(lens.sum()), dtype=torch.int64) - torch.repeat_interleave(starts, lens)
    trem = torch.repeat_interleave(lens, lens) - 1 - tloc
    return tloc.to(torch.int32).to(device), trem.to(torch.int32).to(device)
