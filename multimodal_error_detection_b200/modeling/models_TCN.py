"""Drop-in for the TeCNo part of the reference ``MED/modeling/models_TCN.py`` (MultiStageModel,
models_TCN.py:17-137).  TransSVNet / COG are out of scope (they
hard-code the Apple ``mps`` device and need CLIP weights).

Same constructor signature, ``state_dict`` keys (``stage1.conv_1x1``, ``stage1.layers.{i}.conv_dilated``,
``stage1.layers.{i}.conv_1x1``, ``stage1.conv_out_classes``, ``stages.{s}...``), construction order
(=> same seed-42 weights) and output layout ``[stages, 1, C, T]``.

The modules below only HOLD the parameters (so ``state_dict`` / ``load_state_dict`` / the optimiser see the reference's
tensors); the arithmetic of a stage runs on the fused sm_100a kernels of ``csrc/tcn.cu`` through
:class:`multimodal_error_detection_b200.tcn.TcnStageFunction` -- one launch per DilatedResidualLayer, time-major
activations, the inter-stage softmax folded into the consuming stage.  The kernels are specialised for the reference's
configuration (64 feature maps, kernel size 3, one video per forward -- or a ragged batch through ``forward_ragged`` --,
<= 8 classes); any other shape raises ``ValueError``: there is no second implementation to fall back to.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .. import tcn


class DilatedResidualLayer(nn.Module):
    def __init__(self, dilation, in_channels, out_channels, causal_conv=False, kernel_size=3):
        super().__init__()
        self.causal_conv, self.dilation, self.kernel_size = causal_conv, dilation, kernel_size
        pad = dilation * (kernel_size - 1) if causal_conv else dilation
        self.conv_dilated = nn.Conv1d(in_channels, out_channels, kernel_size, padding=pad, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(out_channels, out_channels, 1)
        self.dropout = nn.Dropout()

    def forward(self, x):
        raise RuntimeError("b200med: DilatedResidualLayer only holds parameters; a stage runs as a whole on the fused kernels "
                           "(SingleStageModel / MultiStageModel.forward)")


class SingleStageModel(nn.Module):
    def __init__(self, num_layers, num_f_maps, dim, num_classes, causal_conv=False):
        super().__init__()
        self.num_layers, self.num_f_maps, self.dim, self.num_classes, self.causal_conv = \
            num_layers, num_f_maps, dim, num_classes, causal_conv
        self.conv_1x1 = nn.Conv1d(dim, num_f_maps, 1)
        self.layers = nn.ModuleList([copy.deepcopy(DilatedResidualLayer(2 ** i, num_f_maps, num_f_maps, causal_conv=causal_conv))
                                     for i in range(num_layers)])
        self.conv_out_classes = nn.Conv1d(num_f_maps, num_classes, 1)
        self._cfgs = {}

    # ------------------------------------------------------------------------------ fused path
    def fused_supported(self) -> bool:
        return (len(self.layers) >= 1 and tcn.supported(self.num_f_maps, self.layers[0].kernel_size, self.num_classes, self.dim)
                and all(l.dilation == 2 ** i and l.kernel_size == 3 for i, l in enumerate(self.layers)))

    def fused_ok(self, x: torch.Tensor) -> bool:
        return x.dim() == 3 and x.shape[0] == 1 and x.dtype == torch.float32 and self.fused_supported()

    def _stage_params(self):
        # cached: ~40 nn.Module attribute lookups per call otherwise (the eager frame step is host-bound).  The Parameter
        # objects are stable (load_state_dict / .to() / the optimiser's re-homing change .data in place).
        cache = self.__dict__.get("_param_cache")
        if cache is None:
            ps = [self.conv_1x1.weight, self.conv_1x1.bias]
            for l in self.layers:
                ps += [l.conv_dilated.weight, l.conv_dilated.bias, l.conv_1x1.weight, l.conv_1x1.bias]
            ps += [self.conv_out_classes.weight, self.conv_out_classes.bias]
            cache = self.__dict__["_param_cache"] = (ps, [l.dropout for l in self.layers])
        return cache[0]

    def run_fused(self, x: torch.Tensor, softmax_in: bool = False, seed: int = 0, layer_base: int = 0, geom=(None, None),
                  seed_dev=None, precision: str = "fp32"):
        """x [1, F, T] (or the previous stage's logits [1, C, T] with ``softmax_in``) -> logits [1, C, T]."""
        tcn.require_cuda(x)
        cfg = self._cfgs.get(softmax_in)
        if cfg is None:
            cfg = self._cfgs[softmax_in] = tcn.StageConfig(len(self.layers), self.causal_conv, softmax_in)
        cfg.layer_base, cfg.seed, cfg.seed_dev = layer_base, seed, seed_dev
        cfg.precision, cfg.grad_enabled = precision, torch.is_grad_enabled()
        params = self._stage_params()
        cfg.drop_p = [float(d.p) if (self.training and d.training) else 0.0 for d in self.__dict__["_param_cache"][1]]
        cfg.tloc, cfg.trem = geom
        xin = x[0] if softmax_in else x[0].t()      # [C, T] logits, or [T, F] rows (a free view of the [1, T, F] batch)
        return tcn.TcnStageFunction.apply(xin, cfg, *params).unsqueeze(0)

    def check_supported(self, x: torch.Tensor) -> None:
        tcn.require_cuda(x)
        if not self.fused_ok(x):
            raise ValueError("b200med TeCNo kernels cover the reference's configuration -- 64 feature maps, kernel size 3, dilation "
                             f"2^i, <= 8 classes, fp32 input [1, F, T] -- got maps={self.num_f_maps}, classes={self.num_classes}, "
                             f"input {tuple(x.shape)} {x.dtype}")

    def forward(self, x):
        self.check_supported(x)
        return self.run_fused(x)


class MultiStageModel(nn.Module):
    def __init__(self, mstcn_stages, mstcn_layers, mstcn_f_maps, mstcn_f_dim, out_features, mstcn_causal_conv):
        super().__init__()
        self.name = "TeCNo"
        self.num_stages, self.num_layers, self.num_f_maps = mstcn_stages, mstcn_layers, mstcn_f_maps
        self.dim, self.num_classes, self.causal_conv = mstcn_f_dim, out_features, mstcn_causal_conv
        self.stage1 = SingleStageModel(mstcn_layers, mstcn_f_maps, mstcn_f_dim, out_features, causal_conv=mstcn_causal_conv)
        self.stages = nn.ModuleList([copy.deepcopy(SingleStageModel(mstcn_layers, mstcn_f_maps, out_features, out_features,
                                                                    causal_conv=mstcn_causal_conv))
                                     for _ in range(mstcn_stages - 1)])
        self.smoothing = False
        self.impl = "b200"       # every forward runs on csrc/tcn.cu (kept as an attribute for callers that report it)
        # CUDA-graph training (engine.FrameTrainStep): a host seed fixed at capture + an int64 DEVICE counter advanced inside
        # the captured step, so that every replay draws a fresh dropout mask
        self.graph_seed = None   # (host seed, device counter tensor) or None
        # "bf16" (set by define_model_objects from exp_kwargs['precision']): INFERENCE runs the layers on the tcgen05 kernel
        # (bf16 operands, fp32 accumulation and residual stream, 2e-2 bar); training always uses the fp32 kernels
        self.precision = "fp32"

    def _seed(self) -> int:
        if not self.training:
            return 0
        return int(torch.randint(0, 2 ** 62, (1,)).item())     # host generator: reproducible under torch.manual_seed

    def forward(self, x, geom=(None, None)):
        self.stage1.check_supported(x)
        for s in self.stages:
            if not s.fused_supported():
                raise ValueError("b200med TeCNo kernels cover 64 feature maps, kernel size 3, dilation 2^i and <= 8 classes")
        seed_dev = None
        if self.training and self.graph_seed is not None:
            seed, seed_dev = self.graph_seed
            seed_dev.add_(1)
        else:
            seed = self._seed()
        out = self.stage1.run_fused(x, False, seed, 0, geom, seed_dev, self.precision)
        outs = [out]
        for i, s in enumerate(self.stages):
            out = s.run_fused(out, True, seed, (i + 1) * self.num_layers, geom, seed_dev, self.precision)
            outs.append(out)
        return torch.stack(outs, dim=0)

    @torch.no_grad()
    def forward_ragged(self, frames: torch.Tensor, lengths) -> torch.Tensor:
        """Batched inference over several videos concatenated along time: frames [sum(T_v), F] (rows of the frame
        table), lengths = T_v per video -> logits [stages, C, sum(T_v)]; taps never cross a video boundary, so the
        result equals running every video on its own (the reference's DataLoader(batch_size=1) loop)."""
        geom = tcn.ragged_geometry(lengths, frames.device)
        if int(sum(int(n) for n in lengths)) != frames.shape[0]:
            raise ValueError("lengths must sum to the number of frame rows")
        x = frames.contiguous().float().unsqueeze(0).permute(0, 2, 1)
        return self.forward(x, geom)[:, 0]
