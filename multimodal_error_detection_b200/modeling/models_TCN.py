"""Drop-in for the TeCNo part of the reference ``MED/modeling/models_TCN.py`` (MultiStageModel,
models_TCN.py:17-137).  TransSVNet / COG are out of scope (they
hard-code the Apple ``mps`` device and need CLIP weights).

Same constructor signature, ``state_dict`` keys (``stage1.conv_1x1``, ``stage1.layers.{i}.conv_dilated``,
``stage1.layers.{i}.conv_1x1``, ``stage1.conv_out_classes``, ``stages.{s}...``), construction order
(=> same seed-42 weights) and output layout ``[stages, 1, C, T]``.  The dilated causal stack runs on
stock torch conv layers on the GPU in this round (its fused time-tiled kernel is SURVEY section 8f row 2).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F


class DilatedResidualLayer(nn.Module):
    def __init__(self, dilation, in_channels, out_channels, causal_conv=False, kernel_size=3):
        super().__init__()
        self.causal_conv, self.dilation, self.kernel_size = causal_conv, dilation, kernel_size
        pad = dilation * (kernel_size - 1) if causal_conv else dilation
        self.conv_dilated = nn.Conv1d(in_channels, out_channels, kernel_size, padding=pad, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(out_channels, out_channels, 1)
        self.dropout = nn.Dropout()

    def forward(self, x):
        y = F.relu(self.conv_dilated(x))
        if self.causal_conv:
            y = y[:, :, :-(self.dilation * 2)]   # drop the right overhang -> causal
        return x + self.dropout(self.conv_1x1(y))


class SingleStageModel(nn.Module):
    def __init__(self, num_layers, num_f_maps, dim, num_classes, causal_conv=False):
        super().__init__()
        self.conv_1x1 = nn.Conv1d(dim, num_f_maps, 1)
        self.layers = nn.ModuleList([copy.deepcopy(DilatedResidualLayer(2 ** i, num_f_maps, num_f_maps, causal_conv=causal_conv))
                                     for i in range(num_layers)])
        self.conv_out_classes = nn.Conv1d(num_f_maps, num_classes, 1)

    def forward(self, x):
        out = self.conv_1x1(x)
        for layer in self.layers:
            out = layer(out)
        return self.conv_out_classes(out)


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

    def forward(self, x):
        out = self.stage1(x)
        outs = [out]
        for s in self.stages:
            out = s(F.softmax(out, dim=1))
            outs.append(out)
        return torch.stack(outs, dim=0)
