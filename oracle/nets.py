"""Oracle: the classifier networks as plain torch-CPU modules.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference's arithmetic for these layers lives in PyTorch (``nn.Linear``, ``nn.Conv1d``,
``nn.LSTM``, ``nn.BatchNorm1d``; pinned torch 2.6, this image has 2.11 -- SURVEY.md §8c), so
the oracle calls the same ops.  What is restated here is the *composition*: layer order,
parameter names (``state_dict`` keys must interchange with the reference's checkpoints) and
the construction / re-initialisation order, which fixes the seed-42 weights
(SURVEY Appendix A-10).
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


class OracleFeatureExtractor(nn.Module):
    """Per-frame MLP in_dim -> hidden... -> out_dim with ReLU between layers.
    Reference MED/modeling/models.py:6-47 (xavier-normal weights, every bias 0.1)."""

    def __init__(self, input_dim=2048, output_dim=32, hidden_dims=(512, 256)):
        super().__init__()
        dims = [input_dim] + list(hidden_dims)
        layers = OrderedDict()
        for i in range(len(hidden_dims)):
            layers[f"linear_{i}"] = nn.Linear(dims[i], dims[i + 1])
            layers[f"relu_{i}"] = nn.ReLU()
        layers["output"] = nn.Linear(dims[-1], output_dim)
        self.linear = nn.Sequential(layers)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0.1)

    def forward(self, x):
        return self.linear(x)


def _conv_block(cin, cout):
    return [nn.Conv1d(cin, cout, kernel_size=3, stride=1), nn.MaxPool1d(2, 2),
            nn.Dropout(p=0.2), nn.BatchNorm1d(cout)]


class OracleCNN(nn.Module):
    """Window CNN head.  Reference MED/modeling/models.py:49-131: two (W=10) or three (W=30)
    conv/pool/dropout/BN blocks, then Linear 256/32/16/C with ReLU+BN.  Only W in {10, 30}
    is defined there (:66, :78).  Init quirk (:122-131): kaiming-normal(fan_out) convs,
    xavier-normal linears, and only the LAST module's bias is set to 0.1."""

    def __init__(self, in_features=58, window_size=30, n_classes=1):
        super().__init__()
        self.name = "SimpleCNN"
        self.window_size, self.in_features, self.n_classes = window_size, in_features, n_classes
        if window_size == 10:
            chans = [in_features, 64, 128]
        elif window_size == 30:
            chans = [in_features, 64, 128, 256]
        else:
            raise AttributeError("the reference CNN defines no layers for this window size")
        blocks = []
        for a, b in zip(chans[:-1], chans[1:]):
            blocks += _conv_block(a, b)
        self.convolutional_layers = nn.Sequential(*blocks, nn.Flatten())
        length = window_size
        for _ in chans[1:]:
            length = (length - 2) // 2
        n_features = chans[-1] * length
        self.linear_layers = nn.Sequential(
            nn.Linear(n_features, 256), nn.ReLU(), nn.BatchNorm1d(256),
            nn.Linear(256, 32), nn.ReLU(), nn.BatchNorm1d(32),
            nn.Linear(32, 16), nn.ReLU(), nn.BatchNorm1d(16),
            nn.Linear(16, n_classes))
        last = None
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
            last = m
        if last.bias is not None:
            nn.init.constant_(last.bias, 0.1)

    def forward(self, x):
        return self.linear_layers(self.convolutional_layers(x))


class OracleLSTM(nn.Module):
    """Window LSTM head.  Reference MED/modeling/models.py:135-220: [B,F,W] -> transpose ->
    nn.LSTM(F, H, layers, batch_first, dropout .2) -> ReLU -> last step -> Linear 256/64/C with
    ReLU+BN.  nn.LSTM keeps its default init; Linear layers get xavier-normal and zero bias."""

    def __init__(self, in_features=58, window_size=30, num_layers=3, hidden_size=128, n_classes=1):
        super().__init__()
        self.name = "SimpleLSTM"
        self.window_size, self.in_features = window_size, in_features
        self.layer_dim, self.hidden_size, self.n_classes = num_layers, hidden_size, n_classes
        self.lstm = nn.LSTM(input_size=in_features, hidden_size=hidden_size, num_layers=num_layers,
                            batch_first=True, dropout=0.2)
        self.linear_layers = nn.Sequential(
            nn.Flatten(), nn.Linear(hidden_size, 256), nn.ReLU(), nn.BatchNorm1d(256),
            nn.Linear(256, 64), nn.ReLU(), nn.BatchNorm1d(64), nn.Linear(64, n_classes))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        seq, _ = self.lstm(x.transpose(1, 2).contiguous())
        return self.linear_layers(F.relu(seq)[:, -1, :])


class _OracleDilatedLayer(nn.Module):
    """Reference MED/modeling/models_TCN.py:104-137."""

    def __init__(self, dilation, channels, causal):
        super().__init__()
        self.dilation, self.causal_conv = dilation, causal
        pad = dilation * 2 if causal else dilation
        self.conv_dilated = nn.Conv1d(channels, channels, 3, padding=pad, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(channels, channels, 1)
        self.dropout = nn.Dropout()

    def forward(self, x):
        y = F.relu(self.conv_dilated(x))
        if self.causal_conv:
            y = y[:, :, :-(self.dilation * 2)]
        return x + self.dropout(self.conv_1x1(y))


class _OracleStage(nn.Module):
    """Reference MED/modeling/models_TCN.py:76-101."""

    def __init__(self, num_layers, f_maps, dim, num_classes, causal):
        super().__init__()
        self.conv_1x1 = nn.Conv1d(dim, f_maps, 1)
        self.layers = nn.ModuleList(
            [copy.deepcopy(_OracleDilatedLayer(2 ** i, f_maps, causal)) for i in range(num_layers)])
        self.conv_out_classes = nn.Conv1d(f_maps, num_classes, 1)

    def forward(self, x):
        h = self.conv_1x1(x)
        for layer in self.layers:
            h = layer(h)
        return self.conv_out_classes(h)


class OracleTeCNo(nn.Module):
    """Multi-stage causal TCN ("TeCNo").  Reference MED/modeling/models_TCN.py:17-53: stage 1 on
    the features, each later stage on softmax(previous logits); output [stages, 1, C, T]."""

    def __init__(self, mstcn_stages, mstcn_layers, mstcn_f_maps, mstcn_f_dim, out_features,
                 mstcn_causal_conv):
        super().__init__()
        self.name = "TeCNo"
        self.stage1 = _OracleStage(mstcn_layers, mstcn_f_maps, mstcn_f_dim, out_features, mstcn_causal_conv)
        self.stages = nn.ModuleList(
            [copy.deepcopy(_OracleStage(mstcn_layers, mstcn_f_maps, out_features, out_features, mstcn_causal_conv))
             for _ in range(mstcn_stages - 1)])

    def forward(self, x):
        logits = self.stage1(x)
        outs = [logits]
        for st in self.stages:
            logits = st(F.softmax(logits, dim=1))
            outs.append(logits)
        return torch.stack(outs, dim=0)


def build_head(exp_kwargs: dict, in_features: int, window_size: int):
    """Reference MED/modeling/modeling_utils.py:3043-3117 (supported heads only)."""
    name = exp_kwargs["model_name"]
    n_out = exp_kwargs.get("out_features", 1)
    if name == "SimpleCNN":
        return OracleCNN(in_features, window_size, n_out)
    if name == "SimpleLSTM":
        return OracleLSTM(in_features, window_size, hidden_size=exp_kwargs["hidden_size"],
                          num_layers=exp_kwargs["num_layers"], n_classes=n_out)
    if name == "TeCNo":
        return OracleTeCNo(exp_kwargs["mstcn_stages"], exp_kwargs["mstcn_layers"], exp_kwargs["mstcn_f_maps"],
                           exp_kwargs["mstcn_f_dim"], exp_kwargs["out_features"], exp_kwargs["mstcn_causal_conv"])
    raise ValueError(f"Model {name} is not supported.")


def build_objects(exp_kwargs: dict, in_features_dict: dict, class_counts, window_size: int = 0):
    """Reference MED/modeling/modeling_utils.py:194-262: seed 42, head first, then the feature
    extractor (that order fixes the weights), Adam over FE+head params with coupled weight
    decay, criterion by (pos_weight, error_type, dataset_type), cosine LR per epoch."""
    torch.manual_seed(42)
    head = build_head(exp_kwargs, in_features_dict[exp_kwargs["data_type"]], window_size)
    if exp_kwargs["data_type"] != "kinematics":
        fe = OracleFeatureExtractor(2048, exp_kwargs["video_dims"], [512, 256])
        params = list(fe.parameters()) + list(head.parameters())
    else:
        fe, params = None, list(head.parameters())
    opt = torch.optim.Adam(params, lr=exp_kwargs["lr"], weight_decay=exp_kwargs["weight_decay"])
    if exp_kwargs["pos_weight"]:
        if exp_kwargs["error_type"] == "global":
            crit = nn.BCEWithLogitsLoss(pos_weight=torch.tensor(class_counts[0] / class_counts[1], dtype=torch.float32))
        else:
            crit = nn.CrossEntropyLoss(weight=torch.tensor(class_counts, dtype=torch.float32))
    elif exp_kwargs["dataset_type"] == "window":
        crit = nn.BCEWithLogitsLoss() if exp_kwargs["error_type"] == "global" else nn.CrossEntropyLoss()
    else:
        crit = nn.CrossEntropyLoss(reduction="none") if exp_kwargs["error_type"] == "sequential" else nn.CrossEntropyLoss()
    sched = (torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=exp_kwargs["n_epochs"], eta_min=1e-6)
             if exp_kwargs["lr_scheduler"] else None)
    return fe, head, crit, opt, sched


def disable_dropout(*modules):
    """Parity runs compare train-mode gradients with dropout switched off on both sides
    (CPU RNG streams cannot be replayed on the GPU -- SURVEY.md §7 'hard parts')."""
    for mod in modules:
        if mod is None:
            continue
        for m in mod.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
            if isinstance(m, nn.LSTM):
                m.dropout = 0.0
