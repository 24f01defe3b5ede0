"""Oracle: batching, loss and the train / validation loops (torch CPU + sklearn).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  This module is also the timed CPU
baseline of ``bench.py`` (``cpu_baseline`` / ``--impl reference``, kind "port"): it keeps the
reference's cost structure -- per-sample ``__getitem__`` normalisation, default collate,
main-thread DataLoader, eager fp32 modules, sklearn metrics on the host every batch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from sklearn.metrics import accuracy_score, confusion_matrix, f1_score, jaccard_score
from torch.utils.data import DataLoader, Dataset


class OracleWindowDataset(Dataset):
    """Materialised windows, standardised per sample as ``(x - mean) / std``.
    Reference MED/dataset/CustomWindowDataset.py:22-74 (subtract, then divide: :58, :60)."""

    def __init__(self, image, kin, g, e7, subjects, stats=None):
        self.image, self.kin, self.g, self.e7, self.subjects = image, kin, g, e7, list(subjects)
        self.stats = stats or {}
        n = len(e7)
        p = e7[:, -1].sum() / n
        self.binary_error_distribution = (1 - p, p)
        self.specific_error_distribution = (n / (e7[:, :-1].sum(axis=0) + 1e-5)).tolist()

    def __len__(self):
        return len(self.image)

    def __getitem__(self, i):
        x, k = self.image[i], self.kin[i]
        if "image" in self.stats:
            x = (x - self.stats["image"]["mean"]) / self.stats["image"]["std"]
        if "kinematics" in self.stats:
            k = (k - self.stats["kinematics"]["mean"]) / self.stats["kinematics"]["std"]
        return x, k, self.g[i], self.e7[i], self.subjects[i]


def make_loaders(train_ds, test_ds, batch_size):
    """Reference MED/dataset/dataset_utils.py:526-527: shuffled train loader driven by a
    ``Generator().manual_seed(42)``, ordered test loader, no workers, no drop_last."""
    tr = DataLoader(train_ds, batch_size=batch_size, shuffle=True, generator=torch.Generator().manual_seed(42))
    te = DataLoader(test_ds, batch_size=batch_size, shuffle=False, generator=torch.Generator().manual_seed(42))
    return tr, te


def select_labels(e7: torch.Tensor, exp_kwargs: dict) -> torch.Tensor:
    """Reference MED/modeling/modeling_utils.py:137-191: 'global' -> last column,
    'all_errors' -> columns 0..5; the frame path indexes one more leading dim."""
    table = {"No Error": 0, "Out_Of_View": 1, "Multiple_Attempts": 2, "Needle_Position": 3,
             "Out_Of_View_Multiple_Attempts": 4, "Multiple_Attempts_Needle_Position": 5,
             "global": -1, "all_errors": [0, 1, 2, 3, 4, 5]}
    if "error_type" not in exp_kwargs:
        raise ValueError("error_type must be defined in exp_kwargs.")
    if exp_kwargs["error_type"] not in table:
        raise ValueError(f"Error type {exp_kwargs['error_type']} is not supported.")
    col = table[exp_kwargs["error_type"]]
    return e7[:, col] if exp_kwargs["dataset_type"] == "window" else e7[:, :, col]


def fuse_inputs(images, kin, fe, exp_kwargs):
    """Reference MED/modeling/modeling_utils.py:19-84: FE on the image stream, concat with the
    kinematics on the feature axis, permute to [B, F, W]."""
    dt = exp_kwargs["data_type"]
    if dt == "multimodal":
        x = torch.cat((fe(images), kin), dim=2).permute(0, 2, 1)
    elif dt == "kinematics":
        x = kin.permute(0, 2, 1)
    elif dt == "video":
        x = images.permute(0, 2, 1) if exp_kwargs["video_dims"] == 2048 else fe(images).permute(0, 2, 1)
    else:
        raise ValueError(f"Data type {dt} is not supported.")
    if x.size(0) == 0:
        raise ValueError("Inputs tensor is empty. Check the data loader and the inputs.")
    return x


def loss_fn(outputs, labels, criterion, dataset_type):
    """Reference MED/modeling/modeling_utils.py:265-297.  Window: squeeze dim 1, apply the
    criterion.  Frame: soft two-column targets [1-e, e], CE per stage, mean over stages."""
    if dataset_type == "window":
        outputs = outputs.squeeze(1) if outputs.dim() > 1 and outputs.size(1) == 1 else outputs
        return criterion(outputs, labels), outputs
    target = torch.cat((1 - labels, labels), dim=0).transpose(1, 0)
    total = 0.0
    for j in range(outputs.shape[0]):
        total = total + criterion(outputs[j].squeeze().transpose(1, 0), target)
    return total / (outputs.shape[0] * 1.0), outputs


def binary_metrics(y, p):
    return (f1_score(y, p, average="binary", pos_label=1), f1_score(y, p, average="weighted"),
            accuracy_score(y, p), jaccard_score(y, p, average="binary", pos_label=1))


def train_epoch(model, fe, loader, criterion, optimizer, scheduler, exp_kwargs, host_metrics=True):
    """Reference MED/modeling/modeling_utils.py:300-407 ('global' error type).  Metrics are the
    MEAN over batches of per-batch sklearn scores (:377-381, :398-402); the CM is summed."""
    model.train()
    if fe is not None:
        fe.train()
    tot = np.zeros(5)
    cm = np.zeros((2, 2), dtype=int)
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = select_labels(e7, exp_kwargs).float()
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        loss, out = loss_fn(out, y, criterion, exp_kwargs["dataset_type"])
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        tot[0] += loss.item()
        if not host_metrics:
            continue
        if exp_kwargs["dataset_type"] == "frame":
            pred = torch.max(out[-1].squeeze().transpose(1, 0).data, 1)[1]
            y = y.squeeze(0)
        else:
            pred = (torch.sigmoid(out) > 0.5).float()
        yn, pn = y.detach().numpy(), pred.detach().numpy()
        tot[1:] += binary_metrics(yn, pn)
        cm += confusion_matrix(yn, pn, labels=[0, 1])
    if scheduler is not None:
        scheduler.step()
    tot /= len(loader)
    return (*tot.tolist(), cm)


@torch.no_grad()
def validate_epoch(model, fe, loader, criterion, exp_kwargs):
    """Reference MED/modeling/modeling_utils.py:688-790: metrics POOLED over all samples."""
    model.eval()
    if fe is not None:
        fe.eval()
    loss_sum, ys, ps, probs = 0.0, [], [], []
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = select_labels(e7, exp_kwargs).float()
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        if out.dim() > 1 and out.size(1) == 1:
            out = out.squeeze(1)
        loss, out = loss_fn(out, y, criterion, exp_kwargs["dataset_type"])
        loss_sum += loss.item()
        if exp_kwargs["dataset_type"] == "frame":
            pred = torch.max(out[-1].squeeze().transpose(1, 0).data, 1)[1].float()
            y = y.squeeze(0)
        else:
            sg = torch.sigmoid(out)
            pred = (sg > 0.5).float()
            probs += sg.tolist()
        ys += y.tolist()
        ps += pred.tolist()
    f1, f1w, acc, jac = binary_metrics(ys, ps)
    return loss_sum / len(loader), f1, f1w, acc, jac, confusion_matrix(ys, ps, labels=[0, 1]), ps, probs, ys


def train_epoch_es(model, fe, loader, criterion, optimizer, scheduler, exp_kwargs):
    """Error-specific (6-class) epoch.  Reference MED/modeling/modeling_utils.py:410-539 with the
    one restatement SURVEY.md §8c prescribes: the class index is passed to CrossEntropyLoss as
    ``long`` (the committed code passes float and raises on CPU)."""
    model.train()
    if fe is not None:
        fe.train()
    loss_sum, ys, ps = 0.0, [], []
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = torch.argmax(select_labels(e7, exp_kwargs).float(), dim=1).view(-1)
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        loss = criterion(out, y.long())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        loss_sum += loss.item()
        ps += torch.argmax(torch.softmax(out, dim=1), dim=1).tolist()
        ys += y.tolist()
    if scheduler is not None:
        scheduler.step()
    return (loss_sum / len(loader), *multiclass_summary(ys, ps))


def multiclass_summary(ys, ps):
    yb, pb = [int(v != 0) for v in ys], [int(v != 0) for v in ps]
    return (f1_score(yb, pb, average="binary", pos_label=1), f1_score(ys, ps, average="macro"),
            accuracy_score(yb, pb), accuracy_score(ys, ps),
            jaccard_score(yb, pb, average="binary", pos_label=1), jaccard_score(ys, ps, average="macro"),
            confusion_matrix(yb, pb), confusion_matrix(ys, ps))


def train_epoch_sequential(model, fe, loader, optimizer, scheduler, exp_kwargs):
    """Cascade stage-2 epoch.  Reference MED/modeling/modeling_utils.py:543-684: labels 0..5,
    mask = (label != 0), targets label-1, per-sample CE times mask, sum / mask.sum().
    Restatement: the -1 targets of masked rows are clamped to 0 before the CE
    (they are multiplied by 0 afterwards; unclamped they raise on CPU)."""
    model.train()
    if fe is not None:
        fe.train()
    loss_sum, ys, ps = 0.0, [], []
    ce = nn.CrossEntropyLoss(reduction="none")
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = torch.argmax(select_labels(e7, exp_kwargs).float(), dim=1).view(-1)
        mask = (y != 0).float()
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        loss = ce(out, (y - 1).clamp(min=0)) * mask
        loss = loss.sum() / mask.sum() if mask.sum() > 0 else loss
        if loss.dim() > 0:
            loss = loss.sum()
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        loss_sum += loss.item()
        pred = torch.argmax(torch.softmax(out, dim=1), dim=1) + 1
        pred = torch.where(y == 0, torch.zeros_like(pred), pred)
        ps += pred.tolist()
        ys += y.tolist()
    if scheduler is not None:
        scheduler.step()
    return loss_sum / len(loader), ys, ps


@torch.no_grad()
def validate_epoch_es(model, fe, loader, criterion, exp_kwargs):
    """Error-specific validation.  Reference MED/modeling/modeling_utils.py:793-904 (class index cast to long)
    -> (loss, *multiclass_summary, probs of class 1, preds, labels)."""
    model.eval()
    if fe is not None:
        fe.eval()
    loss_sum, ys, ps, probs = 0.0, [], [], []
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = torch.argmax(select_labels(e7, exp_kwargs).float(), dim=1).view(-1)
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        loss_sum += criterion(out, y.long()).item()
        sm = torch.softmax(out, dim=1)
        ps += torch.argmax(sm, dim=1).tolist()
        probs += sm[:, 1].tolist()
        ys += y.tolist()
    return (loss_sum / len(loader), *multiclass_summary(ys, ps), probs, ps, ys)


@torch.no_grad()
def validate_epoch_sequential(model, fe, binary_model, binary_fe, loader, exp_kwargs):
    """Cascade validation.  Reference MED/modeling/modeling_utils.py:907-1053: the binary model's RAW logit thresholded at
    0.5 (:979-980) gates the 5-class model; the loss keeps the committed [B] * [B, 1] -> [B, B] broadcast (:989-996): the
    sum over ALL per-sample losses when any window fired, else their mean.  -> (loss, preds_all, preds_specific,
    labels_all, labels_specific)."""
    for m in (model, fe, binary_model, binary_fe):
        m.eval()
    ce = nn.CrossEntropyLoss(reduction="none")
    loss_sum, pa, ps, la, ls = 0.0, [], [], [], []
    for batch in loader:
        images, kin, g, e7, subj = batch[:5]
        y = torch.argmax(select_labels(e7, exp_kwargs).float(), dim=1).view(-1)
        fired = (binary_model(fuse_inputs(images, kin, binary_fe, exp_kwargs)) > 0.5).float()      # [B, 1]
        out = model(fuse_inputs(images, kin, fe, exp_kwargs))
        loss = ce(out, (y - 1).clamp(min=0)) * fired                                                # [B] * [B, 1] -> [B, B]
        loss = loss.sum() / fired.sum() if fired.sum() > 0 else loss.mean()
        loss_sum += loss.item()
        pred = torch.argmax(torch.softmax(out, dim=1), dim=1) + 1
        f = fired.reshape(-1) > 0
        pred = torch.where(f, pred, torch.zeros_like(pred))
        pa += pred.tolist(); la += y.tolist()
        sel = f & (y > 0)
        ps += pred[sel].tolist(); ls += y[sel].tolist()
    return loss_sum / len(loader), pa, ps, la, ls
