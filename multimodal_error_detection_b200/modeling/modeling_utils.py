"""Drop-in for the hot-path functions of the reference ``MED/modeling/modeling_utils.py``.

Same function names, argument order and return tuples as the reference (file:line cited per
function).  What changes is where the work happens:

* batches are built on the device by the fused gather/standardise kernel (K1) from the resident frame
  table -- no per-sample ``__getitem__``, no collate, no 42 MB host->device copy per batch;
* the FeatureExtractor runs on the b200med GEMM kernels (K2);
* loss, its gradient, predictions and confusion counts come from ONE fused kernel (K3); per-batch
  losses / counts stay on the device and are read back ONCE per epoch (the reference syncs 1 + 10
  (+3B) times per batch, modeling_utils.py:366, 377-392);
* Adam is one fused kernel over a flat parameter buffer, after one gradient all-reduce when
  data-parallel.

Out of scope: Siamese loops, TransSVNet / COG loops, mlflow retrieval.
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import pandas as pd
import torch
import torch.nn as nn

from .. import metrics as M
from .. import ops
from ..dataset.CustomWindowDataset import DeviceWindowLoader
from ..optim import FusedAdam
from ..table import cuda_device
from .models import CNN, LSTM, FeatureExtractor
from .models_TCN import MultiStageModel

ERROR_COLUMNS = {"No Error": 0, "Out_Of_View": 1, "Multiple_Attempts": 2, "Needle_Position": 3,
                 "Out_Of_View_Multiple_Attempts": 4, "Multiple_Attempts_Needle_Position": 5,
                 "global": -1, "all_errors": [0, 1, 2, 3, 4, 5]}


# =====================================================================================================
# criteria: callable like the torch modules the reference builds, backed by the fused K3 kernel
# =====================================================================================================
class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, pos_weight, out=None):
        r = ops.bce_logits(logits.detach().contiguous().float(), labels.contiguous().float(), pos_weight, want_probs=True, out=out)
        ctx.set_materialize_grads(False)      # no zero-filled gradients for the probs / preds / counts outputs (three fills per step)
        ctx.save_for_backward(r["dlogits"])
        ctx.shape = logits.shape
        ctx.mark_non_differentiable(r["probs"], r["preds"], r["counts"])
        return r["loss"].reshape(()), r["probs"], r["preds"], r["counts"]

    @staticmethod
    def backward(ctx, gl, *_):
        (dl,) = ctx.saved_tensors
        return (dl * gl).reshape(ctx.shape), None, None, None


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, mask, target_shift, reduction, pred_shift, pred_mask_mode, cm_classes):
        r = ops.ce_logits(logits.detach().contiguous().float(), target.to(torch.int32).contiguous(), weight, mask,
                          target_shift=target_shift, reduction=reduction, want_probs=True, pred_shift=pred_shift,
                          pred_mask_mode=pred_mask_mode, cm_classes=cm_classes)
        ctx.set_materialize_grads(False)      # no zero-filled gradients for the probs / preds / counts outputs (three fills per step)
        ctx.save_for_backward(r["dlogits"])
        ctx.mark_non_differentiable(r["probs"], r["preds"], r["cm"])
        return r["loss"].reshape(()), r["probs"], r["preds"], r["cm"]

    @staticmethod
    def backward(ctx, gl, *_):
        (dl,) = ctx.saved_tensors
        return dl * gl, None, None, None, None, None, None, None, None


class _FrameCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, e):
        r = ops.ce_frame(logits.detach().contiguous().float(), e.contiguous().float())
        ctx.set_materialize_grads(False)      # no zero-filled gradients for the probs / preds / counts outputs (three fills per step)
        ctx.save_for_backward(r["dlogits"])
        ctx.mark_non_differentiable(r["preds"], r["counts"])
        return r["loss"].reshape(()), r["preds"], r["counts"]

    @staticmethod
    def backward(ctx, gl, *_):
        (dl,) = ctx.saved_tensors
        return dl * gl, None


class FusedBCEWithLogitsLoss(nn.Module):
    """``nn.BCEWithLogitsLoss(pos_weight=...)`` (mean reduction) on the K3 kernel.  After a call,
    ``.last`` holds (probs, preds, counts[tn, fp, fn, tp]) of that batch."""

    def __init__(self, pos_weight=None):
        super().__init__()
        self.pos_weight = None if pos_weight is None else float(pos_weight)
        self.last = None
        self.static_out = None        # engine.WindowTrainStep: persistent loss / probs / preds / counts buffers the kernel fills

    def forward(self, outputs, labels):
        loss, probs, preds, counts = _BCEFn.apply(outputs, labels, 1.0 if self.pos_weight is None else self.pos_weight,
                                                  self.static_out)
        self.last = (probs, preds, counts)
        return loss


class FusedCrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(weight=..., reduction=...)`` on the K3 kernel for class-index targets
    (window path) and for the frame path's two-column soft targets (see :func:`compute_loss`)."""

    def __init__(self, weight=None, reduction="mean"):
        super().__init__()
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32)
        self.reduction = reduction
        self.last = None

    def forward(self, outputs, labels, mask=None, target_shift=0, reduction=None, pred_shift=0, pred_mask_mode=0,
                cm_classes=None):
        w = None if self.weight is None else self.weight.to(outputs.device)
        if reduction is None and self.reduction not in ("mean", "sum"):
            raise NotImplementedError(f"FusedCrossEntropyLoss(reduction={self.reduction!r}): the K3 kernel returns a reduced loss "
                                      "(mean / sum / masked mean); per-sample losses are not produced")
        red = {"mean": 0, "sum": 2}[self.reduction] if reduction is None else reduction
        loss, probs, preds, cm = _CEFn.apply(outputs, labels, w, mask, target_shift, red, pred_shift, pred_mask_mode,
                                             cm_classes or outputs.shape[1])
        self.last = (probs, preds, cm)
        return loss


def _bce_pos_weight(criterion) -> float:
    pw = getattr(criterion, "pos_weight", None)
    return 1.0 if pw is None else float(pw)


# =====================================================================================================
# glue with the reference's signatures
# =====================================================================================================
def define_inputs(images, kinematics, feature_extractor, exp_kwargs: dict, device) -> torch.Tensor:
    """Reference modeling_utils.py:19-84: FE on the image stream, concat with the kinematics on the
    feature axis, permute to [B, F, W] (COG is out of scope, so the permute always happens)."""
    dt = exp_kwargs["data_type"]
    if dt == "multimodal":
        from ..heads import concat_features
        feats = feature_extractor(images.to(device))
        inputs = concat_features(feats, kinematics.to(device)).permute(0, 2, 1)
    elif dt == "kinematics":
        inputs = kinematics.permute(0, 2, 1).to(device)
    elif dt == "video":
        images = images.to(device)
        inputs = (images if exp_kwargs["video_dims"] == 2048 else feature_extractor(images).float()).permute(0, 2, 1)
    else:
        raise ValueError(f"Data type {dt} is not supported.")
    if inputs.size(0) == 0:
        raise ValueError("Inputs tensor is empty. Check the data loader and the inputs.")
    return inputs


def define_error_labels(e_labels: torch.Tensor, exp_kwargs: dict) -> torch.Tensor:
    """Reference modeling_utils.py:137-191."""
    if "error_type" not in exp_kwargs:
        raise ValueError("error_type must be defined in exp_kwargs.")
    if exp_kwargs["error_type"] not in ERROR_COLUMNS:
        raise ValueError(f"Error type {exp_kwargs['error_type']} is not supported. Supported error types are: "
                         f"{list(ERROR_COLUMNS.keys())}.")
    col = ERROR_COLUMNS[exp_kwargs["error_type"]]
    if exp_kwargs["dataset_type"] == "window":
        return e_labels[:, col]
    if exp_kwargs["dataset_type"] == "frame":
        return e_labels[:, :, col]
    raise ValueError(f"Dataset type {exp_kwargs['dataset_type']} is not supported.")


def instantiate_model(exp_kwargs: dict, in_features: int, window_size: int, device=None) -> nn.Module:
    """Reference modeling_utils.py:3043-3117 (heads on the hot path only)."""
    name = exp_kwargs["model_name"]
    n_out = exp_kwargs["out_features"] if "out_features" in exp_kwargs else 1
    if name == "SimpleCNN":
        return CNN(in_features=in_features, window_size=window_size, n_classes=n_out)
    if name == "SimpleLSTM":
        return LSTM(in_features=in_features, window_size=window_size, hidden_size=exp_kwargs["hidden_size"],
                    num_layers=exp_kwargs["num_layers"], n_classes=n_out)
    if name == "TeCNo":
        return MultiStageModel(exp_kwargs["mstcn_stages"], exp_kwargs["mstcn_layers"], exp_kwargs["mstcn_f_maps"],
                               exp_kwargs["mstcn_f_dim"], exp_kwargs["out_features"], exp_kwargs["mstcn_causal_conv"])
    raise ValueError(f"Model {name} is not supported.")


def _configure_precision(model, exp_kwargs: dict, precision: str) -> None:
    """``exp_kwargs['precision']`` -> the head: "fp32" (exact-math kernels, 1e-5 bar) or "bf16" (tcgen05 kernels, 2e-2 bar).
    Every layer of every head runs on b200med kernels in both modes; there is no other implementation to select."""
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision {precision!r} is not supported. Supported: 'fp32', 'bf16'.")
    if precision == "bf16" and not ops.has_tcgen05():
        raise RuntimeError("precision='bf16' needs the tcgen05 kernels (compute capability 10.x)")
    if isinstance(model, (MultiStageModel, LSTM)):
        model.precision = precision


def define_model_objects(exp_kwargs: dict, in_features_dict: dict, device, class_counts: tuple, window_size: int = 0):
    """Reference modeling_utils.py:194-262 -> (feature_extractor, model, criterion, optimizer, scheduler).

    Seed 42, head built BEFORE the feature extractor, both on the host RNG (that order fixes the
    weights bit for bit), then moved to the GPU.  ``exp_kwargs['precision']``
    (optional, absent in the reference) selects "fp32" (default, 1e-5 parity) or "bf16" (tcgen05)."""
    device = torch.device(device) if device is not None else cuda_device()
    if device.type != "cuda":
        raise RuntimeError("b200med runs on CUDA devices only (no CPU fallback)")
    precision = exp_kwargs.get("precision", "fp32")
    torch.manual_seed(42)
    model = instantiate_model(exp_kwargs, in_features_dict[exp_kwargs["data_type"]], window_size, device).to(device)
    _configure_precision(model, exp_kwargs, precision)
    if exp_kwargs["data_type"] != "kinematics":
        feature_extractor = FeatureExtractor(input_dim=2048, output_dim=exp_kwargs["video_dims"], hidden_dims=[512, 256],
                                             precision=precision).to(device)
        params = list(feature_extractor.parameters()) + list(model.parameters())
    else:
        feature_extractor, params = None, list(model.parameters())
    optimizer = FusedAdam(params, lr=exp_kwargs["lr"], weight_decay=exp_kwargs["weight_decay"]).prepare()
    print("Number of parameters to optimize:", sum(p.numel() for p in params if p.requires_grad))

    criterion = None
    if exp_kwargs["pos_weight"]:
        if exp_kwargs["error_type"] == "global":
            pw = torch.tensor(float(class_counts[0]) / float(class_counts[1]), dtype=torch.float32) \
                if not torch.is_tensor(class_counts[0]) else (class_counts[0] / class_counts[1]).float()
            criterion = FusedBCEWithLogitsLoss(pos_weight=float(pw))
        elif exp_kwargs["error_type"] == "all_errors":
            criterion = FusedCrossEntropyLoss(weight=torch.tensor([float(c) for c in class_counts], dtype=torch.float32))
    elif exp_kwargs["dataset_type"] == "window":
        if exp_kwargs["error_type"] == "global":
            criterion = FusedBCEWithLogitsLoss()
        elif exp_kwargs["error_type"] == "all_errors":
            criterion = FusedCrossEntropyLoss()
    elif exp_kwargs["dataset_type"] == "frame":
        criterion = FusedCrossEntropyLoss(reduction="none" if exp_kwargs["error_type"] == "sequential" else "mean")

    scheduler = (torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=exp_kwargs["n_epochs"], eta_min=1e-6)
                 if exp_kwargs["lr_scheduler"] else None)
    return feature_extractor, model, criterion, optimizer, scheduler


def compute_loss(outputs: torch.Tensor, e_labels: torch.Tensor, criterion, dataset_type: str):
    """Reference modeling_utils.py:265-297 -> (loss, outputs).  Window: squeeze dim 1 and apply the
    criterion.  Frame: CE of every stage against the soft targets [1-e, e], mean over stages -- one
    K3 launch for all stages."""
    if dataset_type == "window":
        if outputs.dim() > 1 and outputs.size(1) == 1:
            outputs = outputs.squeeze(1)
        if isinstance(criterion, (FusedBCEWithLogitsLoss, nn.BCEWithLogitsLoss)):
            if not isinstance(criterion, FusedBCEWithLogitsLoss):
                criterion = FusedBCEWithLogitsLoss(_bce_pos_weight(criterion))
            return criterion(outputs, e_labels), outputs
        if isinstance(criterion, nn.CrossEntropyLoss):
            criterion = FusedCrossEntropyLoss(criterion.weight, criterion.reduction)
        return criterion(outputs, e_labels.long()), outputs
    if dataset_type == "frame":
        # the frame kernel is the reference's loss as written (:278-295): 2 classes, soft targets [1-e, e], unweighted mean
        if outputs.shape[-2] != 2:
            raise NotImplementedError(f"frame loss: {outputs.shape[-2]} output classes (the reference's soft-target loss has 2)")
        if getattr(criterion, "weight", None) is not None or getattr(criterion, "reduction", "mean") not in ("mean", "none"):
            raise NotImplementedError("frame loss: class weights / sum reduction are not part of the reference's frame path")
        loss, preds, counts = _FrameCEFn.apply(outputs, e_labels.reshape(-1).to(outputs.device))
        if hasattr(criterion, "last"):
            criterion.last = (None, preds, counts)
        compute_loss.last_frame = (preds, counts)
        return loss, outputs
    raise ValueError(f"Dataset type {dataset_type} is not supported.")


# =====================================================================================================
# epoch bookkeeping: everything stays on the device until the epoch ends
# =====================================================================================================
class _HostLossRing:
    """Per-step device -> host read of the loss with a lag of one step: push() enqueues an asynchronous copy of this
    step's loss into pinned memory and then reads (on the host) the PREVIOUS step's value, whose copy has had a whole
    step to land.  `values` ends up with every step's loss, in order."""

    def __init__(self, device):
        self.buf = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.values = []

    def push(self, loss: torch.Tensor):
        slot = self.k & 1
        self.buf[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.ev[slot].record()
        if self.k > 0:
            self._read(slot ^ 1)
        self.k += 1

    def _read(self, slot):
        self.ev[slot].synchronize()
        self.values.append(float(self.buf[slot]))

    def drain(self):
        if self.k > 0:
            self._read((self.k - 1) & 1)


class _EpochLog:
    def __init__(self):
        self.losses, self.counts, self.extra = [], [], {}

    def add(self, loss, counts=None, **kw):
        self.losses.append(loss.detach().reshape(1))
        if counts is not None:
            self.counts.append(counts.reshape(1, -1))
        for k, v in kw.items():
            self.extra.setdefault(k, []).append(v)

    def losses_host(self) -> np.ndarray:
        return torch.cat(self.losses).cpu().numpy().astype(np.float64) if self.losses else np.zeros(0)

    def counts_host(self) -> np.ndarray:
        return torch.cat(self.counts).cpu().numpy() if self.counts else np.zeros((0, 4), dtype=np.int64)

    def cat_host(self, key) -> np.ndarray:
        xs = self.extra.get(key, [])
        return torch.cat([x.reshape(-1) for x in xs]).cpu().numpy() if xs else np.zeros(0)


def _batch_scores(counts4: np.ndarray):
    """Per-batch (f1, f1_weighted, acc, jaccard, sklearn-style cm contribution) from (tn, fp, fn, tp)."""
    cm = np.asarray(counts4, dtype=np.int64).reshape(2, 2)
    small = M.sklearn_cm(cm)
    # reference quirk: `train_cm += confusion_matrix(...)` broadcasts a 1x1 matrix over the 2x2 accumulator
    contrib = cm if small.shape == (2, 2) else np.full((2, 2), int(small.sum()), dtype=np.int64)
    return M.f1_binary(cm), M.f1_avg(cm, "weighted"), M.accuracy(cm), M.jaccard_binary(cm), contrib


def _epoch_scores(counts: np.ndarray):
    """(sum over batches of (f1, f1_weighted, acc, jaccard), summed sklearn-style cm) for an epoch's per-batch binary counts
    [n, 4] = (tn, fp, fn, tp): `_batch_scores` for all batches at once (the per-batch loop cost 0.1 ms per batch on the host,
    2 ms of every 20-step epoch).  Same float64 formulas element by element, sums taken batch after batch like the
    reference's `train_f1 += f1_score(...)` (MED/modeling/modeling_utils.py:377-381, 398-402)."""
    c = np.asarray(counts, dtype=np.int64).reshape(-1, 4)
    if c.shape[0] == 0:
        return np.zeros(4), np.zeros((2, 2), dtype=int)
    tn, fp, fn, tp = (c[:, i].astype(np.float64) for i in range(4))

    def div(a, b):
        out = np.zeros_like(a)
        np.divide(a, b, out=out, where=b != 0)
        return out
    f1_pos = div(2 * tp, 2 * tp + fp + fn)
    f1_neg = div(2 * tn, 2 * tn + fn + fp)
    sup_neg, sup_pos = tn + fp, fn + tp
    # labels absent from y_true and y_pred carry zero support and a 0/0 := 0 score: they add exactly +0.0, as in M.f1_avg
    f1w = div(f1_neg * sup_neg + f1_pos * sup_pos, sup_neg + sup_pos)
    acc = div(tn + tp, tn + fp + fn + tp)
    jac = div(tp, tp + fp + fn)
    tot = np.array([np.cumsum(v)[-1] for v in (f1_pos, f1w, acc, jac)])          # sequential sums, batch after batch
    # reference quirk: `train_cm += confusion_matrix(...)` broadcasts a 1x1 matrix (one label present) over the 2x2 accumulator
    present_neg = (tn + fn + tn + fp) > 0
    present_pos = (fp + tp + fn + tp) > 0
    both = present_neg & present_pos
    cm = np.zeros((2, 2), dtype=np.int64)
    cm += c[both].sum(0).reshape(2, 2)
    cm += int(c[~both].sum())
    return tot, cm.astype(int)


def _set_train(model, feature_extractor, exp_kwargs, train: bool):
    mods = [model] if exp_kwargs["data_type"] == "kinematics" else [feature_extractor, model]
    for m in mods:
        m.train(train)


def _dp_world() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class _EarlyExchange:
    """Data-parallel gradient exchange in two pieces (see models.EARLY_EXCHANGE_HOOK): everything but the FeatureExtractor's
    first layer is all-reduced UNDER that layer's weight-gradient GEMM, the first layer's own gradient after the backward."""

    def __init__(self, optimizer):
        self.opt = optimizer
        self.work = None            # async all-reduce of the early range
        self.split = None           # element offset in chunk 0 where the early range starts

    def hook(self, params, grads):
        import torch.distributed as dist
        from .. import lstm_stack
        opt = self.opt
        opt.prepare()
        if len(opt.chunks) != 1:
            return grads            # several chunks: leave the exchange to _allreduce_grads
        chunk = opt.chunks[0]
        slots = {id(p): (o, n) for p, o, n in chunk.members}
        first = [id(params[0]), id(params[1])]
        if any(i not in slots for i in first):
            return grads
        end_first = max(slots[i][0] + slots[i][1] for i in first)
        split = (end_first + 3) // 4 * 4
        if any(o < split for i, (o, n) in slots.items() if i not in first):
            return grads            # the first layer does not lead the flat buffer: one exchange after the backward
        lstm_stack.join_pending()   # LSTM weight gradients still on their side stream
        dst, src = [], []
        out = list(grads)
        for k in range(2, len(params)):                    # the FeatureExtractor's other layers: computed a moment ago
            if grads[k] is None or id(params[k]) not in slots:
                continue
            o, n = slots[id(params[k])]
            view = chunk.grad[o:o + n].view_as(params[k])
            dst.append(view); src.append(grads[k].detach())
            out[k] = view
        seen = set(id(q) for q in params)
        for p, o, n in chunk.members:                      # the head: its gradients already sit in .grad
            if id(p) in seen or p.grad is None:
                continue
            view = chunk.grad[o:o + n].view_as(p)
            if p.grad.data_ptr() != view.data_ptr():
                dst.append(view); src.append(p.grad.detach())
                p.grad = view
        if dst:
            torch._foreach_copy_(dst, src)
        self.split = split
        self.work = dist.all_reduce(chunk.grad[split:], op=dist.ReduceOp.SUM, async_op=True)
        return out

    def finish(self) -> bool:
        """All-reduce the late range and join the early one; False when the hook did not run (nothing exchanged yet)."""
        import torch.distributed as dist
        if self.work is None:
            return False
        self.opt._refresh_active()
        dist.all_reduce(self.opt.chunks[0].grad[:self.split], op=dist.ReduceOp.SUM)
        self.work.wait()
        self.work = None
        return True


_EARLY = {"state": None}


def _backward(loss, optimizer=None, weight: float = 1.0, exchange: bool = True):
    """loss.backward() inside a train loop: the LSTM weight-gradient GEMMs stay on their side stream past the end of the
    backward (lstm_stack.DEFER_JOIN); the next consumer of the gradients (_allreduce_grads / optimizer.step, both through
    FusedAdam._refresh_active) joins it.  ``weight``: data-parallel share of this rank's batch (uneven shards of a short
    batch), applied to the gradient only -- the reported loss stays the rank's own mean.  With several ranks the gradient
    exchange starts INSIDE the backward (:class:`_EarlyExchange`) unless ``exchange`` is False (collective-free warm-up steps)."""
    from .. import lstm_stack
    from . import models as _models
    fused = isinstance(optimizer, FusedAdam)
    lstm_stack.DEFER_JOIN = fused      # only FusedAdam joins the side stream before using the gradients
    early = None
    # the peer-memory exchange (one kernel over NVLink, ~20 us) needs no head start; the NCCL one is split in two
    peer = fused and getattr(optimizer, "_peer", None) is not None
    if exchange and fused and not peer and _dp_world() > 1 and os.environ.get("B200MED_EARLY_EXCHANGE", "1") != "0":
        early = _EARLY["state"] = _EarlyExchange(optimizer)
        _models.EARLY_EXCHANGE_HOOK = early.hook
    try:
        (loss if weight == 1.0 else loss * weight).backward()
    finally:
        lstm_stack.DEFER_JOIN = False
        _models.EARLY_EXCHANGE_HOOK = None


def _setup_exchange(optimizer):
    """Entry of every train loop (all ranks pass here together, outside any graph capture): with several ranks, map the flat
    gradient buffer into peer memory once so that the per-step exchange is the one-kernel NVLink path (parallel.PeerAllReduce)."""
    if _dp_world() > 1 and isinstance(optimizer, FusedAdam) and torch.cuda.is_available() \
            and not torch.cuda.is_current_stream_capturing():
        optimizer.enable_peer_exchange()


def _allreduce_grads(optimizer):
    """Data-parallel exchange: a sum all-reduce of the flat gradient buffer -- in two pieces when the
    backward started it early (everything but the FeatureExtractor's first layer travels under that layer's weight-gradient
    GEMM), else in one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        early, _EARLY["state"] = _EARLY["state"], None
        if early is None or early.opt is not optimizer or not early.finish():
            optimizer._refresh_active()                      # gradients of first-time parameters move into the chunks
            if not (isinstance(optimizer, FusedAdam) and optimizer.peer_all_reduce()):
                for buf in optimizer.grad_buffers():
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        optimizer.grad_scale = 1.0 / dist.get_world_size()


def _window_batches(loader, exp_kwargs, device, image_dtype):
    """Yield (images, kinematics, g, e7, host_idx) per batch.  DeviceWindowLoader: one K1 launch per
    batch straight from the resident table; anything else: the reference's tuple protocol."""
    if isinstance(loader, DeviceWindowLoader):
        ds = loader.dataset
        need_img = exp_kwargs["data_type"] != "kinematics"
        for idx, n_global in loader.index_batches_with_global():
            if idx.numel() == 0:
                continue
            _window_batches.dp_weight = loader.dp_weight(idx.numel(), n_global)
            didx = idx.pin_memory().to(device, non_blocking=True)      # pinned host -> device, 8 B per window
            images, kin = ds.gather_batch(didx, image_dtype=image_dtype if need_img else torch.float32,
                                          exact=image_dtype == torch.float32)
            yield images, kin, ds.g_labels_data.index_select(0, didx), ds.e_labels_data.index_select(0, didx), idx
    else:
        for batch in loader:
            _window_batches.dp_weight = 1.0
            images, kin, g, e7, subject = batch[:5]
            yield images.to(device), kin.to(device), g.to(device), e7.to(device), subject


_window_batches.dp_weight = 1.0      # data-parallel weight of the batch last yielded (see DeviceWindowLoader.dp_weight)


def _subjects(loader, idx):
    if isinstance(loader, DeviceWindowLoader):
        return loader.dataset.subjects_of(idx.tolist())
    return list(idx)


def _image_dtype(feature_extractor):
    return torch.bfloat16 if getattr(feature_extractor, "precision", "fp32") == "bf16" else torch.float32


# =====================================================================================================
# train / validate: binary ("global") window and frame paths
# =====================================================================================================
def train_single_epoch(model, feature_extractor, train_dataloader, criterion, optimizer, scheduler, device, exp_kwargs):
    """Reference modeling_utils.py:300-407.  Returns (loss, f1, f1_weighted, acc, jaccard, cm) -- the MEAN
    over batches of per-batch scores and the summed confusion matrix (:398-402) -- plus
    (probs, preds, labels, subjects) lists when ``return_train_preds``."""
    device = torch.device(device)
    _setup_exchange(optimizer)
    _set_train(model, feature_extractor, exp_kwargs, True)
    if _graph_step_ok(train_dataloader, feature_extractor, criterion, exp_kwargs):
        return _train_epoch_graph(model, feature_extractor, train_dataloader, criterion, optimizer, scheduler, device, exp_kwargs)
    if _frame_graph_ok(train_dataloader, exp_kwargs):
        return _train_epoch_frame_graph(model, feature_extractor, train_dataloader, criterion, optimizer, scheduler, device,
                                        exp_kwargs)
    log = _EpochLog()
    subjects_all = []
    frame = exp_kwargs["dataset_type"] == "frame"
    batches = _frame_batches(train_dataloader, device) if frame else \
        _window_batches(train_dataloader, exp_kwargs, device, _image_dtype(feature_extractor))
    for images, kin, g, e7, who in batches:
        y = define_error_labels(e7, exp_kwargs).float()
        inputs = define_inputs(images, kin, feature_extractor, exp_kwargs, device)
        outputs = model(inputs)
        loss, outputs = compute_loss(outputs, y, criterion, exp_kwargs["dataset_type"])
        optimizer.zero_grad()
        _backward(loss, optimizer, 1.0 if frame else _window_batches.dp_weight)
        _allreduce_grads(optimizer)
        optimizer.step()
        if exp_kwargs.get("host_sync") == "step":
            loss.item()      # the reference's per-batch `train_loss += loss.item()` (:366); default: one read per epoch
        if frame:
            preds, counts = compute_loss.last_frame
            probs = None
        else:
            probs, preds, counts = criterion.last if hasattr(criterion, "last") and criterion.last else _rescore(outputs, y)
        if exp_kwargs["return_train_preds"]:
            log.add(loss, counts, preds=preds, labels=y.reshape(-1), **({} if probs is None else {"probs": probs}))
            subjects_all += _subjects(train_dataloader, who) if not frame else [who] * int(preds.numel())
        else:
            log.add(loss, counts)
    if scheduler is not None:
        scheduler.step()
    n_batches = max(len(log.losses), 1)
    losses, counts = log.losses_host(), log.counts_host()
    tot, cm = _epoch_scores(counts)
    res = (float(losses.sum() / n_batches), *(tot / n_batches).tolist(), cm)
    if exp_kwargs["return_train_preds"]:
        return (*res, log.cat_host("probs").tolist(), log.cat_host("preds").tolist(), log.cat_host("labels").tolist(),
                subjects_all)
    return res


def _graph_step_ok(loader, feature_extractor, criterion, exp_kwargs) -> bool:
    """The CUDA-graph step serves the binary window path over a DeviceWindowLoader.  Default: on in the bf16
    throughput mode, off in the fp32 parity mode; ``exp_kwargs['cuda_graph']`` overrides."""
    want = exp_kwargs.get("cuda_graph", getattr(feature_extractor, "precision", "fp32") == "bf16")
    return bool(want) and isinstance(loader, DeviceWindowLoader) and exp_kwargs["dataset_type"] == "window" \
        and exp_kwargs["error_type"] == "global" and isinstance(criterion, FusedBCEWithLogitsLoss) \
        and exp_kwargs["data_type"] != "kinematics"


def _frame_graph_ok(loader, exp_kwargs) -> bool:
    """Frame path: one CUDA graph per resident video, replayed once per epoch (engine.FrameTrainStep).  Opt-in through
    ``exp_kwargs['cuda_graph']``: capturing costs three eager steps per video, which pays off from the second epoch on."""
    from ..dataset.CustomFrameDataset import FrameLoader
    return bool(exp_kwargs.get("cuda_graph", False)) and isinstance(loader, FrameLoader) and exp_kwargs["dataset_type"] == "frame" \
        and exp_kwargs["error_type"] == "global"


def _train_epoch_frame_graph(model, feature_extractor, loader, criterion, optimizer, scheduler, device, exp_kwargs):
    """train_single_epoch for the frame path with every video's step replayed from its own captured CUDA graph.  Same
    return tuple and the same arithmetic as the eager loop (videos whose capture fails run eagerly, loudly)."""
    from ..engine import FrameTrainStep
    key = (id(loader.dataset), id(model), id(feature_extractor), id(criterion))
    cache = getattr(optimizer, "_b200_frame_steps", None)
    if cache is None or cache["key"] != key:
        cache = optimizer._b200_frame_steps = {"key": key, "steps": {}, "pool": torch.cuda.graph_pool_handle(), "failed": False}
        # data parallel: bring the communicator up once, at the same point on every rank, before any capture records a collective
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(torch.zeros(1, device=device))
            torch.cuda.synchronize()
    log = _EpochLog()
    subjects_all = []
    want_preds = exp_kwargs["return_train_preds"]
    for i in loader.indices():
        step = cache["steps"].get(i)
        if step is None:
            images, kin, g, e7, subject, skill = loader.dataset[i]
            step = FrameTrainStep(images.unsqueeze(0), kin.unsqueeze(0), e7.unsqueeze(0), feature_extractor, model, criterion,
                                  optimizer, exp_kwargs, pool=cache["pool"])
            step.subject = subject
            if not cache["failed"]:
                try:
                    step.capture()
                except Exception as e:      # capture is an optimisation: fall back to eager launches, loudly
                    cache["failed"] = True
                    print(f"b200med: CUDA graph capture of the frame step failed ({type(e).__name__}: {e}); running eagerly")
            cache["steps"][i] = step
        step.run()
        if exp_kwargs.get("host_sync") == "step":
            step.loss.item()
        if want_preds:
            log.add(step.loss.clone(), step.counts.clone(), preds=step.preds.clone(), labels=step.labels.reshape(-1))
            subjects_all += [step.subject] * int(step.preds.numel())
        else:
            log.add(step.loss.clone(), step.counts.clone())
    if scheduler is not None:
        scheduler.step()
    n_batches = max(len(log.losses), 1)
    losses, counts = log.losses_host(), log.counts_host()
    tot, cm = _epoch_scores(counts)
    res = (float(losses.sum() / n_batches), *(tot / n_batches).tolist(), cm)
    if want_preds:
        return (*res, log.cat_host("probs").tolist(), log.cat_host("preds").tolist(), log.cat_host("labels").tolist(),
                subjects_all)
    return res


def _train_epoch_graph(model, feature_extractor, loader, criterion, optimizer, scheduler, device, exp_kwargs):
    """train_single_epoch with every full batch replayed from ONE captured CUDA graph (engine.WindowTrainStep): per
    step the host copies the batch's window indices (pinned, 8 B per window) and replays; the short last batch runs
    eagerly.  Same return tuple and the same arithmetic as the eager loop."""
    from ..engine import WindowTrainStep
    ds = loader.dataset
    # the captured step serves this rank's share of a FULL global batch (data parallel: loader.batch_size is the global
    # batch; every other batch -- the short last one, uneven shards -- runs eagerly)
    world = max(1, int(getattr(loader, "world_size", 1)))
    B = loader.batch_size // world if loader.batch_size % world == 0 else -1
    key = (id(ds), id(model), id(feature_extractor), id(criterion), B)
    stepper = getattr(optimizer, "_b200_stepper", None)
    if stepper is None or stepper.key != key:
        stepper = WindowTrainStep(ds, feature_extractor, model, criterion, optimizer, exp_kwargs, max(B, 1),
                                  prefetch=bool(exp_kwargs.get("prefetch_gather", False)))
        stepper.key = key
        optimizer._b200_stepper = stepper
    log = _EpochLog()
    subjects_all = []
    want_preds = exp_kwargs["return_train_preds"]
    # host_sync == "step": the reference reads the loss on the host every batch (`train_loss += loss.item()`, :366).  Here
    # every step's loss is copied to pinned host memory asynchronously and read one step later, so the read of step k-1
    # overlaps the execution of step k instead of draining the GPU queue.
    step_sync = exp_kwargs.get("host_sync") == "step"
    ring = _HostLossRing(device) if step_sync else None
    def _with_lookahead(it):
        """(batch, next batch or None) pairs."""
        prev = None
        for cur in it:
            if prev is not None:
                yield prev, cur
            prev = cur
        if prev is not None:
            yield prev, None

    for (idx, n_global), nxt_pair in _with_lookahead(p for p in loader.index_batches_with_global() if p[0].numel() > 0):
        n = idx.numel()
        idx_next = None if nxt_pair is None else nxt_pair[0]
        if nxt_pair is not None and nxt_pair[1] != loader.batch_size:
            idx_next = None                               # the next global batch is short: it runs eagerly, no prefetch for it
        if n == B and n_global == loader.batch_size:
            if not (stepper.prefetch and stepper._primed):
                stepper.load(idx.pin_memory())           # start of a sequence (or no prefetch): stage (and gather) this batch
            if stepper.graph is None and not getattr(stepper, "graph_failed", False):
                try:
                    stepper.capture()
                except Exception as e:      # capture is an optimisation: fall back to eager launches, loudly
                    stepper.graphs, stepper.graph_failed = [None, None], True
                    print(f"b200med: CUDA graph capture failed ({type(e).__name__}: {e}); running the step eagerly")
            # prefetch mode: the NEXT full batch is gathered inside this step (under the LSTM recurrence)
            nxt = idx_next.pin_memory() if (stepper.prefetch and idx_next is not None and idx_next.numel() == B) else None
            stepper.run(nxt)
            loss, counts = stepper.loss.clone(), stepper.counts.clone()
            extra = dict(preds=stepper.preds.clone(), labels=stepper.labels.clone(), probs=stepper.probs.clone()) if want_preds else {}
        else:
            didx = idx.pin_memory().to(device, non_blocking=True)
            images, kin = ds.gather_batch(didx, image_dtype=_image_dtype(feature_extractor),
                                          exact=_image_dtype(feature_extractor) == torch.float32)
            y = define_error_labels(ds.e_labels_data.index_select(0, didx), exp_kwargs).float()
            outputs = model(define_inputs(images, kin, feature_extractor, exp_kwargs, device))
            loss, outputs = compute_loss(outputs, y, criterion, "window")
            optimizer.zero_grad()
            _backward(loss, optimizer, loader.dp_weight(n, n_global))
            _allreduce_grads(optimizer)
            optimizer.step()
            probs, preds, counts = criterion.last
            loss = loss.detach().reshape(1)
            extra = dict(preds=preds, labels=y.reshape(-1), probs=probs) if want_preds else {}
        if step_sync:
            ring.push(loss)
        log.add(loss, counts, **extra)
        if want_preds:
            subjects_all += ds.subjects_of(idx.tolist())
    if step_sync:
        ring.drain()
    if scheduler is not None:
        scheduler.step()
    n_batches = max(len(log.losses), 1)
    tot, cm = _epoch_scores(log.counts_host())
    res = (float(log.losses_host().sum() / n_batches), *(tot / n_batches).tolist(), cm)
    if want_preds:
        return (*res, log.cat_host("probs").tolist(), log.cat_host("preds").tolist(), log.cat_host("labels").tolist(), subjects_all)
    return res


def _rescore(outputs, y):
    r = ops.bce_logits(outputs.detach().contiguous().float(), y.contiguous().float(), want_grad=False, want_probs=True)
    return r["probs"], r["preds"], r["counts"]


def _frame_batches(loader, device):
    """One video per step (reference train_frame.ipynb: DataLoader(batch_size=1))."""
    for batch in loader:
        images, kin, g, e7, subject = batch[:5]
        who = subject[0] if isinstance(subject, (tuple, list)) else subject
        yield images.to(device), kin.to(device), g.to(device), e7.to(device), who


def validate_single_epoch(model, feature_extractor, test_dataloader, criterion, device, exp_kwargs):
    """Reference modeling_utils.py:688-790 -> 13-tuple (loss, f1, f1_weighted, acc, jaccard, cm,
    inference_rate, preds, probs, labels, labels_specific, gesture_labels, subjects); scores are POOLED
    over all samples (:782-786)."""
    device = torch.device(device)
    _set_train(model, feature_extractor, exp_kwargs, False)
    log = _EpochLog()
    frame = exp_kwargs["dataset_type"] == "frame"
    subjects_all, labels_all = [], []
    fwd_ms = 0.0
    with torch.no_grad():
        batches = _frame_batches(test_dataloader, device) if frame else \
            _window_batches(test_dataloader, exp_kwargs, device, _image_dtype(feature_extractor))
        for images, kin, g, e7, who in batches:
            y = define_error_labels(e7, exp_kwargs).float()
            inputs = define_inputs(images, kin, feature_extractor, exp_kwargs, device)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            outputs = model(inputs)
            t1.record()
            if frame:
                r = ops.ce_frame(outputs.contiguous().float(), y.reshape(-1), want_grad=False)
                preds, probs = r["preds"], None
                e_rows = e7[0]
                subjects_all += [(who,)] * int(preds.numel())
            else:
                outputs = outputs.squeeze(1) if outputs.dim() > 1 and outputs.size(1) == 1 else outputs
                r = ops.bce_logits(outputs.contiguous().float(), y.contiguous(), _bce_pos_weight(criterion), want_grad=False,
                                   want_probs=True)
                preds, probs = r["preds"], r["probs"]
                e_rows = e7
                subjects_all += _subjects(test_dataloader, who)
            labels_all.append(e_rows)
            log.add(r["loss"], r["counts"], preds=preds, labels=y.reshape(-1), gest=g.reshape(-1).float(),
                    **({} if probs is None else {"probs": probs}))
            last_events = (t0, t1)
    if log.losses:
        torch.cuda.synchronize()
        fwd_ms = last_events[0].elapsed_time(last_events[1])   # device time of the LAST batch's forward (:779)
    n_batches = max(len(log.losses), 1)
    cm = log.counts_host().sum(0).reshape(2, 2) if log.counts else np.zeros((2, 2), dtype=np.int64)
    preds = log.cat_host("preds").tolist()
    probs = log.cat_host("probs").tolist() if not frame else [None] * len(preds)
    e_all = torch.cat(labels_all).cpu() if labels_all else torch.zeros(0, 7)
    return (float(log.losses_host().sum() / n_batches), M.f1_binary(cm), M.f1_avg(cm, "weighted"), M.accuracy(cm),
            M.jaccard_binary(cm), M.sklearn_cm(cm), fwd_ms, preds, probs, [row for row in e_all],
            log.cat_host("labels").tolist(), log.cat_host("gest").tolist(), subjects_all)


# =====================================================================================================
# error-specific (6-class) loops
# =====================================================================================================
def _class_targets(e7, exp_kwargs):
    spec = define_error_labels(e7, exp_kwargs).float()
    return torch.argmax(spec, dim=2 if exp_kwargs["dataset_type"] == "frame" else 1).view(-1)


def _es_summary(cm6: np.ndarray):
    cmb = M.binarise(cm6)
    return (M.f1_binary(cmb), M.f1_avg(cm6, "macro"), M.accuracy(cmb), M.accuracy(cm6), M.jaccard_binary(cmb),
            M.jaccard_avg(cm6, "macro"), M.sklearn_cm(cmb), M.sklearn_cm(cm6))


def train_single_epoch_ES(model, feature_extractor, train_dataloader, criterion, optimizer, scheduler, device, exp_kwargs):
    """Reference modeling_utils.py:410-539 (window path).  The committed code hands CrossEntropyLoss a
    float class index and raises on CPU/CUDA; the index is used as ``long`` here, which is
    what the call means.  Scores are pooled over the epoch (:519-528)."""
    device = torch.device(device)
    _setup_exchange(optimizer)
    _set_train(model, feature_extractor, exp_kwargs, True)
    crit = criterion if isinstance(criterion, FusedCrossEntropyLoss) else FusedCrossEntropyLoss(
        getattr(criterion, "weight", None), getattr(criterion, "reduction", "mean"))
    log = _EpochLog()
    C = None
    cm = None
    for images, kin, g, e7, who in _window_batches(train_dataloader, exp_kwargs, device, _image_dtype(feature_extractor)):
        y = _class_targets(e7, exp_kwargs)
        outputs = model(define_inputs(images, kin, feature_extractor, exp_kwargs, device))
        C = outputs.shape[1]
        loss = crit(outputs, y, cm_classes=max(C, 6))
        optimizer.zero_grad()
        _backward(loss, optimizer, _window_batches.dp_weight)
        _allreduce_grads(optimizer)
        optimizer.step()
        cm = crit.last[2] if cm is None else cm + crit.last[2]
        log.add(loss, None, preds=crit.last[1], labels=y)
    if scheduler is not None:
        scheduler.step()
    n_batches = max(len(log.losses), 1)
    cm6 = cm.cpu().numpy() if cm is not None else np.zeros((6, 6), dtype=np.int64)
    res = (float(log.losses_host().sum()) / n_batches, *_es_summary(cm6))
    if exp_kwargs["return_train_preds"]:
        preds, labels = log.cat_host("preds").astype(int).tolist(), log.cat_host("labels").astype(int).tolist()
        return (*res, [], preds, labels, [int(v != 0) for v in labels], [int(v != 0) for v in preds])
    return res


def validate_single_epoch_ES(model, feature_extractor, test_dataloader, criterion, device, exp_kwargs):
    """Reference modeling_utils.py:793-904 -> 17-tuple."""
    device = torch.device(device)
    _set_train(model, feature_extractor, exp_kwargs, False)
    crit = criterion if isinstance(criterion, FusedCrossEntropyLoss) else FusedCrossEntropyLoss(
        getattr(criterion, "weight", None), getattr(criterion, "reduction", "mean"))
    log = _EpochLog()
    subjects_all, cm, fwd_ms, last_events = [], None, 0.0, None
    with torch.no_grad():
        for images, kin, g, e7, who in _window_batches(test_dataloader, exp_kwargs, device, _image_dtype(feature_extractor)):
            y = _class_targets(e7, exp_kwargs)
            inputs = define_inputs(images, kin, feature_extractor, exp_kwargs, device)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            outputs = model(inputs)
            t1.record()
            last_events = (t0, t1)
            w = None if crit.weight is None else crit.weight.to(device)
            r = ops.ce_logits(outputs.contiguous().float(), y.to(torch.int32), w, None, want_grad=False, want_probs=True,
                              cm_classes=max(outputs.shape[1], 6))
            cm = r["cm"] if cm is None else cm + r["cm"]
            log.add(r["loss"], None, preds=r["preds"], labels=y, probs=r["probs"][:, 1], gest=g.reshape(-1).float())
            subjects_all += _subjects(test_dataloader, who)
    if last_events is not None:
        torch.cuda.synchronize()
        fwd_ms = last_events[0].elapsed_time(last_events[1])
    n_batches = max(len(log.losses), 1)
    cm6 = cm.cpu().numpy() if cm is not None else np.zeros((6, 6), dtype=np.int64)
    preds, labels = log.cat_host("preds").astype(int).tolist(), log.cat_host("labels").astype(int).tolist()
    return (float(log.losses_host().sum()) / n_batches, *_es_summary(cm6), fwd_ms, log.cat_host("probs").tolist(), preds,
            labels, [int(v != 0) for v in labels], [int(v != 0) for v in preds], log.cat_host("gest").tolist(), subjects_all)


# =====================================================================================================
# cascade ("sequential") loops: a frozen binary model gates a 5-class error-type model
# =====================================================================================================
def _seq_summary(cm_all: np.ndarray, cm_spec: np.ndarray):
    return (M.f1_avg(cm_all, "macro"), M.f1_avg(cm_spec, "macro"), M.f1_avg(cm_spec, "weighted"), M.accuracy(cm_all),
            M.accuracy(cm_spec), M.jaccard_avg(cm_all, "macro"), M.jaccard_avg(cm_spec, "macro"),
            M.jaccard_avg(cm_spec, "weighted"), M.sklearn_cm(cm_all), M.sklearn_cm(cm_spec))


def train_single_epoch_Sequential(model, feature_extractor, train_dataloader, criterion, optimizer, device, scheduler,
                                  exp_kwargs):
    """Reference modeling_utils.py:543-684 -> 11-tuple.  Labels 0..5, mask = (label != 0), CE on
    label-1 for the masked rows, sum / mask.sum().  The committed code feeds target -1 for unmasked
    rows (raises off-MPS); those rows carry zero weight, so the target is clamped to 0."""
    device = torch.device(device)
    _setup_exchange(optimizer)
    _set_train(model, feature_extractor, exp_kwargs, True)
    crit = FusedCrossEntropyLoss()
    log = _EpochLog()
    cm = None
    for images, kin, g, e7, who in _window_batches(train_dataloader, exp_kwargs, device, _image_dtype(feature_extractor)):
        y = _class_targets(e7, exp_kwargs)
        mask = (y != 0).float()
        outputs = model(define_inputs(images, kin, feature_extractor, exp_kwargs, device))
        loss = crit(outputs, y, mask=mask, target_shift=-1, reduction=1, pred_shift=1, pred_mask_mode=1, cm_classes=6)
        optimizer.zero_grad()
        _backward(loss, optimizer, _window_batches.dp_weight)
        _allreduce_grads(optimizer)
        optimizer.step()
        cm = crit.last[2] if cm is None else cm + crit.last[2]
        log.add(loss, None)
    if scheduler is not None:
        scheduler.step()
    n_batches = max(len(log.losses), 1)
    cm_all = cm.cpu().numpy() if cm is not None else np.zeros((6, 6), dtype=np.int64)
    cm_spec = cm_all.copy()
    cm_spec[0, :] = 0          # error-specific lists hold only the rows whose true label is an error (:658-662)
    return (float(log.losses_host().sum()) / n_batches, *_seq_summary(cm_all, cm_spec))


def validate_single_epoch_Sequential(model, feature_extractor, binary_model, binary_feature_extractor, test_dataloader,
                                     device, exp_kwargs):
    """Reference modeling_utils.py:907-1053 -> 19-tuple.  The binary model's RAW logit is thresholded
    at 0.5 (:979-980); the loss keeps the reference's [B]*[B,1] broadcast: sum of all per-sample losses if
    any window fired, else their mean (:989-996)."""
    device = torch.device(device)
    for m in (model, feature_extractor, binary_model, binary_feature_extractor):
        m.eval()
    log = _EpochLog()
    cm_all = cm_spec = None
    total_ms, last_w = 0.0, 1
    with torch.no_grad():
        for images, kin, g, e7, who in _window_batches(test_dataloader, exp_kwargs, device, _image_dtype(feature_extractor)):
            y = _class_targets(e7, exp_kwargs)
            fired = (binary_model(define_inputs(images, kin, binary_feature_extractor, exp_kwargs, device)) > 0.5).float().reshape(-1)
            inputs = define_inputs(images, kin, feature_extractor, exp_kwargs, device)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            outputs = model(inputs)
            t1.record()
            r = ops.ce_logits(outputs.contiguous().float(), y.to(torch.int32), None, fired.contiguous(), target_shift=-1,
                              reduction=3, want_grad=False, want_probs=True, pred_shift=1, pred_mask_mode=2, cm_classes=6)
            cm_all = r["cm"] if cm_all is None else cm_all + r["cm"]
            sel = (fired > 0) & (y > 0)
            spec = ops.confusion(y.to(torch.int32)[sel].contiguous(), r["preds"][sel].contiguous(), 6)
            cm_spec = spec if cm_spec is None else cm_spec + spec
            log.add(r["loss"], None, preds=r["preds"], labels=y, sel=sel.float(), probs=r["probs"])
            torch.cuda.synchronize()
            total_ms += t0.elapsed_time(t1)
            last_w = images.shape[1]
    n_batches = max(len(log.losses), 1)
    z = np.zeros((6, 6), dtype=np.int64)
    cm_all = cm_all.cpu().numpy() if cm_all is not None else z
    cm_spec = cm_spec.cpu().numpy() if cm_spec is not None else z
    preds, labels = log.cat_host("preds").astype(int), log.cat_host("labels").astype(int)
    sel = log.cat_host("sel") > 0
    probs = log.cat_host("probs").reshape(len(preds), -1) if len(preds) else np.zeros((0, 5))
    return (float(log.losses_host().sum()) / n_batches, *_seq_summary(cm_all, cm_spec), total_ms / last_w,
            preds.tolist(), preds[sel].tolist(), probs[sel].tolist(), labels.tolist(), labels[sel].tolist(), [], [])


# =====================================================================================================
# frame -> window post-processing, fold aggregation, checkpoint IO
# =====================================================================================================
def window_predictions(predictions, e_labels, gestures, subjects, window_size=10, stride=6, binary=True):
    """Reference modeling_utils.py:2695-2777: re-window frame-level predictions with the walk of
    ``window_data`` (K0 kernel), window value = mean of the frame predictions thresholded ``>= 0.5`` or
    rounded half-to-even (vote kernel).  Subjects are visited in SORTED order (``np.unique``, :2722)."""
    dev = cuda_device()
    subjects = np.asarray(subjects)
    predictions, e_labels, gestures = np.asarray(predictions), np.asarray(e_labels), np.asarray(gestures)
    uniq, inv = np.unique(subjects, return_inverse=True)
    order = np.argsort(inv, kind="stable")
    counts = np.bincount(inv, minlength=len(uniq))
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(dev)
    g = torch.from_numpy(gestures[order].astype(np.float32)).to(dev)
    r = ops.window_index(g, offsets, window_size, stride)
    starts = r["starts"]
    votes = ops.window_vote(torch.from_numpy(predictions[order].astype(np.float32)).to(dev), starts, window_size, binary)
    first = order[starts.cpu().numpy()]
    pw = votes.cpu().to(torch.float32 if binary else torch.float64).reshape(-1, 1)
    ew = torch.tensor(e_labels[first]).reshape(-1, 1)
    gw = torch.tensor(gestures[first]).reshape(-1, 1)
    return pw, ew, gw, pd.DataFrame([str(s) for s in subjects[first]], columns=["subject"])


def frame2window(outs, test_all_preds, test_all_labels, test_all_gest_labels, test_all_subjects, window_size=10, stride=6,
                 binary=True):
    """Reference modeling_utils.py:2780-2825."""
    wp, wl, wg, ws = {}, {}, {}, {}
    for out in outs:
        if out in test_all_preds:
            wp[out], wl[out], wg[out], ws[out] = window_predictions(
                np.array(test_all_preds[out]), np.array(test_all_labels[out]), np.array(test_all_gest_labels[out]),
                np.array(test_all_subjects[out]), window_size=window_size, stride=stride, binary=binary)
    return wp, wl, wg, ws


def _pooled_cm(labels: np.ndarray, preds: np.ndarray, n_classes: int) -> np.ndarray:
    dev = cuda_device()
    t = torch.from_numpy(labels.astype(np.int32)).to(dev)
    p = torch.from_numpy(preds.astype(np.int32)).to(dev)
    return ops.confusion(t, p, n_classes).cpu().numpy()


def _weighted_mean_std(values, weights):
    values = np.asarray(values, dtype=np.float64)
    mean = np.average(values, weights=weights)
    return mean, np.average((values - mean) ** 2, weights=weights) ** 0.5


def compute_window_metrics(outs, test_all_preds, test_all_labels, test_all_gest_labels, test_all_subjects, window_size=10,
                           stride=6, binary=True):
    """Reference modeling_utils.py:2828-2917 -> (summary DataFrame, summed confusion matrix)."""
    wp, wl, wg, ws = frame2window(outs, test_all_preds, test_all_labels, test_all_gest_labels, test_all_subjects,
                                  window_size=window_size, stride=stride, binary=binary)
    f1s, accs, jacs, cms, samples = [], [], [], [], []
    for out in wp:
        preds, labels = wp[out].numpy().flatten(), wl[out].numpy().flatten()
        cm = _pooled_cm(labels, preds, 2 if binary else 6)
        f1s.append(M.f1_binary(cm) if binary else M.f1_avg(cm, "weighted"))
        jacs.append(M.jaccard_binary(cm) if binary else M.jaccard_avg(cm, "weighted"))
        accs.append(M.accuracy(cm))
        cms.append(M.sklearn_cm(cm))
        samples.append(len(wp[out]))
    (mf, sf), (ma, sa), (mj, sj) = (_weighted_mean_std(v, samples) for v in (f1s, accs, jacs))
    summary_df = pd.DataFrame({"F1": [f"{mf:.3f} ± {sf:.3f}"], "Accuracy": [f"{ma:.3f} ± {sa:.3f}"],
                               "Jaccard": [f"{mj:.3f} ± {sj:.3f}"]}, index=["Windowed Metrics"])
    return summary_df, np.sum(np.array(cms), axis=0)


def create_summary_df(LOSO_f1_train, LOSO_f1_test, LOSO_acc_train, LOSO_acc_test, LOSO_jaccard_train, LOSO_jaccard_test,
                      samples_train, samples_test, inference_rates, train_times) -> pd.DataFrame:
    """Reference modeling_utils.py:2979-3025: sample-weighted mean ± std across folds."""
    df = pd.DataFrame(index=["Train", "Test"], columns=["F1", "Accuracy", "Jaccard", "Train Time", "Inference Rate"])
    cells = {("Train", "F1"): (LOSO_f1_train, samples_train), ("Test", "F1"): (LOSO_f1_test, samples_test),
             ("Train", "Accuracy"): (LOSO_acc_train, samples_train), ("Test", "Accuracy"): (LOSO_acc_test, samples_test),
             ("Train", "Jaccard"): (LOSO_jaccard_train, samples_train), ("Test", "Jaccard"): (LOSO_jaccard_test, samples_test)}
    for (row, col), (vals, w) in cells.items():
        mean, std = _weighted_mean_std(vals, w)
        df.loc[row, col] = f"{mean:.3f} ± {std:.3f}"
    df.loc["Train", "Train Time"] = f"{np.mean(train_times):.2f} ± {np.std(train_times):.2f}"
    df.loc["Test", "Train Time"] = np.nan
    df.loc["Train", "Inference Rate"] = np.nan
    df.loc["Test", "Inference Rate"] = f"{np.mean(inference_rates):.2f} ± {np.std(inference_rates):.2f}"
    return df


def soft_vote_ensemble(probs_a, probs_b, labels):
    """Soft vote of two window models + scores (reference ensemble.ipynb cell 6): returns
    (preds, acc, f1, jaccard, cm)."""
    dev = cuda_device()
    to = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float32)).to(dev)
    preds, counts = ops.soft_vote(to(probs_a), to(probs_b), to(labels))
    cm = counts.cpu().numpy().reshape(2, 2)
    return preds.cpu().numpy().astype(int), M.accuracy(cm), M.f1_binary(cm), M.jaccard_binary(cm), cm


def cascade_ensemble(binary_preds, multiclass_preds):
    """Cascade of a binary and a multi-class model (reference ensemble.ipynb cell 15, lines 53-63)."""
    dev = cuda_device()
    to = lambda x: torch.as_tensor(np.asarray(x, dtype=np.int32)).to(dev)
    return ops.cascade(to(binary_preds), to(multiclass_preds)).cpu().numpy()


def roc_auc_score(y_true, y_score) -> float:
    """``sklearn.metrics.roc_auc_score(y_true, y_score)`` for binary labels, as the reference imports it
    (modeling_utils.py:7) and calls it on stored per-sample probabilities (:1124, :1243), computed on the device
    (csrc/metrics.cu: integer rank counting, ties count one half).  Raises ValueError when only one class is present,
    like sklearn."""
    dev = cuda_device()
    to = lambda x: (x.detach() if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float32))).to(dev, torch.float32)
    auc, stats = ops.roc_auc(to(y_score).contiguous(), to(y_true).contiguous())
    host = stats.cpu()
    if int(host[0]) == 0 or int(host[1]) == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return float(auc.cpu())


def save_model(best_model: dict, model_path: str) -> None:
    """Reference modeling_utils.py:3028-3040: ``{'feature_extractor': state_dict, 'model': state_dict}``."""
    torch.save({"feature_extractor": best_model["feature_extractor"], "model": best_model["model"]}, model_path)
    print(f"Model saved to {model_path}")


def _load_blob(model_path: str) -> dict:
    return torch.load(model_path, map_location="cpu", weights_only=False)


def load_model_file(model_path: str, feature_extractor, model, device=None):
    """Load one ``save_model`` file (written by the reference or by this package -- same keys and shapes) into existing
    modules, in place: parameters that FusedAdam re-homed into its flat buffer stay views of it."""
    blob = _load_blob(model_path)
    if feature_extractor is not None and blob.get("feature_extractor") is not None:
        feature_extractor.load_state_dict(blob["feature_extractor"])
    model.load_state_dict(blob["model"])
    return feature_extractor, model


def load_model_local(model_folder: str, out: str, setting: str, exp_kwargs: dict, in_features: int, window_size: int,
                     device) -> tuple:
    """Reference modeling_utils.py:2241-2295: build the head (and the FeatureExtractor unless the data type is
    kinematics or ``video_dims == 2048``), load ``<model_folder>/best_model_<setting>_<out>.pt`` -> (feature_extractor, model).
    (TransSVNet checkpoints load their TeCNo trunk, :2262-2267.)"""
    device = torch.device(device) if device is not None else cuda_device()
    kw = dict(exp_kwargs, model_name="TeCNo") if exp_kwargs["model_name"] == "TransSVNet" else exp_kwargs
    model = instantiate_model(kw, in_features, window_size).to(device)
    precision = exp_kwargs.get("precision", "fp32")
    _configure_precision(model, exp_kwargs, precision)
    feature_extractor = None
    if exp_kwargs["data_type"] != "kinematics" and exp_kwargs["video_dims"] != 2048:
        feature_extractor = FeatureExtractor(input_dim=2048, output_dim=exp_kwargs["video_dims"], hidden_dims=[512, 256],
                                             precision=precision).to(device)
    best_model = _load_blob(os.path.join(model_folder, f"best_model_{setting}_{out}.pt"))
    model.load_state_dict(best_model["model"])
    if feature_extractor is not None:
        feature_extractor.load_state_dict(best_model["feature_extractor"])
    return feature_extractor, model


def load_binary_model_local(model_folder: str, model_name: str, outs: list, exp_kwargs: dict, device):
    """Reference modeling_utils.py:2298-2329: the frozen binary models of the cascade, one per fold -- always
    ``LSTM(58, 10, hidden 128, 3 layers, 1 class)`` + FeatureExtractor -- read from
    ``models/<data_type>/<frequency>Hz/<model_name>/best_model_LOSO_<out>.pt`` (the reference overwrites its
    ``model_folder`` argument with that relative path, :2304; kept) -> (model_dict, fe_dict)."""
    device = torch.device(device) if device is not None else cuda_device()
    model_folder = f'models/{exp_kwargs["data_type"]}/{exp_kwargs["frequency"]}Hz/{model_name}/'
    precision = exp_kwargs.get("precision", "fp32")
    model_dict, fe_dict = {}, {}
    for out in outs:
        model = LSTM(in_features=58, window_size=10, hidden_size=128, num_layers=3, n_classes=1).to(device)
        _configure_precision(model, exp_kwargs, precision)
        feature_extractor = FeatureExtractor(input_dim=2048, output_dim=exp_kwargs["video_dims"], hidden_dims=[512, 256],
                                             precision=precision).to(device)
        best_model = _load_blob(os.path.join(model_folder, f"best_model_LOSO_{out}.pt"))
        feature_extractor.load_state_dict(best_model["feature_extractor"])
        model.load_state_dict(best_model["model"])
        model_dict[out], fe_dict[out] = model, feature_extractor
    print(f"Loaded binary models from {model_folder}.")
    return model_dict, fe_dict


def create_binary_mask(preds_binary: dict, subjects: dict, out: str, fold_data_path: str, exp_kwargs: dict = None):
    """Reference modeling_utils.py:2920-2976: the binary model's predictions of fold ``out`` as the cascade's mask, with
    the Needle-Drop positions recorded in the fold's ``mask_position_ND_<subject>.pth`` files removed (when
    ``delete_ND``) so that the mask lines up with the error-specific predictions -> (binary_mask, binary_subjects).
    Host-side index bookkeeping over a fold's prediction list (10^3..10^5 entries, once per fold)."""
    binary_mask = np.array(preds_binary[out])
    binary_subjects = np.array(subjects[out])
    nd_files = []
    if exp_kwargs["delete_ND"]:
        for file in os.listdir(fold_data_path):
            if file.startswith("mask_position_ND_") and file.endswith(".pth"):
                subject = file.split(".")[0].replace("mask_position_ND_", "")
                nd_files.append((subject, torch.load(os.path.join(fold_data_path, file))))
    print(binary_mask.shape)
    for subject, mask_position_ND in nd_files:
        rows = np.where(binary_subjects == subject)[0]
        if len(rows) == 0:
            continue
        print(f"Found mask_position_ND for subject {subject} in {out} set.")
        expanded = np.zeros_like(binary_mask, dtype=bool)
        expanded[rows] = np.asarray(mask_position_ND, dtype=bool)
        binary_mask, binary_subjects = binary_mask[~expanded], binary_subjects[~expanded]
    print(binary_mask.shape)
    return binary_mask, binary_subjects
