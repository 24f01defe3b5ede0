"""Host-side scores from the confusion counts produced on the device by K3.

The reference calls sklearn's f1_score / accuracy_score / jaccard_score / confusion_matrix on host
copies of every batch (MED/modeling/modeling_utils.py:377-381, 519-528, 782-786).  Here the device
kernels return the (at most 8x8) count matrix and these closed forms give the same numbers:
sklearn computes every one of these scores from exactly these counts, with 0/0 := 0.
"""
from __future__ import annotations

import numpy as np


def _div(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    out = np.zeros_like(a)
    np.divide(a, b, out=out, where=b != 0)
    return out


def present_labels(cm: np.ndarray) -> np.ndarray:
    """sklearn scores only the labels that occur in y_true or y_pred."""
    cm = np.asarray(cm)
    return np.flatnonzero((cm.sum(0) + cm.sum(1)) > 0)


def per_class(cm: np.ndarray):
    cm = np.asarray(cm, dtype=np.float64)
    tp = np.diag(cm)
    fp = cm.sum(0) - tp
    fn = cm.sum(1) - tp
    return tp, fp, fn, cm.sum(1)


def f1_binary(cm, pos: int = 1) -> float:
    tp, fp, fn, _ = per_class(cm)
    return float(_div(2 * tp[pos], 2 * tp[pos] + fp[pos] + fn[pos]))


def jaccard_binary(cm, pos: int = 1) -> float:
    tp, fp, fn, _ = per_class(cm)
    return float(_div(tp[pos], tp[pos] + fp[pos] + fn[pos]))


def accuracy(cm) -> float:
    cm = np.asarray(cm, dtype=np.float64)
    return float(_div(np.trace(cm), cm.sum()))


def _averaged(values, support, labels, average):
    values, support = values[labels], support[labels]
    if average == "macro":
        return float(values.mean()) if len(values) else 0.0
    tot = support.sum()
    return float((values * support).sum() / tot) if tot > 0 else 0.0


def f1_avg(cm, average: str) -> float:
    tp, fp, fn, sup = per_class(cm)
    return _averaged(_div(2 * tp, 2 * tp + fp + fn), sup, present_labels(cm), average)


def jaccard_avg(cm, average: str) -> float:
    tp, fp, fn, sup = per_class(cm)
    return _averaged(_div(tp, tp + fp + fn), sup, present_labels(cm), average)


def sklearn_cm(cm) -> np.ndarray:
    """confusion_matrix(y, p) without ``labels=``: rows / columns of the labels that are present."""
    cm = np.asarray(cm)
    lab = present_labels(cm)
    return cm[np.ix_(lab, lab)].astype(np.int64)


def binarise(cm) -> np.ndarray:
    """6-class (label, pred) counts -> error / no-error counts (label != 0, pred != 0)."""
    cm = np.asarray(cm, dtype=np.int64)
    return np.array([[cm[0, 0], cm[0, 1:].sum()], [cm[1:, 0].sum(), cm[1:, 1:].sum()]], dtype=np.int64)
