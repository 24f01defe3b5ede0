"""Oracle: window construction and label transforms (integer / index work, bit-exact bar).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pure numpy + Python loops.
"""
from __future__ import annotations

import numpy as np


def subjects_in_order(subject_names):
    """Distinct subjects in order of first appearance, and each subject's frame rows.

    Follows ``window_data`` (reference MED/dataset/dataset_utils.py:193-194), which uses
    pandas ``.unique()`` (appearance order) and a boolean match per subject."""
    names = np.asarray(subject_names)
    seen, order = set(), []
    for s in names.tolist():
        if s not in seen:
            seen.add(s)
            order.append(s)
    rows = {s: np.flatnonzero(names == s) for s in order}
    return order, rows


def window_starts_subject(g_subject: np.ndarray, W: int, S: int) -> list:
    """Local start positions of the windows of ONE subject.

    Reference MED/dataset/dataset_utils.py:211-240:
      * begin at the first frame whose gesture is non-zero (:211-212);
      * loop while ``start < n - W`` -- strict, so the last legal window is never emitted (:214);
      * compare only the two END-POINT gestures of the candidate window (:220-221);
      * mismatch -> advance by one frame (:226); match -> emit, advance by the stride (:239).
    """
    n = len(g_subject)
    nz = np.flatnonzero(g_subject != 0)
    if len(nz) == 0:
        raise IndexError("subject has no non-zero gesture")  # reference: gesture_indices[0] raises
    pos = int(nz[0])
    out = []
    while pos < n - W:
        if g_subject[pos] != g_subject[pos + W - 1]:
            pos += 1
            continue
        out.append(pos)
        pos += S
    return out


def window_starts(g: np.ndarray, subject_names, W: int, S: int):
    """All windows of a flat table.  Returns (rows [n, W] int64 global frame rows,
    subject name per window).  ``rows[:, 0]`` is the global start when subjects are contiguous
    (SURVEY Appendix A-1)."""
    g = np.asarray(g).reshape(-1)
    order, rows = subjects_in_order(subject_names)
    win_rows, win_subj = [], []
    for s in order:
        idx = rows[s]
        for p in window_starts_subject(g[idx], W, S):
            win_rows.append(idx[p:p + W])
            win_subj.append(s)
    if win_rows:
        return np.stack(win_rows).astype(np.int64), win_subj
    return np.zeros((0, W), dtype=np.int64), win_subj


def window_data(image, kin, g, e5, subject_names, W=10, S=6):
    """Materialised windows like the reference returns them (dataset_utils.py:229-258):
    frames of each window, and the gesture / error labels of its FIRST frame (:232-233)."""
    rows, subj = window_starts(g, subject_names, W, S)
    g = np.asarray(g).reshape(-1, 1)
    first = rows[:, 0]
    return (np.asarray(image)[rows], np.asarray(kin)[rows],
            g[first].astype(np.float32).reshape(-1, 1), np.asarray(e5)[first], subj)


def powerset_error_labels(e5: np.ndarray, delete_nd: bool = True):
    """5-column (OOV, ND, MA, NP, Error) -> 7-column (NoErr, OOV, MA, NP, OOV+MA, MA+NP, Error)
    int32 labels and the Needle-Drop-only mask.

    Branch order follows reference MED/dataset/dataset_utils.py:793-843 (the same chain is
    duplicated in MED/dataset/CustomFrameDataset.py:195-245): the first matching rule wins."""
    e5 = np.asarray(e5, dtype=np.float32)
    n = e5.shape[0]
    out = np.zeros((n, 7), dtype=np.int32)
    nd_mask = np.zeros(n, dtype=bool)
    for i in range(n):
        oov, nd, ma, npos, err = (e5[i, j] for j in range(5))
        if err == 1:
            out[i, 6] = 1
            single = (np.float32(oov) + np.float32(nd) + np.float32(ma) + np.float32(npos)) == 1
            if (oov == 1 and single) or (oov == 1 and nd == 1):
                out[i, 1] = 1
            elif (ma == 1 and single) or (ma == 1 and nd == 1):
                out[i, 2] = 1
            elif (npos == 1 and single) or (npos == 1 and oov == 1):
                out[i, 3] = 1
            elif oov == 1 and ma == 1:
                out[i, 4] = 1
            elif ma == 1 and npos == 1:
                out[i, 5] = 1
            elif nd == 1:
                if delete_nd:
                    out[i, 6] = 0
                    nd_mask[i] = True
            # else: unrecognised combination, row stays [0,0,0,0,0,0,1] (:837-838)
        else:
            out[i, 0] = 1
    return out, nd_mask


def class_balance(e7: np.ndarray):
    """``binary_error_distribution`` and ``specific_error_distribution`` of
    ``CustomWindowDataset`` (reference MED/dataset/CustomWindowDataset.py:42-46), computed the way
    torch does it there: int32 sums, true division by the row count, weights in float32."""
    import torch
    t = torch.from_numpy(np.asarray(e7, dtype=np.int32))
    n = len(t)
    p = t[:, -1].sum() / n
    binary = (float(1 - p), float(p))
    specific = (n / (t[:, :-1].sum(axis=0) + 1e-5)).tolist()
    return binary, specific


def window_predictions(predictions, e_labels, gestures, subjects, W=10, S=6, binary=True):
    """Frame-level predictions -> window-level predictions.

    Reference MED/modeling/modeling_utils.py:2695-2777: same walk as ``window_data`` but the
    subjects come from ``np.unique`` (SORTED order, :2722), the window value is the mean of the
    frame predictions, thresholded ``>= 0.5`` (binary, :2754) or ``np.round`` (half-to-even,
    :2758), labels are those of the first frame (:2760)."""
    predictions = np.asarray(predictions)
    e_labels = np.asarray(e_labels)
    gestures = np.asarray(gestures)
    subjects = np.asarray(subjects)
    pw, ew, gw, sw = [], [], [], []
    for s in np.unique(subjects):
        idx = np.flatnonzero(subjects == s)
        for p in window_starts_subject(gestures[idx], W, S):
            m = np.mean(predictions[idx[p:p + W]])
            pw.append((1.0 if m >= 0.5 else 0.0) if binary else np.round(m))
            ew.append(e_labels[idx[p]])
            gw.append(gestures[idx[p]])
            sw.append(s)
    return np.asarray(pw), np.asarray(ew), np.asarray(gw), sw


def soft_vote(p_a, p_b):
    """Ensemble soft vote of two window models: ``(p_a + p_b) / 2 >= 0.5``
    (reference ensemble.ipynb cell 6, source lines 10-19)."""
    return ((np.asarray(p_a, dtype=np.float64) + np.asarray(p_b, dtype=np.float64)) / 2 >= 0.5).astype(np.int64)


def cascade(binary_preds, multiclass_preds):
    """Ensemble cascade: multiclass prediction where the binary model fired, else 0
    (reference ensemble.ipynb cell 15, source lines 53-63)."""
    b = np.asarray(binary_preds)
    out = np.zeros_like(np.asarray(multiclass_preds))
    out[b == 1] = np.asarray(multiclass_preds)[b == 1]
    return out
