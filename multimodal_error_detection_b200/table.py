"""Device-resident per-frame feature table and the window index over it.

Data layout in HBM: one row per frame, subjects (trials) contiguous:
    image  [N, 2048] f32   (8 KB rows, 16-byte aligned -> 128-bit vector / TMA bulk friendly)
    kin    [N, 26]   f32
    g      [N]       f32   gesture id
    e5     [N, 5]    f32   (OOV, ND, MA, NP, Error)
    offsets[n_subjects+1] i64   first row of every subject
A window is (start row, W): its frames are W consecutive rows, so nothing is duplicated -- the
reference materialises every window (W/S-fold copy of the table, dataset_utils.py:243-244).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


def cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200med needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def factorize_subjects(names: Sequence) -> tuple:
    """Subject codes in order of FIRST APPEARANCE (pandas ``.unique()`` order used by window_data,
    dataset_utils.py:193) -> (codes [N] int64 numpy, unique names list)."""
    names = np.asarray(names, dtype=object)
    uniq, first, inv = np.unique(names.astype(str), return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")            # sorted-unique position -> appearance rank
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return rank[inv].astype(np.int64), [str(uniq[i]) for i in order]


class FrameTable:
    """Flat per-frame table on the GPU with contiguous subjects."""

    def __init__(self, image: torch.Tensor, kin: torch.Tensor, g: torch.Tensor, e5: torch.Tensor,
                 subject_names: Sequence, device: Optional[torch.device] = None, extra_streams=None):
        device = device or cuda_device()
        codes, self.subjects = factorize_subjects(subject_names)
        n = len(codes)
        order = None
        if n and np.any(np.diff(codes) < 0):
            # Subjects interleaved in the input: the reference walks each subject's own row list
            # (dataset_utils.py:194, 207), which equals a stable regrouping by subject.
            order = np.argsort(codes, kind="stable")
            codes = codes[order]
        counts = np.bincount(codes, minlength=len(self.subjects)) if n else np.zeros(0, dtype=np.int64)
        self.offsets_host = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.row_order = order  # table row -> original row (None = identity)

        def put(t, dtype=torch.float32):
            t = torch.as_tensor(t)
            if order is not None:
                t = t[torch.from_numpy(order).to(t.device)]
            return t.to(device=device, dtype=dtype).contiguous()

        self.image = put(image)
        self.kin = put(kin)
        self.g = put(g).reshape(-1)
        self.e5 = put(e5).reshape(-1, 5)
        self.extra = [put(x) for x in (extra_streams or [])]
        self.offsets = torch.from_numpy(self.offsets_host).to(device)
        self.device = device

    @property
    def n_frames(self) -> int:
        return self.g.numel()

    def window_index(self, W: int, S: int) -> "WindowIndex":
        r = ops.window_index(self.g, self.offsets, W, S, self.e5)
        return WindowIndex(self, W, S, r["starts"], r["g_win"], r["e5_win"], r["subj_win"])


class WindowIndex:
    """Start rows + first-frame labels of every window of a FrameTable (K0 output)."""

    def __init__(self, table: FrameTable, W: int, S: int, starts, g_win, e5_win, subj_win):
        self.table, self.W, self.S = table, W, S
        self.starts, self.g_win, self.e5_win, self.subj_win = starts, g_win, e5_win, subj_win

    def __len__(self):
        return self.starts.numel()

    def subject_names(self) -> List[str]:
        idx = self.subj_win.cpu().numpy()
        return [self.table.subjects[i] for i in idx]

    def select(self, keep: torch.Tensor) -> "WindowIndex":
        """Boolean-mask the windows (Needle-Drop deletion, dataset_utils.py:442-453)."""
        return WindowIndex(self.table, self.W, self.S, self.starts[keep].contiguous(), self.g_win[keep].contiguous(),
                           self.e5_win[keep].contiguous(), self.subj_win[keep].contiguous())
