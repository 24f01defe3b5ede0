"""Drop-in for the reference ``MED/dataset/CustomWindowDataset.py`` backed by a device-resident
frame table and the fused gather/standardise kernel (K1).

Same constructor, ``__len__``, ``__getitem__`` 5-tuple and the two class-balance attributes as the
reference class (CustomWindowDataset.py:22-74).  Two ways to build it:

* the reference way -- materialised windows ``image_data [n, W, 2048]`` etc.; the windows are then
  treated as a table of ``n*W`` rows with ``starts = arange(n) * W``;
* :meth:`from_index` -- a :class:`~multimodal_error_detection_b200.table.WindowIndex` over the flat
  table (what ``retrieve_dataloaders_window`` uses): no window is ever materialised on the host.

Batches are produced by :meth:`gather_batch` -- one K1 launch per batch instead of B Python
``__getitem__`` calls + ``default_collate`` + an H2D copy (modeling_utils.py:40-42).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .. import ops
from ..parallel import shard_batch
from ..table import WindowIndex, cuda_device


class CustomWindowDataset(Dataset):

    def __init__(self, image_data, kinematics_data, g_labels_data, e_labels_data, subject_data,
                 feature_standardization_dict={}):
        dev = cuda_device()
        image_data = torch.as_tensor(image_data)
        kinematics_data = torch.as_tensor(kinematics_data)
        n, W = image_data.shape[0], image_data.shape[1]
        self.W = W
        self._image_table = image_data.to(dev, torch.float32).reshape(n * W, -1).contiguous()
        self._kin_table = kinematics_data.to(dev, torch.float32).reshape(n * W, -1).contiguous()
        self._starts = (torch.arange(n, dtype=torch.int32, device=dev) * W).contiguous()
        self.g_labels_data = torch.as_tensor(g_labels_data).to(dev)
        self.e_labels_data = torch.as_tensor(e_labels_data).to(dev)
        self.subject_data = subject_data
        self._subjects = list(subject_data["subject"]) if hasattr(subject_data, "__getitem__") and "subject" in subject_data else list(subject_data)
        self._finish(feature_standardization_dict)

    @classmethod
    def from_index(cls, index: WindowIndex, e7: torch.Tensor, feature_standardization_dict={}, subjects: Optional[List[str]] = None):
        import pandas as pd
        self = cls.__new__(cls)
        self.W = index.W
        self._image_table, self._kin_table = index.table.image, index.table.kin
        self._starts = index.starts
        self.g_labels_data = index.g_win.reshape(-1, 1)
        self.e_labels_data = e7
        self._subjects = subjects if subjects is not None else index.subject_names()
        self.subject_data = pd.DataFrame(self._subjects, columns=["subject"])
        self.index = index
        self._finish(feature_standardization_dict)
        return self

    def _finish(self, stats):
        self.feature_standardization_dict = stats
        dev = self._image_table.device
        D_img, D_kin = self._image_table.shape[1], self._kin_table.shape[1]
        self._img_stats = self._kin_stats = None
        for key, value in (stats or {}).items():
            if key == "image":
                self._img_stats = (ops.expand_stat(value["mean"], D_img, self.W, dev), ops.expand_stat(value["std"], D_img, self.W, dev))
            elif key == "kinematics":
                self._kin_stats = (ops.expand_stat(value["mean"], D_kin, self.W, dev), ops.expand_stat(value["std"], D_kin, self.W, dev))
        # class balance, computed with the reference's own torch expressions (CustomWindowDataset.py:42-46)
        e = self.e_labels_data.cpu()
        self._e_host = e
        n = len(e)
        self.binary_error_distribution = (1 - e[:, -1].sum() / n, e[:, -1].sum() / n)
        self.specific_error_distribution = (n / (e[:, :-1].sum(axis=0) + 1e-5)).tolist()

    # -- reference-compatible views (materialise on demand, on the device) -------------------------
    @property
    def image_data(self):
        return self.gather_batch(None, standardise=False)[0]

    @property
    def kinematics_data(self):
        return self.gather_batch(None, standardise=False)[1]

    def __len__(self):
        return self._starts.numel()

    def subjects_of(self, idx) -> List[str]:
        return [self._subjects[int(i)] for i in idx]

    def gather_batch(self, idx: Optional[torch.Tensor], standardise: bool = True, image_dtype=torch.float32,
                     image_out: Optional[torch.Tensor] = None, kin_out: Optional[torch.Tensor] = None,
                     kin_col: int = 0, exact: bool = True, variant: int = 0, starts: Optional[torch.Tensor] = None):
        """K1: (images [B, W, D_img], kinematics [B, W, D_kin]) for window indices ``idx`` (CUDA int64/int32
        tensor, or None = all windows).  ``kin_out``/``kin_col`` let the kinematics land directly in the
        concat buffer the head consumes."""
        if starts is None:
            starts = self._starts if idx is None else self._starts.index_select(0, idx.to(self._starts.device, torch.long))
        B, W = starts.numel(), self.W
        dev = self._image_table.device
        D_img, D_kin = self._image_table.shape[1], self._kin_table.shape[1]
        if image_out is None:
            image_out = torch.empty(B, W, D_img, dtype=image_dtype, device=dev)
        if kin_out is None:
            kin_out = torch.empty(B, W, D_kin, dtype=torch.float32, device=dev)
        im = self._img_stats if standardise else None
        km = self._kin_stats if standardise else None
        streams = [ops.GatherStream(self._image_table, im[0] if im else None, im[1] if im else None, image_out, 0, exact),
                   ops.GatherStream(self._kin_table, km[0] if km else None, km[1] if km else None, kin_out, kin_col, True)]
        if B:
            ops.gather_norm(streams, starts.contiguous(), W, variant)
        return image_out, kin_out

    def __getitem__(self, idx):
        i = torch.tensor([int(idx)], device=self._starts.device)
        image, kin = self.gather_batch(i)
        return image[0], kin[0], self.g_labels_data[idx], self.e_labels_data[idx], self._subjects[int(idx)]


class _IndexOnly(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class DeviceWindowLoader:
    """Stands in for ``torch.utils.data.DataLoader`` over a CustomWindowDataset
    (reference dataset_utils.py:526-527).  Batch COMPOSITION comes from a real torch DataLoader over an
    index-only dataset driven by the same ``Generator().manual_seed(42)``, so the shuffled order is
    the reference's bit for bit (SURVEY section 7 "sampler parity"); batch CONTENT comes from K1.

    Iterating yields the reference's 5-tuples ``(images, kinematics, g_labels, e_labels, subject)`` as
    CUDA tensors; the train / validation loops use :meth:`index_batches` and fuse the gather into the
    step instead.  With ``world_size > 1`` every rank walks the same global permutation and takes its
    contiguous share of each global batch (data-parallel sharding)."""

    def __init__(self, dataset: CustomWindowDataset, batch_size: int, shuffle: bool = False, generator=None,
                 rank: int = 0, world_size: int = 1):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, batch_size, shuffle, generator
        self.rank, self.world_size = rank, world_size
        self.max_batches = None        # optional cap on batches per pass (bounded benchmark epochs)
        # the real thing, kept as the definition of "the reference's order" (tests compare index_batches() against it)
        self._order = DataLoader(_IndexOnly(len(dataset)), batch_size=batch_size, shuffle=shuffle, generator=generator,
                                 collate_fn=lambda items: torch.as_tensor(items, dtype=torch.int64))

    def __len__(self):
        return len(self._order)

    def _global_batches(self):
        """The batches torch's DataLoader would produce, without its per-sample Python loop (8192 `__getitem__` calls and a
        list round trip per batch cost more host time than the GPU needs for the step).  Mirrors the generator consumption
        of torch.utils.data exactly: `_BaseDataLoaderIter.__init__` draws a base seed (`random_` on an int64 scalar), then
        RandomSampler draws `randperm(n)` and -- once the pass is exhausted -- a second, discarded `randperm(n)`
        (`num_samples % n` tail).  tests/test_host_logic.py pins this against the real DataLoader."""
        n, B = len(self.dataset), self.batch_size
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for lo in range(0, n, B):
            yield order[lo:lo + B]
        if self.shuffle:
            torch.randperm(n, generator=self.generator)

    def index_batches(self):
        """Host int64 index tensors, one per (rank-local) batch (the consumer pins them: torch's caching host allocator
        keeps a pinned block alive until the asynchronous upload that reads it has run)."""
        for idx, _ in self.index_batches_with_global():
            yield idx

    def index_batches_with_global(self):
        """(rank-local index tensor, size of the GLOBAL batch it was cut from).  The second value lets a data-parallel train
        loop weight its gradient by n_local / n_global when the shards of a short last batch differ by one window."""
        for k, idx in enumerate(self._global_batches()):
            if self.max_batches is not None and k >= self.max_batches:
                break
            n_global = idx.numel()
            if self.world_size > 1:
                if n_global < self.world_size:
                    continue       # fewer windows than ranks: dropped on EVERY rank (a rank without a step would stall the all-reduce)
                idx = shard_batch(idx, self.rank, self.world_size)
            yield idx, n_global

    def dp_weight(self, n_local: int, n_global: int) -> float:
        """Factor that turns this rank's mean-reduced loss into its share of the GLOBAL batch mean under a SUM all-reduce
        scaled by 1 / world_size: n_local * world_size / n_global (exactly 1 for even shards)."""
        if self.world_size <= 1 or n_global <= 0:
            return 1.0
        return float(n_local) * self.world_size / float(n_global)

    def __iter__(self):
        ds = self.dataset
        for idx in self.index_batches():
            didx = idx.to(ds._starts.device, non_blocking=True)
            images, kin = ds.gather_batch(didx)
            yield (images, kin, ds.g_labels_data.index_select(0, didx), ds.e_labels_data.index_select(0, didx),
                   tuple(ds.subjects_of(idx.tolist())))
