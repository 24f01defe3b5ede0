"""Drop-in for the hot-path functions of the reference ``MED/dataset/dataset_utils.py``.

Same names and argument meaning as the reference (file:line cited per function); the work happens in
the K0 / K1 CUDA kernels over a device-resident frame table.  Siamese pair construction
(dataset_utils.py:282-353, 534-757) is out of scope.
"""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np
import pandas as pd
import torch

from .. import ops
from ..table import FrameTable, WindowIndex, cuda_device
from .CustomWindowDataset import CustomWindowDataset, DeviceWindowLoader


def _trial_arrays(fold_data_path: str, pkl_file: str, video_data_path: Optional[str]):
    """One trial's per-frame arrays from either on-disk schema (dataset_utils.py:73-96 / :118-136)."""
    with open(os.path.join(fold_data_path, pkl_file), "rb") as f:
        data = pickle.load(f)
    if video_data_path is not None:
        with open(os.path.join(video_data_path, pkl_file), "rb") as f:
            vid = pickle.load(f)
        n = vid["feature"].shape[0]
        image = torch.as_tensor(np.asarray(vid["feature"]).reshape(n, 2048), dtype=torch.float32)
    else:
        n = data["image_feats"].shape[0]
        image = torch.as_tensor(data["image_feats"]).reshape(n, 2048).to(torch.float32)
    kin = torch.as_tensor(data["kinematics_feats"]).to(torch.float32)
    g = torch.as_tensor(np.asarray(data["g_labels"]).reshape(n, 1)).to(torch.float32)
    e = torch.as_tensor(np.asarray(data["e_labels"]) if not torch.is_tensor(data["e_labels"]) else data["e_labels"]
                        ).reshape(n, 5).to(torch.float32)
    return image, kin[:n], g, e


def load_data(fold_data_path: str, csv_filename: str, video_data_path: str = None) -> tuple:
    """Flat per-frame host tensors of every trial listed in ``csv_filename``
    (reference dataset_utils.py:36-157): (image [N,2048], kinematics [N,26], g [N,1], e [N,5],
    subject DataFrame[N])."""
    csv_file = pd.read_csv(os.path.join(fold_data_path, csv_filename), header=None, names=["files"])
    images, kins, gs, es, subjects = [], [], [], [], []
    for pkl_file in csv_file["files"]:
        if not pkl_file.endswith(".pkl"):
            continue
        image, kin, g, e = _trial_arrays(fold_data_path, pkl_file, video_data_path)
        images.append(image); kins.append(kin); gs.append(g); es.append(e)
        subjects += [pkl_file[:-4]] * image.shape[0]
    cat = lambda xs, w: torch.cat(xs) if xs else torch.empty((0, w))
    return (cat(images, 2048), cat(kins, 26), cat(gs, 1), cat(es, 5), pd.DataFrame({"subject": subjects}))


def window_index(image_data, kinematics_data, g_labels_data, e_labels_data, subject_data,
                 window_size: int = 10, stride: int = 6) -> WindowIndex:
    """Native form of ``window_data``: upload the flat table once and return the window index (start
    rows + first-frame labels) without materialising any window."""
    names = subject_data["subject"].tolist() if isinstance(subject_data, pd.DataFrame) else list(subject_data)
    table = FrameTable(image_data, kinematics_data, g_labels_data, e_labels_data, names)
    return table.window_index(window_size, stride)


def window_data(image_data, kinematics_data, g_labels_data, e_labels_data, subject_data, window_size=10, stride=6):
    """Reference dataset_utils.py:161-258: same 5-tuple -- (image_windows [n,W,D], kinematics_windows
    [n,W,26], g_labels_windows [n,1], e_labels_windows [n,5], subject_windows DataFrame[n]) -- with the
    window walk done by the K0 kernel and the windows materialised ON THE DEVICE by the K1 gather
    (no standardisation).  Prefer :func:`window_index` + :class:`CustomWindowDataset.from_index`, which
    never materialise."""
    idx = window_index(image_data, kinematics_data, g_labels_data, e_labels_data, subject_data, window_size, stride)
    n, W = len(idx), window_size
    t = idx.table
    image = torch.empty(n, W, t.image.shape[1], dtype=t.image.dtype, device=t.device)
    kin = torch.empty(n, W, t.kin.shape[1], dtype=t.kin.dtype, device=t.device)
    if n:
        ops.gather_norm([ops.GatherStream(t.image, out=image), ops.GatherStream(t.kin, out=kin)], idx.starts, W)
    subject_windows = pd.DataFrame(idx.subject_names(), columns=["subject"])
    return image, kin, idx.g_win.reshape(-1, 1), idx.e5_win, subject_windows


def compute_window_size_stride(frequency: int = 30) -> tuple:
    """2 s windows, 1.33 s stride (reference dataset_utils.py:262-279)."""
    return int(2 * frequency), int(4 / 3 * frequency)


def powerset_error_labels(e_labels_data: torch.Tensor, delete_ND: bool = True) -> tuple:
    """Reference dataset_utils.py:760-845: (int32 [n,7] powerset labels, bool [n] Needle-Drop mask).
    Computed by the K0 powerset kernel; results are returned on the input's device."""
    src = torch.as_tensor(e_labels_data)
    e7, mask = ops.powerset(src.to(cuda_device(), torch.float32).contiguous(), delete_ND)
    return e7.to(src.device), mask.to(src.device)


def load_and_window(fold_data_path: str, window_size: int = 30, stride: int = 20, video_data_path: str = None):
    """Reference dataset_utils.py:357-402, returning window INDICES (train, test) instead of
    materialised windows."""
    out = []
    for csv in ("train.csv", "test.csv"):
        flat = load_data(fold_data_path, csv, video_data_path=video_data_path)
        out.append(window_index(*flat, window_size=window_size, stride=stride))
    return tuple(out)


def load_feature_standardization(fold_data_path: str) -> dict:
    """mean/std files of a fold (reference dataset_utils.py:457-464)."""
    ld = lambda n: torch.load(os.path.join(fold_data_path, n))
    return {"image": {"mean": ld("mean_features.pth"), "std": ld("std_features.pth")},
            "kinematics": {"mean": ld("mean_kinematics.pth"), "std": ld("std_kinematics.pth")}}


def dataset_from_index(index: WindowIndex, delete_ND: bool, stats: dict) -> CustomWindowDataset:
    """powerset labels -> optional Needle-Drop window deletion -> dataset (dataset_utils.py:433-453, 508-524)."""
    e7, nd_mask = ops.powerset(index.e5_win, delete_ND)
    if delete_ND:
        keep = ~nd_mask
        index, e7 = index.select(keep), e7[keep].contiguous()
    return CustomWindowDataset.from_index(index, e7, stats)


def retrieve_dataloaders_window(fold_data_path: str, exp_kwargs: dict, window_size: int = 30, stride: int = 20,
                                video_data_path: str = None, rank: int = 0, world_size: int = 1):
    """Reference dataset_utils.py:405-531: (train_dataloader, test_dataloader); ``.dataset`` exposes
    ``binary_error_distribution`` / ``specific_error_distribution``.  ``rank`` / ``world_size`` shard
    every global batch across data-parallel ranks (absent in the single-process reference)."""
    if exp_kwargs.get("siamese"):
        raise ValueError("Siamese datasets are outside the b200med hot path (SURVEY.md section 2, row 8).")
    train_idx, test_idx = load_and_window(fold_data_path, window_size, stride, video_data_path)
    stats = load_feature_standardization(fold_data_path)
    train_dataset = dataset_from_index(train_idx, exp_kwargs["delete_ND"], stats)
    test_dataset = dataset_from_index(test_idx, exp_kwargs["delete_ND"], stats)
    train_dataloader = DeviceWindowLoader(train_dataset, exp_kwargs["batch_size"], shuffle=True,
                                          generator=torch.Generator().manual_seed(42), rank=rank, world_size=world_size)
    test_dataloader = DeviceWindowLoader(test_dataset, exp_kwargs["batch_size"], shuffle=False,
                                         generator=torch.Generator().manual_seed(42), rank=rank, world_size=world_size)
    print(f"Number of training windows: {len(train_dataset)}")
    print(f"Number of testing windows: {len(test_dataset)}")
    return train_dataloader, test_dataloader
