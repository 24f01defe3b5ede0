"""Drop-in for the reference ``MED/dataset/CustomFrameDataset.py`` (one whole video per item).

Same constructor, ``__len__``, ``__getitem__`` 6-tuple ``(images, kinematics, g_labels, e_labels,
subject, skill_level)``, ``get_n_frames`` and ``powerset_error_labels`` as the reference class
(CustomFrameDataset.py:12-247).  The reference unpickles the trial and runs a T-iteration Python
label loop on EVERY ``__getitem__`` of every epoch (:49-56, :84, :195-245); here each trial is read
once, its label transform / Needle-Drop deletion / kinematics standardisation run once on the device
(K0 powerset kernel, row-standardise kernel), and ``__getitem__`` returns views of resident tensors.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset

from .. import ops
from ..table import cuda_device
from .dataset_utils import _trial_arrays, load_feature_standardization

SKILL_OF_SUBJECT = {"B": 0, "G": 0, "H": 0,      # novice        (reference CustomFrameDataset.py:26-34)
                    "C": 1, "F": 1,              # intermediate
                    "D": 2, "E": 2, "I": 2}      # expert


class CustomFrameDataset(Dataset):

    def __init__(self, fold_data_path: str, video_data_path: str = None, error_type: str = "global",
                 csv_filename: str = "trains.csv", delete_ND: bool = True):
        self.fold_data_path, self.video_data_path = fold_data_path, video_data_path
        self.error_type, self.delete_ND = error_type, delete_ND
        self.feature_standardization_dict = load_feature_standardization(fold_data_path)
        self.csv_file = pd.read_csv(os.path.join(fold_data_path, csv_filename), header=None, names=["files"])
        self.skill_level_dict = {"B": "Novice", "C": "Intermediate", "D": "Expert", "E": "Expert", "F": "Intermediate",
                                 "G": "Novice", "H": "Novice", "I": "Expert"}
        self._cache = {}

    def __len__(self):
        return len(self.csv_file)

    def _load(self, idx: int):
        dev = cuda_device()
        file_name = self.csv_file["files"].iloc[idx]
        image, kin, g, e5 = _trial_arrays(self.fold_data_path, file_name, self.video_data_path)
        image, kin, g, e5 = (t.to(dev).contiguous() for t in (image, kin, g, e5))
        # always computed with delete_ND=True; the mask is APPLIED only if self.delete_ND (:84-90)
        e7, nd_mask = ops.powerset(e5, True)
        if self.delete_ND:
            keep = ~nd_mask
            image, kin, g, e7 = image[keep].contiguous(), kin[keep].contiguous(), g[keep].contiguous(), e7[keep].contiguous()
        st = self.feature_standardization_dict.get("kinematics")
        if st is not None and kin.shape[0]:   # images are NOT standardised on the frame path (:93-95)
            D = kin.shape[1]
            kin = ops.standardise_rows(kin, ops.expand_stat(st["mean"], D, 1, dev), ops.expand_stat(st["std"], D, 1, dev))
        subject = file_name[:-4]
        letter = subject[-4]
        if letter not in SKILL_OF_SUBJECT:
            raise KeyError(letter)
        skill = torch.zeros((kin.size(0), 3), device=dev)
        skill[:, SKILL_OF_SUBJECT[letter]] = 1
        return image, kin, g, e7, subject, skill

    def __getitem__(self, idx):
        idx = int(idx)
        if idx not in self._cache:
            self._cache[idx] = self._load(idx)
        return self._cache[idx]

    def load_feature_standardization_dict(self):
        return load_feature_standardization(self.fold_data_path)

    def get_n_frames(self):
        """Total frames over the fold's trials (reference CustomFrameDataset.py:130-160)."""
        total = 0
        for file in self.csv_file["files"]:
            if not file.endswith(".pkl"):
                raise ValueError(f"File {file} does not end with .pkl. Please check the CSV file.")
            path = os.path.join(self.fold_data_path, file)
            if not os.path.exists(path):
                raise FileNotFoundError(f"File {path} does not exist. Please check the path.")
            with open(path, "rb") as f:
                data = pickle.load(f)
            if "image_feats" in data:
                total += data["image_feats"].shape[0]
            elif "feature" in data:
                total += data["feature"].shape[0]
            else:
                raise KeyError(f"Neither 'image_feats' nor 'feature' found in {file}. Please check the data format.")
        return total

    def powerset_error_labels(self, e_labels_data: torch.Tensor, delete_ND: bool = True) -> tuple:
        from .dataset_utils import powerset_error_labels
        return powerset_error_labels(e_labels_data, delete_ND)


class FrameLoader:
    """``DataLoader(frame_dataset, batch_size=1, shuffle=..., generator=...)`` (reference train_frame.ipynb
    cell 2, lines 58-66): same shuffled video order (a real torch sampler over the indices), batches are the
    device-resident trial tensors with a leading batch dimension of 1."""

    def __init__(self, dataset: CustomFrameDataset, shuffle: bool = False, generator=None, rank: int = 0, world_size: int = 1,
                 even_shards: bool = None):
        from torch.utils.data import DataLoader
        from .CustomWindowDataset import _IndexOnly
        self.dataset, self.rank, self.world_size = dataset, rank, world_size
        # Data-parallel TRAINING all-reduces the gradients once per video, so every rank must take the same number of steps:
        # the shuffled order is padded with its own head up to a multiple of world_size (default for shuffle=True).
        # Validation has no collective per video and shards the videos exactly (default for shuffle=False).
        self.even_shards = shuffle if even_shards is None else bool(even_shards)
        self._order = DataLoader(_IndexOnly(len(dataset)), batch_size=1, shuffle=shuffle, generator=generator,
                                 collate_fn=lambda items: int(items[0]))

    def __len__(self):
        return len(self._order)

    def indices(self):
        """This rank's video indices in the order of one pass (advances the sampler's generator like a pass does)."""
        if self.world_size <= 1:
            yield from self._order
            return
        order = list(self._order)
        if self.even_shards and len(order) % self.world_size:
            order += order[:self.world_size - len(order) % self.world_size]
        yield from order[self.rank::self.world_size]

    def __iter__(self):
        for i in self.indices():
            images, kin, g, e7, subject, skill = self.dataset[i]
            yield images.unsqueeze(0), kin.unsqueeze(0), g.unsqueeze(0), e7.unsqueeze(0), (subject,), skill.unsqueeze(0)
