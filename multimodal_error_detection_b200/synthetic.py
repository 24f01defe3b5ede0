"""Seeded synthetic JIGSAWS-like folds.

The reference ships no data (its ``/data`` is git-ignored), so every parity test, the
golden-vector generator and ``bench.py`` draw their per-frame tables from here.  Only numpy's
``PCG64`` is used, so a (seed, shape) pair gives the same bytes in the build container and on
the GPU box.

Layout mirrors what ``load_data`` consumes (reference ``MED/dataset/dataset_utils.py:36-157``):
one trial = image_feats [T,2048] f32, kinematics_feats [T,26] f32, g_labels [T] int,
e_labels [T,5] f32 with columns (Out_Of_View, Needle_Drop, Multiple_Attempts, Needle_Position,
Error).
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass, field

import numpy as np

IMAGE_DIM = 2048
KIN_DIM = 26
SUBJECT_LETTERS = "BCDEFGHI"  # reference CustomFrameDataset.py:26-34


@dataclass
class Trial:
    name: str
    image: np.ndarray      # [T, image_dim] float32
    kin: np.ndarray        # [T, 26] float32
    g: np.ndarray          # [T] int64
    e5: np.ndarray         # [T, 5] float32


@dataclass
class Fold:
    train: list = field(default_factory=list)
    test: list = field(default_factory=list)
    mean_image: np.ndarray = None
    std_image: np.ndarray = None
    mean_kin: np.ndarray = None
    std_kin: np.ndarray = None


def gesture_track(rng: np.random.Generator, T: int, lead_zeros: int = 3,
                  run_lo: int = 8, run_hi: int = 60, zero_run_prob: float = 0.05) -> np.ndarray:
    """Piecewise-constant gesture ids; the first ``lead_zeros`` frames are 0 so the
    "first non-zero gesture" rule (dataset_utils.py:211-212) is exercised; a few later
    runs are 0 too (Appendix A-3: later zero runs are *not* excluded)."""
    g = np.zeros(T, dtype=np.int64)
    t = min(lead_zeros, T)
    prev = 0
    while t < T:
        run = int(rng.integers(run_lo, run_hi + 1))
        lab = int(rng.integers(1, 10))
        if rng.random() < zero_run_prob:
            lab = 0
        elif lab == prev:
            lab = lab % 9 + 1
        g[t:t + run] = lab
        prev = lab
        t += run
    return g


def error_track(rng: np.random.Generator, g: np.ndarray, p_error: float = 0.5,
                p_rare: float = 0.15) -> np.ndarray:
    """Per-gesture-run error labels.  85 % of error runs carry exactly one of OOV/MA/NP;
    the rest walk through every other branch of the powerset rule (two-label combos,
    ND-only, ND+x, and an "unrecognised" all-zero-but-Error row)."""
    T = len(g)
    e = np.zeros((T, 5), dtype=np.float32)
    bounds = np.flatnonzero(np.diff(g)) + 1
    starts = np.concatenate([[0], bounds])
    ends = np.concatenate([bounds, [T]])
    rare = [(1, 0, 0, 1), (1, 0, 1, 0), (0, 0, 1, 1), (0, 1, 0, 0), (1, 1, 0, 0),
            (0, 1, 1, 0), (0, 1, 0, 1), (1, 1, 1, 1), (0, 0, 0, 0)]
    for s, t in zip(starts, ends):
        if rng.random() >= p_error:
            continue
        if rng.random() < p_rare:
            combo = rare[int(rng.integers(0, len(rare)))]
        else:
            k = (0, 2, 3)[int(rng.integers(0, 3))]
            combo = tuple(1 if j == k else 0 for j in range(4))
        e[s:t, :4] = np.asarray(combo, dtype=np.float32)
        e[s:t, 4] = 1.0
    return e


def make_trial(rng: np.random.Generator, name: str, T: int, image_dim: int = IMAGE_DIM,
               p_error: float = 0.5) -> Trial:
    g = gesture_track(rng, T)
    e5 = error_track(rng, g, p_error=p_error)
    # post-ReLU pooled ResNet features are non-negative
    image = np.maximum(rng.standard_normal((T, image_dim), dtype=np.float32), 0.0)
    kin = rng.standard_normal((T, KIN_DIM), dtype=np.float32)
    # weak class signal so that a few epochs of training move the metrics
    image[:, :8] += e5[:, 4:5] * 0.75
    kin[:, :4] += e5[:, 4:5] * 0.5
    return Trial(name=name, image=image, kin=kin, g=g, e5=e5)


def trial_name(i: int) -> str:
    task = "Suturing" if i % 2 == 0 else "Needle_Passing"
    letter = SUBJECT_LETTERS[i % len(SUBJECT_LETTERS)]
    return f"{task}_{letter}{(i // len(SUBJECT_LETTERS)) + 1:03d}"


def make_fold(seed: int = 42, n_train: int = 6, n_test: int = 2, t_lo: int = 300, t_hi: int = 900,
              image_dim: int = IMAGE_DIM, p_error: float = 0.5) -> Fold:
    rng = np.random.Generator(np.random.PCG64(seed))
    fold = Fold()
    for i in range(n_train + n_test):
        T = int(rng.integers(t_lo, t_hi + 1))
        tr = make_trial(rng, trial_name(i), T, image_dim=image_dim, p_error=p_error)
        (fold.train if i < n_train else fold.test).append(tr)
    img = np.concatenate([t.image for t in fold.train])
    kin = np.concatenate([t.kin for t in fold.train])
    fold.mean_image = img.mean(0).astype(np.float32)
    fold.std_image = (img.std(0) + 1e-3).astype(np.float32)
    fold.mean_kin = kin.mean(0).astype(np.float32)
    fold.std_kin = (kin.std(0) + 1e-3).astype(np.float32)
    return fold


def write_fold(fold: Fold, path: str) -> str:
    """Write the on-disk layout the reference loaders read:
    ``<trial>.pkl`` + ``train.csv``/``test.csv`` + four ``*.pth`` statistics files."""
    import torch
    os.makedirs(path, exist_ok=True)
    for split, trials in (("train", fold.train), ("test", fold.test)):
        with open(os.path.join(path, f"{split}.csv"), "w") as f:
            for tr in trials:
                f.write(f"{tr.name}.pkl\n")
        for tr in trials:
            blob = {
                "image_feats": torch.from_numpy(tr.image.copy()),
                "kinematics_feats": torch.from_numpy(tr.kin.copy()),
                "g_labels": tr.g.copy(),
                "e_labels": torch.from_numpy(tr.e5.copy()),
                "frames": np.arange(len(tr.g)),
            }
            with open(os.path.join(path, f"{tr.name}.pkl"), "wb") as f:
                pickle.dump(blob, f)
    torch.save(torch.from_numpy(fold.mean_image.copy()), os.path.join(path, "mean_features.pth"))
    torch.save(torch.from_numpy(fold.std_image.copy()), os.path.join(path, "std_features.pth"))
    torch.save(torch.from_numpy(fold.mean_kin.copy()), os.path.join(path, "mean_kinematics.pth"))
    torch.save(torch.from_numpy(fold.std_kin.copy()), os.path.join(path, "std_kinematics.pth"))
    return path


def write_fold_video_schema(fold: Fold, path: str, video_path: str) -> tuple:
    """The reference's SECOND on-disk schema (dataset_utils.py:73-96, used with ``video_data_path``): image features
    live in ``<video_path>/<trial>.pkl`` as a numpy ``'feature'`` array [T, 2048]; kinematics / labels stay in the fold's
    own ``<trial>.pkl`` (whose image_feats are then ignored).  Returns (fold path, video path)."""
    write_fold(fold, path)
    os.makedirs(video_path, exist_ok=True)
    for tr in fold.train + fold.test:
        # a DIFFERENT image stream than the fold's own pickle, so that a loader reading the wrong file is caught
        with open(os.path.join(video_path, f"{tr.name}.pkl"), "wb") as f:
            pickle.dump({"feature": (tr.image[:, ::-1] * 0.5 + 0.25).astype(np.float32).copy()}, f)
    return path, video_path


def flat_tables(trials):
    """Concatenate trials into the flat per-frame arrays ``load_data`` returns, plus the
    per-frame subject names and the contiguous subject offsets."""
    image = np.concatenate([t.image for t in trials])
    kin = np.concatenate([t.kin for t in trials])
    g = np.concatenate([t.g for t in trials]).astype(np.float32).reshape(-1, 1)
    e5 = np.concatenate([t.e5 for t in trials])
    names = np.concatenate([[t.name] * len(t.g) for t in trials])
    offsets = np.concatenate([[0], np.cumsum([len(t.g) for t in trials])]).astype(np.int64)
    return image, kin, g, e5, names, offsets


def label_tracks(seed: int, n_videos: int, t_lo: int = 300, t_hi: int = 900, p_error: float = 0.5):
    """Labels only (no features) for the large throughput tables, whose feature streams are
    drawn on the device by bench.py.  Returns g [N] f32, e5 [N,5] f32, offsets [V+1] i64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gs, es, lens = [], [], []
    for _ in range(n_videos):
        T = int(rng.integers(t_lo, t_hi + 1))
        g = gesture_track(rng, T)
        gs.append(g)
        es.append(error_track(rng, g, p_error=p_error))
        lens.append(T)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return np.concatenate(gs).astype(np.float32), np.concatenate(es), offsets
