"""Ensemble inference (BASELINE config 5): frame model + window model over a
device-resident frame table, fused on the device.

The reference fuses STORED per-sample outputs in ``ensemble.ipynb`` (cell 6: soft vote ``(p_a + p_b) / 2 >= 0.5``; cell 15:
cascade) and bridges frame -> window predictions with ``window_predictions`` (MED/modeling/modeling_utils.py:2695-2777).
Here the two forwards run too:

* frame model (TeCNo): videos are concatenated along time and run in ragged passes (``MultiStageModel.forward_ragged``:
  taps never cross a video) instead of one ``DataLoader(batch_size=1)`` step per video;
* window model: K1 gather -> FeatureExtractor -> head per batch of window indices, probabilities from the K3 kernel;
* bridge + fusion: window vote of the frame predictions over the SAME window index (K0), soft vote, confusion counts.

Everything is sharded by video (a window never crosses a video): each rank owns a contiguous range of videos, the only
exchange is the sum of the confusion counts.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .table import FrameTable, WindowIndex


def video_passes(lengths, frames_per_pass: int):
    """Group consecutive videos into ragged passes of at most ``frames_per_pass`` frames (a video longer than the budget
    is a pass of its own): list of (first video, one past the last video)."""
    passes, v, n_videos = [], 0, len(lengths)
    while v < n_videos:
        w, n = v, 0
        while w < n_videos and (w == v or n + int(lengths[w]) <= frames_per_pass):
            n += int(lengths[w]); w += 1
        passes.append((v, w))
        v = w
    return passes


def frame_features(feature_extractor, image_table: torch.Tensor, r0: int, r1: int) -> torch.Tensor:
    """FeatureExtractor output [r1 - r0, out] f32 for the rows [r0, r1) of the resident fp32 frame table (inference).  bf16 mode:
    consecutive rows are "windows" of 128 frames for the fused gather + first-layer kernel (csrc/gather_gemm.cu), so the table
    rows go fp32 -> bf16 operand tile -> tcgen05.mma inside ONE kernel -- no separate cast pass (8 KB read + 4 KB written per
    frame) in front of the GEMM and no bf16 copy read back by it; a ragged tail is one more, overlapping, block."""
    n = r1 - r0
    first = feature_extractor.linear[0]
    if (getattr(feature_extractor, "precision", "fp32") != "bf16" or n < 128 or torch.is_grad_enabled()
            or not ops.gather_linear_supported(image_table, 128, first.out_features, 1)):
        return feature_extractor(image_table[r0:r1]).float()
    nb, rem = divmod(n, 128)
    starts = torch.arange(r0, r0 + nb * 128, 128, dtype=torch.int32, device=image_table.device)
    if rem:
        starts = torch.cat([starts, torch.tensor([r1 - 128], dtype=torch.int32, device=image_table.device)])
    feats = feature_extractor.forward_table(image_table, None, None, starts, 128).float()          # [blocks, 128, out]
    out = feats[:nb].reshape(nb * 128, -1)
    return out if not rem else torch.cat([out, feats[nb, 128 - rem:]], dim=0)


@torch.no_grad()
def frame_model_predictions(table: FrameTable, feature_extractor, model, exp_kwargs: dict, kin_stats: Optional[dict] = None,
                            frames_per_pass: int = 1 << 17) -> torch.Tensor:
    """Frame-level predictions [N] f32 (argmax of the LAST stage, modeling_utils.py:370 / :751) for every frame of the
    table.  Frame-path inputs: raw image features through the FeatureExtractor, kinematics standardised
    (CustomFrameDataset.py:93-95), concatenated on the feature axis (modeling_utils.py:41-42)."""
    model.eval()
    if feature_extractor is not None:
        feature_extractor.eval()
    off = table.offsets_host
    lengths = np.diff(off)
    N, dev = table.n_frames, table.device
    preds = torch.empty(N, dtype=torch.float32, device=dev)
    dt = exp_kwargs["data_type"]
    # ragged geometry of the WHOLE table, once, on the device (a pass takes slices: no host work, no copy per pass)
    lens_dev = torch.from_numpy(lengths.astype(np.int64)).to(dev)
    tloc = (torch.arange(N, device=dev) - torch.repeat_interleave(table.offsets[:-1], lens_dev, output_size=N)).to(torch.int32)
    trem = (torch.repeat_interleave(lens_dev, lens_dev, output_size=N) - 1 - tloc).to(torch.int32)
    for v, w in video_passes(lengths, frames_per_pass):
        r0, r1 = int(off[v]), int(off[w])
        cols = []
        if dt in ("multimodal", "video"):
            cols.append(table.image[r0:r1] if (dt == "video" and exp_kwargs["video_dims"] == 2048)
                        else frame_features(feature_extractor, table.image, r0, r1))
        if dt in ("multimodal", "kinematics"):
            kin = table.kin[r0:r1]
            if kin_stats is not None:
                D = kin.shape[1]
                kin = ops.standardise_rows(kin.contiguous(), ops.expand_stat(kin_stats["mean"], D, 1, dev),
                                           ops.expand_stat(kin_stats["std"], D, 1, dev))
            cols.append(kin)
        frames = cols[0] if len(cols) == 1 else torch.cat(cols, dim=1)
        logits = model(frames.contiguous().float().unsqueeze(0).permute(0, 2, 1), (tloc[r0:r1], trem[r0:r1]))   # [stages, 1, C, n]
        preds[r0:r1] = torch.argmax(logits[-1, 0], dim=0).float()
    return preds


@torch.no_grad()
def window_model_probabilities(dataset, feature_extractor, model, exp_kwargs: dict, batch_size: int = 18944) -> torch.Tensor:
    """sigmoid(logit) [n] f32 of the binary window model for every window of the dataset (validate_single_epoch's forward,
    modeling_utils.py:735-752, without the per-sample host loop).  Default batch = 148 SMs x 128 windows: the persistent LSTM
    recurrence kernels own 128 windows per CTA, so this is the batch that puts one CTA on every SM."""
    from .modeling import modeling_utils as mu
    model.eval()
    if feature_extractor is not None:
        feature_extractor.eval()
    n, dev = len(dataset), dataset._starts.device
    probs = torch.empty(n, dtype=torch.float32, device=dev)
    image_dtype = mu._image_dtype(feature_extractor)
    zeros = torch.zeros(batch_size, dtype=torch.float32, device=dev)
    # bf16 mode, window lengths the fused kernel serves, a head that takes the kinematics from the table: the gather runs inside
    # the first FeatureExtractor layer (no bf16 batch is written in inference) and inside the LSTM's first operand (no concat)
    first = feature_extractor.linear[0] if feature_extractor is not None else None
    stat_rows = dataset._img_stats[0].shape[0] if dataset._img_stats is not None else 1
    fused = (feature_extractor is not None and image_dtype == torch.bfloat16 and exp_kwargs["data_type"] == "multimodal"
             and getattr(model, "accepts_parts", lambda: False)()
             and ops.gather_linear_supported(dataset._image_table, dataset.W, first.out_features, stat_rows))
    for lo in range(0, n, batch_size):
        hi = min(n, lo + batch_size)
        if fused:
            from .lstm_stack import WindowParts
            starts = dataset._starts[lo:hi].contiguous()
            im, km = dataset._img_stats, dataset._kin_stats
            feats = feature_extractor.forward_table(dataset._image_table, im[0] if im else None, im[1] if im else None, starts, dataset.W)
            out = model(feats, parts=WindowParts(dataset._kin_table, km[0] if km else None, km[1] if km else None, starts))
        else:
            idx = torch.arange(lo, hi, device=dev)
            images, kin = dataset.gather_batch(idx, image_dtype=image_dtype, exact=image_dtype == torch.float32)
            out = model(mu.define_inputs(images, kin, feature_extractor, exp_kwargs, dev))
        r = ops.bce_logits(out.reshape(-1).float().contiguous(), zeros[: hi - lo], want_grad=False, want_probs=True)
        probs[lo:hi] = r["probs"]
    return probs


@torch.no_grad()
def fuse(frame_preds: torch.Tensor, index: WindowIndex, window_probs: torch.Tensor, window_labels: Optional[torch.Tensor] = None):
    """Bridge + fusion on the device: window value of the frame predictions (mean over the window ``>= 0.5``,
    modeling_utils.py:2752-2754) on the window model's own index, then the soft vote of ensemble.ipynb cell 6 with the
    confusion counts (tn, fp, fn, tp) against ``window_labels``."""
    frame_windows = ops.window_vote(frame_preds, index.starts, index.W, True)
    fused, counts = ops.soft_vote(window_probs, frame_windows, window_labels)
    return dict(frame_windows=frame_windows, fused=fused, counts=counts)


def ensemble_inference(table: FrameTable, dataset, frame_objects, window_objects, kin_stats=None, batch_size: int = 18944,
                       frames_per_pass: int = 1 << 17):
    """frame_objects / window_objects = (feature_extractor, model, exp_kwargs); ``dataset`` = the window dataset built
    over ``table`` (its ``index`` is the window index after Needle-Drop deletion).  Returns the dict of :func:`fuse` plus
    the two models' raw outputs; with torch.distributed initialised the counts are summed over the ranks."""
    from . import parallel
    f_fe, f_model, f_kw = frame_objects
    w_fe, w_model, w_kw = window_objects
    frame_preds = frame_model_predictions(table, f_fe, f_model, f_kw, kin_stats, frames_per_pass)
    window_probs = window_model_probabilities(dataset, w_fe, w_model, w_kw, batch_size)
    from .modeling import modeling_utils as mu
    labels = mu.define_error_labels(dataset.e_labels_data, w_kw).float().contiguous()
    out = fuse(frame_preds, dataset.index, window_probs, labels)
    out["counts"] = parallel.allreduce_sum_(out["counts"])
    out.update(frame_preds=frame_preds, window_probs=window_probs, labels=labels)
    return out
