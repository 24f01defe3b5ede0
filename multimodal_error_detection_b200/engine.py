"""One window-classifier train step as a reusable object (what ``train_single_epoch`` does per batch,
reference MED/modeling/modeling_utils.py:335-366), with static device buffers so that the whole step
-- index lookup -> K1 gather -> K2 FeatureExtractor -> head -> K3 loss -> backward -> gradient
all-reduce -> fused Adam -- can be captured in a CUDA graph and replayed (a step is a few hundred short
kernels; launched one by one from Python the host, not the GPU, sets the pace).

The only per-step input is the batch's WINDOW INDICES (8 bytes per window, copied from pinned host
memory or staged on the device); frames never leave HBM.
"""
from __future__ import annotations

import torch

from . import _lib
from .dataset.CustomWindowDataset import CustomWindowDataset
from .modeling import modeling_utils as mu


class WindowTrainStep:
    def __init__(self, dataset: CustomWindowDataset, feature_extractor, model, criterion, optimizer, exp_kwargs: dict,
                 batch_size: int, gather_variant: int = 0):
        self.ds, self.fe, self.model, self.crit, self.opt, self.kw = dataset, feature_extractor, model, criterion, optimizer, exp_kwargs
        self.B, self.W = batch_size, dataset.W
        dev = dataset._starts.device
        self.device = dev
        self.image_dtype = mu._image_dtype(feature_extractor)
        self.idx = torch.zeros(batch_size, dtype=torch.int64, device=dev)          # the step's only input
        self.label_col = mu.define_error_labels(dataset.e_labels_data, exp_kwargs).float().contiguous()
        D_img, D_kin = dataset._image_table.shape[1], dataset._kin_table.shape[1]
        self.images = torch.empty(batch_size, self.W, D_img, dtype=self.image_dtype, device=dev)
        self.kin = torch.empty(batch_size, self.W, D_kin, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.probs = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.preds = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.labels = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.gather_variant = gather_variant
        self.gather_events = None      # optional (start, end) CUDA events around K1 (eager mode only)
        self.graph = None
        self.launches_per_step = None
        optimizer.prepare()

    def load(self, idx: torch.Tensor):
        """Stage one batch of window indices (pinned host -> device, or device -> device)."""
        self.idx.copy_(idx, non_blocking=True)

    def _body(self):
        starts = self.ds._starts.index_select(0, self.idx)
        labels = self.label_col.index_select(0, self.idx)
        if self.gather_events is not None:
            self.gather_events[0].record()
        self.ds.gather_batch(None, image_out=self.images, kin_out=self.kin, starts=starts,
                             exact=self.image_dtype == torch.float32, variant=self.gather_variant)
        if self.gather_events is not None:
            self.gather_events[1].record()
        inputs = mu.define_inputs(self.images, self.kin, self.fe, self.kw, self.device)
        outputs = self.model(inputs)
        loss, _ = mu.compute_loss(outputs, labels, self.crit, "window")
        self.opt.zero_grad()
        mu._backward(loss, self.opt)
        mu._allreduce_grads(self.opt)
        self.opt.step()
        probs, preds, counts = self.crit.last
        self.loss.copy_(loss.detach().reshape(1))
        self.counts.copy_(counts)
        self.probs.copy_(probs)
        self.preds.copy_(preds)
        self.labels.copy_(labels)

    def capture(self, warmup: int = 3):
        """Run `warmup` real steps on a side stream (the current contents of ``idx`` are used), then record the
        step into a CUDA graph."""
        mu._set_train(self.model, self.fe, self.kw, True)
        # the warm-up steps are real optimiser steps: snapshot every piece of training state and put it back
        # afterwards, so that capturing the graph leaves the training trajectory untouched
        self.opt._refresh_active()
        mods = [m for m in (self.fe, self.model) if m is not None]
        snap = [t.clone() for c in self.opt.chunks for t in (c.param, c.exp_avg, c.exp_avg_sq)] + [self.opt.state_dev.clone()]
        bufs = [b for m in mods for b in m.buffers()]
        snap_bufs = [b.clone() for b in bufs]
        first_step = self.opt._active is None or not any(self.opt._active.values())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._body()
        self.launches_per_step = _lib.launch_count() - n0     # b200med kernels recorded in (and replayed by) the graph
        self.graph = graph
        with torch.no_grad():
            it = iter(snap)
            for c in self.opt.chunks:
                for t in (c.param, c.exp_avg, c.exp_avg_sq):
                    t.copy_(next(it))
            self.opt.state_dev.copy_(next(it))
            for b, sb in zip(bufs, snap_bufs):
                b.copy_(sb)
        self.opt._lr_on_device = None      # the device-side lr was part of the restored state: push it again on the next run
        return self

    def run(self):
        self.opt.sync_lr()             # lr lives in a device scalar; refresh it outside the graph when the scheduler moved it
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss
