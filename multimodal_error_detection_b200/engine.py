"""One window-classifier train step as a reusable object (what ``train_single_epoch`` does per batch,
reference MED/modeling/modeling_utils.py:335-366), with static device buffers so that the whole step
-- K1 gather -> K2 FeatureExtractor -> head -> K3 loss -> backward -> gradient all-reduce -> fused
Adam -- can be captured in a CUDA graph and replayed (a B=512 step is ~100 us of device work, so
launch latency would dominate otherwise; SURVEY.md section 7 "step is tiny").
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .dataset.CustomWindowDataset import CustomWindowDataset
from .modeling import modeling_utils as mu


class WindowTrainStep:
    def __init__(self, dataset: CustomWindowDataset, feature_extractor, model, criterion, optimizer, exp_kwargs: dict,
                 batch_size: int, use_graph: bool = False, gather_variant: int = 0):
        self.ds, self.fe, self.model, self.crit, self.opt, self.kw = dataset, feature_extractor, model, criterion, optimizer, exp_kwargs
        self.B, self.W = batch_size, dataset.W
        dev = dataset._starts.device
        self.device = dev
        self.image_dtype = mu._image_dtype(feature_extractor)
        self.starts = torch.zeros(batch_size, dtype=torch.int32, device=dev)       # static inputs of the step
        self.labels = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        D_img, D_kin = dataset._image_table.shape[1], dataset._kin_table.shape[1]
        self.images = torch.empty(batch_size, self.W, D_img, dtype=self.image_dtype, device=dev)
        self.kin = torch.empty(batch_size, self.W, D_kin, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.gather_variant = gather_variant
        self.gather_events = None      # optional (start, end) CUDA events around K1
        self.graph = None
        self.use_graph = use_graph
        mu._set_train(model, feature_extractor, exp_kwargs, True)
        optimizer.prepare()

    def load(self, starts: torch.Tensor, labels: torch.Tensor):
        """Stage one batch's inputs (device -> device, or pinned host -> device)."""
        self.starts.copy_(starts, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)

    def _body(self):
        if self.gather_events is not None:
            self.gather_events[0].record()
        self.ds.gather_batch(None, image_out=self.images, kin_out=self.kin, starts=self.starts,
                             exact=self.image_dtype == torch.float32, variant=self.gather_variant)
        if self.gather_events is not None:
            self.gather_events[1].record()
        inputs = mu.define_inputs(self.images, self.kin, self.fe, self.kw, self.device)
        outputs = self.model(inputs)
        loss, _ = mu.compute_loss(outputs, self.labels, self.crit, "window")
        self.opt.zero_grad()
        loss.backward()
        mu._allreduce_grads(self.opt)
        self.opt.step()
        self.loss.copy_(loss.detach().reshape(1))
        self.counts.copy_(self.crit.last[2])

    def capture(self, warmup: int = 3):
        """Warm up on a side stream, then capture the step into a CUDA graph."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        return self

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss
