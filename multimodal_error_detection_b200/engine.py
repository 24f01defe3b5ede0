"""One window-classifier train step as a reusable object (what ``train_single_epoch`` does per batch,
reference MED/modeling/modeling_utils.py:335-366), with static device buffers so that the whole step
-- index lookup -> K1 gather -> K2 FeatureExtractor -> head -> K3 loss -> backward -> gradient
all-reduce -> fused Adam -- can be captured in a CUDA graph and replayed (a step is a few hundred short
kernels; launched one by one from Python the host, not the GPU, sets the pace).

The only per-step input is the batch's WINDOW INDICES (8 bytes per window, copied from pinned host
memory or staged on the device); frames never leave HBM.

Software pipelining of K1 (``prefetch=True``): the batch buffers are double-buffered and the gather of step k+1 is
issued on a side stream right after the FeatureExtractor forward of step k, so it runs under the LSTM recurrence of
step k (whose kernels hold a CTA on 64 of the 148 SMs only).  Protocol: ``prime(idx_0)`` once, then
``run(next_idx=idx_{k+1})`` per step; every step still performs exactly one gather.  With ``prefetch=False`` the gather
of step k runs at the start of step k on the main stream (this is also how bench.py times K1 for the roofline).
"""
from __future__ import annotations

import torch

from . import _lib
from .dataset.CustomWindowDataset import CustomWindowDataset
from .modeling import modeling_utils as mu


class WindowTrainStep:
    def __init__(self, dataset: CustomWindowDataset, feature_extractor, model, criterion, optimizer, exp_kwargs: dict,
                 batch_size: int, gather_variant: int = 0, prefetch: bool = False):
        self.ds, self.fe, self.model, self.crit, self.opt, self.kw = dataset, feature_extractor, model, criterion, optimizer, exp_kwargs
        self.B, self.W = batch_size, dataset.W
        dev = dataset._starts.device
        self.device = dev
        self.image_dtype = mu._image_dtype(feature_extractor)
        self.prefetch = prefetch
        self.prefetch_sms = int(exp_kwargs.get("prefetch_sms", 56))     # SMs the prefetching gather may occupy
        self.parity = 0                # which of the two batch buffers holds the CURRENT step's batch
        self._primed = False           # prefetch mode: buffer[parity] already holds the gathered batch
        self.idx2 = [torch.zeros(batch_size, dtype=torch.int64, device=dev) for _ in range(2)]     # the step's only input
        self.label_col = mu.define_error_labels(dataset.e_labels_data, exp_kwargs).float().contiguous()
        D_img, D_kin = dataset._image_table.shape[1], dataset._kin_table.shape[1]
        nbuf = 2 if prefetch else 1        # double-buffered batches only when the next step's gather runs inside this step
        imgs = [torch.empty(batch_size, self.W, D_img, dtype=self.image_dtype, device=dev) for _ in range(nbuf)]
        kins = [torch.empty(batch_size, self.W, D_kin, dtype=torch.float32, device=dev) for _ in range(nbuf)]
        self.images2 = [imgs[0], imgs[-1]]
        self.kin2 = [kins[0], kins[-1]]
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.probs = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.preds = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.labels = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.gather_variant = gather_variant
        # K1 fused into the FeatureExtractor's first layer (csrc/gather_gemm.cu): no materialised image batch on the way in
        from . import ops as _ops
        first = feature_extractor.linear[0] if feature_extractor is not None else None
        stat_rows = dataset._img_stats[0].shape[0] if dataset._img_stats is not None else 1
        self.fused = bool(exp_kwargs.get("fused_gather", True)) and not prefetch and feature_extractor is not None \
            and self.image_dtype == torch.bfloat16 and exp_kwargs["data_type"] == "multimodal" \
            and _ops.gather_linear_supported(dataset._image_table, self.W, first.out_features, stat_rows)
        self.gather_events = None      # optional (start, end) CUDA events around K1
        self.phase_events = None       # optional list of 9 (external) CUDA events at the phase boundaries of the step
        self.graphs = [None, None]     # one captured step per buffer parity
        self.launches_per_step = None
        self._side = torch.cuda.Stream(device=dev)
        optimizer.prepare()
        mu._setup_exchange(optimizer)      # several ranks: gradient exchange over NVLink peer memory (collective set-up)

    # the current batch buffers (what the last / next run() trains on)
    @property
    def idx(self):
        return self.idx2[self.parity]

    @property
    def images(self):
        return self.images2[self.parity]

    @property
    def kin(self):
        return self.kin2[self.parity]

    @property
    def graph(self):
        """Truthy when the step is replayed from captured graphs."""
        return self.graphs[0]

    def _gather(self, which: int, sm_cap: int = 0):
        from . import ops
        starts = ops.take_rows(self.ds._starts, self.idx2[which])
        if self.gather_events is not None:
            self.gather_events[0].record()
        self.ds.gather_batch(None, image_out=self.images2[which], kin_out=self.kin2[which], starts=starts,
                             exact=self.image_dtype == torch.float32, variant=self.gather_variant | (sm_cap << 8))
        if self.gather_events is not None:
            self.gather_events[1].record()

    def load(self, idx: torch.Tensor):
        """Stage the CURRENT step's window indices (pinned host -> device, or device -> device).  In prefetch mode this
        also gathers the batch right away (start of a sequence; same as :meth:`prime`)."""
        self.idx2[self.parity].copy_(idx, non_blocking=True)
        if self.prefetch:
            self._gather(self.parity)
            self._primed = True

    prime = load

    def _mark(self, k: int):
        if self.phase_events is not None:
            self.phase_events[k].record()

    def _body(self, cur: int):
        nxt = 1 - cur
        self._mark(0)
        from . import ops
        prepack = getattr(self.model, "prepack", None)
        if prepack is not None:        # parameter-only operand packing of the head: on the side stream, under the first kernels
            prepack(self.B, self.W)
        labels = ops.take_rows(self.label_col, self.idx2[cur], out=self.labels)
        parts = None
        if self.fused:
            # the image stream is gathered INSIDE the first FeatureExtractor layer; the kinematics (26 columns) either inside the
            # kernel that builds the LSTM's first operand (heads that take WindowParts) or through K1 + one concat kernel
            from .heads import concat_features
            starts = ops.take_rows(self.ds._starts, self.idx2[cur])
            km = self.ds._kin_stats
            by_parts = getattr(self.model, "accepts_parts", lambda: False)()
            if not by_parts:
                ops.gather_norm([ops.GatherStream(self.ds._kin_table, km[0] if km else None, km[1] if km else None, self.kin2[cur], 0, True)],
                                starts, self.W)
            self._mark(1)
            im = self.ds._img_stats
            # heads that build their own first operand round the features to bf16 there: they take them in bf16 right away
            feats = self.fe.forward_table(self.ds._image_table, im[0] if im else None, im[1] if im else None, starts, self.W,
                                          events=self.gather_events, out_bf16=by_parts)
            if by_parts:
                from .lstm_stack import WindowParts
                parts = WindowParts(self.ds._kin_table, km[0] if km else None, km[1] if km else None, starts)
                inputs = feats
            else:
                inputs = concat_features(feats, self.kin2[cur]).permute(0, 2, 1)
        else:
            if not self.prefetch:
                self._gather(cur)
            self._mark(1)
            inputs = mu.define_inputs(self.images2[cur], self.kin2[cur], self.fe, self.kw, self.device)
        self._mark(2)
        if self.phase_events is not None and inputs.requires_grad:
            inputs.register_hook(lambda g: self._mark(5))      # gradient w.r.t. the head input = end of the head's backward
        if self.prefetch:
            # K1 of the NEXT step: forked here so that it runs under the LSTM recurrence of this step
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._gather(nxt, self.prefetch_sms)      # leaves the other SMs to the main stream's kernels
        outputs = self.model(inputs) if parts is None else self.model(inputs, parts=parts)
        self._mark(3)
        static = isinstance(self.crit, mu.FusedBCEWithLogitsLoss)
        if static:      # K3 writes the step's persistent result buffers itself (no copies at the end of the step)
            self.crit.static_out = dict(loss=self.loss, probs=self.probs, preds=self.preds, counts=self.counts)
        try:
            loss, _ = mu.compute_loss(outputs, labels, self.crit, "window")
        finally:
            if static:
                self.crit.static_out = None
        self.opt.zero_grad()
        self._mark(4)
        mu._backward(loss, self.opt)
        self._mark(6)
        mu._allreduce_grads(self.opt)
        self.opt.step()
        self._mark(7)
        if not static:
            probs, preds, counts = self.crit.last
            self.loss.copy_(loss.detach().reshape(1))
            self.counts.copy_(counts)
            self.probs.copy_(probs)
            self.preds.copy_(preds)
        if self.prefetch:
            torch.cuda.current_stream().wait_stream(self._side)
        self._mark(8)

    def capture(self, warmup: int = 3):
        """Run `warmup` real steps on a side stream (the current contents of the index buffers are used), then record
        the step into one CUDA graph per buffer parity."""
        mu._set_train(self.model, self.fe, self.kw, True)
        # the warm-up steps are real optimiser steps: snapshot every piece of training state and put it back
        # afterwards, so that capturing the graphs leaves the training trajectory untouched
        self.opt._refresh_active()
        mods = [m for m in (self.fe, self.model) if m is not None]
        snap = [t.clone() for c in self.opt.chunks for t in (c.param, c.exp_avg, c.exp_avg_sq)] + [self.opt.state_dev.clone()]
        bufs = [b for m in mods for b in m.buffers()]
        snap_bufs = [b.clone() for b in bufs]
        keep = [t.clone() for t in (*self.idx2, *self.images2, *self.kin2)]
        self.idx2[1 - self.parity].copy_(self.idx2[self.parity])          # valid indices for the prefetch of the warm-up steps
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for k in range(warmup):
                self._body((self.parity + k) & 1 if self.prefetch else self.parity)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        graphs = [None, None]
        pars = (0, 1) if self.prefetch else (self.parity,)
        for par in pars:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body(par)
            graphs[par] = g
        self.launches_per_step = (_lib.launch_count() - n0) // len(pars)   # b200med kernels recorded in (and replayed by) a graph
        self.graphs = graphs
        with torch.no_grad():
            it = iter(snap)
            for c in self.opt.chunks:
                for t in (c.param, c.exp_avg, c.exp_avg_sq):
                    t.copy_(next(it))
            self.opt.state_dev.copy_(next(it))
            for b, sb in zip(bufs, snap_bufs):
                b.copy_(sb)
            for t, k in zip((*self.idx2, *self.images2, *self.kin2), keep):
                t.copy_(k)
        self.opt._lr_on_device = None      # the device-side lr was part of the restored state: push it again on the next run
        return self

    def run(self, next_idx: torch.Tensor = None):
        """One train step on the current batch.  Prefetch mode: ``next_idx`` = the window indices of the NEXT step (its
        gather runs inside this step); without it the next step has to :meth:`load` its batch itself."""
        if self.prefetch and not self._primed:
            raise RuntimeError("WindowTrainStep(prefetch=True): call load()/prime() with the first batch of a sequence")
        cur = self.parity
        if self.prefetch and next_idx is not None:
            self.idx2[1 - cur].copy_(next_idx, non_blocking=True)
        self.opt.sync_lr()             # lr lives in a device scalar; refresh it outside the graph when the scheduler moved it
        if self.graphs[cur] is not None:
            self.graphs[cur].replay()
        else:
            self._body(cur)
        if self.prefetch:
            self.parity = 1 - cur
            self._primed = next_idx is not None
        return self.loss


class _TrainState:
    """Snapshot / restore of everything a train step mutates (flat parameters, Adam moments and step scalars, module
    buffers): the warm-up steps that precede a graph capture are real optimiser steps and must leave no trace."""

    def __init__(self, optimizer, modules):
        optimizer._refresh_active()
        self.opt = optimizer
        self.bufs = [b for m in modules if m is not None for b in m.buffers()]
        self.snap = [t.clone() for c in optimizer.chunks for t in (c.param, c.exp_avg, c.exp_avg_sq)] + [optimizer.state_dev.clone()]
        self.snap_bufs = [b.clone() for b in self.bufs]

    def restore(self):
        with torch.no_grad():
            it = iter(self.snap)
            for c in self.opt.chunks:
                for t in (c.param, c.exp_avg, c.exp_avg_sq):
                    t.copy_(next(it))
            self.opt.state_dev.copy_(next(it))
            for b, sb in zip(self.bufs, self.snap_bufs):
                b.copy_(sb)
        self.opt._lr_on_device = None      # the device-side lr was part of the restored state: push it again on the next run


class FrameTrainStep:
    """One frame-classifier train step on ONE video (what ``train_single_epoch`` does per DataLoader(batch_size=1) item,
    reference MED/modeling/modeling_utils.py:335-366) captured in a CUDA graph.  The video's tensors are resident
    (CustomFrameDataset caches them on the device), so the graph reads them in place: a replay has NO input.  A step is
    ~100 short kernels (0.9 ms of device time under ncu) that take 2.3 ms when launched one by one from Python; every
    video of a fold keeps its own graph (its T is baked into the grids), all graphs share one memory pool -- they never
    run concurrently -- and are replayed once per epoch."""

    def __init__(self, images, kin, e7, feature_extractor, model, criterion, optimizer, exp_kwargs: dict, pool=None):
        self.images, self.kin, self.e7 = images, kin, e7
        self.fe, self.model, self.crit, self.opt, self.kw = feature_extractor, model, criterion, optimizer, exp_kwargs
        dev = images.device
        self.device = dev
        T = images.shape[1]
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.preds = torch.zeros(T, dtype=torch.float32, device=dev)
        self.labels = mu.define_error_labels(e7, exp_kwargs).float().contiguous()
        self.graph, self.pool, self.launches_per_step = None, pool, None
        optimizer.prepare()
        if getattr(model, "graph_seed", False) is None:
            model.graph_seed = (int(torch.randint(0, 2 ** 62, (1,)).item()), torch.zeros(1, dtype=torch.int64, device=dev))

    def _body(self, exchange: bool = True):
        inputs = mu.define_inputs(self.images, self.kin, self.fe, self.kw, self.device)
        outputs = self.model(inputs)
        loss, _ = mu.compute_loss(outputs, self.labels, self.crit, "frame")
        self.opt.zero_grad()
        mu._backward(loss, self.opt, exchange=exchange)
        if exchange:
            mu._allreduce_grads(self.opt)
        self.opt.step()
        preds, counts = mu.compute_loss.last_frame
        self.loss.copy_(loss.detach().reshape(1))
        self.counts.copy_(counts)
        self.preds.copy_(preds.reshape(-1))

    def capture(self, warmup: int = 2):
        mu._set_train(self.model, self.fe, self.kw, True)
        state = _TrainState(self.opt, (self.fe, self.model))
        seed_counter = None if getattr(self.model, "graph_seed", None) is None else self.model.graph_seed[1].clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        # The warm-up steps run WITHOUT the gradient all-reduce (they are rolled back anyway).  Under data parallelism the
        # ranks capture different videos at different times -- the shuffled order deals the videos anew every epoch, so in
        # epoch 2 one rank may capture a video while its peer replays a cached graph -- and a real collective inside a
        # capture-time warm-up would pair with the peer's NEXT step (measured on two GPUs: replicas drifted apart).  With
        # collective-free captures every rank issues exactly one all-reduce per step, the one inside the replayed graph.
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body(exchange=False)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.pool):
            self._body()
        self.launches_per_step = _lib.launch_count() - n0
        self.graph = g
        state.restore()
        if seed_counter is not None:
            self.model.graph_seed[1].copy_(seed_counter)
        return self

    def run(self):
        self.opt.sync_lr()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss
