#!/usr/bin/env python
"""bench.py -- train windows/sec of the sliding-window error classifier (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this implementation (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): window 16 / stride 4 over a synthetic
per-frame table (2048-d image stream + 26-d kinematics), FeatureExtractor 2048-512-256-32 + 3-layer
LSTM(128) head, BCE loss, Adam; B = 8192 windows per GPU per step (weak scaling).  A "step" is one full
train step: K1 gather/standardise -> K2 FE forward -> head -> K3 loss/metrics -> backward -> gradient
all-reduce -> fused Adam.  `value` times K steps on the device (CUDA events, barrier + synchronize on
both sides, max over ranks) with the step's window indices already resident in HBM; `e2e` runs the
public ``train_single_epoch`` over the same number of steps with pinned-host index batches copied in
and the loss read back every step.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

W, S = 16, 4
IMAGE_DIM, KIN_DIM = 2048, 26
METRIC, UNIT = "train_windows_per_sec", "windows/s"


def exp_kwargs(batch, precision):
    return dict(dataset_type="window", error_type="global", pos_weight=True, n_epochs=15, batch_size=batch, lr=1e-3,
                lr_scheduler=True, weight_decay=1e-4, num_layers=3, hidden_size=128, video_dims=32, data_type="multimodal",
                delete_ND=True, return_train_preds=False, siamese=False, model_name="SimpleLSTM", precision=precision)


def workload_name(batch, videos):
    return (f"train_window T: W={W} S={S} streams[{IMAGE_DIM},{KIN_DIM}] FE(2048-512-256-32)+LSTM(58,{W},3,128,1) "
            f"BCE+Adam, B={batch}/GPU, table {videos} videos/GPU U[300,900] frames")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in-process every 2 ms
    (a timed region of K 3-ms steps is over before a second `nvidia-smi` process has started), `nvidia-smi` as fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.costs = []
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self._sample_nvml()                      # priming queries: only the COST of the second one is kept (the first pays
            t0 = time.perf_counter()                 # NVML's lazy set-up: ~1.4 ms on a one-GPU box, which pushed the first
            self._sample_nvml()                      # in-region sample past the end of a 34 ms region; the GPU is idle right now)
            self.rows.clear()
            self.first_cost = time.perf_counter() - t0
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
        self.rows.append([sm, self.sm_max] + flags)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [x.strip() for x in out.split(",")]
        if len(c) >= 6 and c[0].replace(".", "").isdigit():
            self.rows.append([float(c[0]), float(c[1])] + [v.lower().startswith("active") for v in c[2:6]])

    def run(self):
        # Sparse on purpose: an NVML query is an RM call that takes ~0.1 ms on a one-GPU box but was measured at ~9 ms per
        # query on a two-GPU box, where sampling every 2 ms stretched the timed steps from 2.8 to 6.5 ms.  The interval backs
        # off to 25x the cost of the last query, so that sampling stays below ~4 % of the region whatever a query costs.
        base = float(os.environ.get("B200MED_CLOCK_INTERVAL_MS", "10")) * 1e-3
        if self.nvml is not None and self._stop_evt.wait(max(base, 25.0 * getattr(self, "first_cost", 0.0))):
            return
        while not self._stop_evt.is_set():
            t0 = time.perf_counter()
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            cost = time.perf_counter() - t0
            self.costs.append(cost)
            self._stop_evt.wait(max(base if self.nvml is not None else 0.1, 25.0 * cost))

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:                            # region shorter than the (backed-off) interval: one sample at its very end
            try:
                self._sample_nvml() if self.nvml is not None else self._sample_smi()
            except Exception:
                pass
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:6]) if v})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi",
                "query_ms": 1e3 * statistics.median(self.costs) if self.costs else None}


def k1_traffic(precision, batch, window, variant, fused=False):
    path = os.path.join(ROOT, "profiles", "k1_fused_traffic.json" if fused else "k1_traffic.json")
    try:
        rec = json.load(open(path))
        c = rec["config"]
        if (c["precision"], c["batch"], c["window"], c["gather_variant"]) == (precision, batch, window, variant):
            return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), f"profiles/{os.path.basename(path)} ({rec['source']})"
    except Exception:
        pass
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("hbm_gbs", 6650.0), p.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ======================================================================================== CPU arms
def oracle_dataset(n_windows, seed):
    """A bounded sample of the workload for the CPU arm, in the reference's representation:
    MATERIALISED windows [n, W, 2048] on the host (dataset_utils.py:243-244)."""
    from multimodal_error_detection_b200 import synthetic
    from oracle import loops, window_index as O
    rng = np.random.Generator(np.random.PCG64(seed))
    n_videos = n_windows // 60 + 4
    g, e5, offsets = synthetic.label_tracks(seed, n_videos, 300, 900)
    N = len(g)
    image = np.maximum(rng.standard_normal((N, IMAGE_DIM), dtype=np.float32), 0)
    kin = rng.standard_normal((N, KIN_DIM), dtype=np.float32)
    names = np.repeat(np.arange(n_videos), np.diff(offsets))
    t0 = time.perf_counter()
    rows, subj = O.window_starts(g, names, W, S)
    rows = rows[:n_windows]
    build_s = time.perf_counter() - t0
    e7, _ = O.powerset_error_labels(e5[rows[:, 0]], True)
    stats = {"image": {"mean": torch.from_numpy(image.mean(0)), "std": torch.from_numpy(image.std(0) + 1e-3)},
             "kinematics": {"mean": torch.from_numpy(kin.mean(0)), "std": torch.from_numpy(kin.std(0) + 1e-3)}}
    ds = loops.OracleWindowDataset(torch.from_numpy(image[rows]), torch.from_numpy(kin[rows]),
                                   torch.from_numpy(g[rows[:, 0]].reshape(-1, 1)), torch.from_numpy(e7), subj[:len(rows)], stats)
    return ds, build_s


def cpu_train_windows_per_sec(n_windows, batch, steps=None, warmup=0):
    """The reference's train_single_epoch cost structure (oracle port) on the host cores:
    windows/s = windows processed / wall time, data loading and sklearn metrics included.  The sample holds `n_windows`
    materialised windows (the reference's representation); `steps` train steps of `batch` windows are timed as repeated
    passes of the oracle's epoch loop over that sample (default: one pass)."""
    from oracle import loops, nets
    kw = exp_kwargs(batch, "fp32")
    ds, build_s = oracle_dataset(n_windows, seed=7)
    fe, model, crit, opt, sched = nets.build_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26},
                                                     ds.binary_error_distribution, W)
    loader, _ = loops.make_loaders(ds, ds, batch)
    per_pass = len(loader)
    steps = per_pass if steps is None else max(1, int(steps))
    if warmup:
        it = iter(loader)
        model.train(); fe.train()
        for _ in range(min(warmup, per_pass)):
            images, kin, g, e7, subj = next(it)
            out = model(loops.fuse_inputs(images, kin, fe, kw))
            loss, _ = loops.loss_fn(out, loops.select_labels(e7, kw).float(), crit, "window")
            opt.zero_grad(); loss.backward(); opt.step()
    passes = (steps + per_pass - 1) // per_pass
    t0 = time.perf_counter()
    for _ in range(passes):
        loops.train_epoch(model, fe, loader, crit, opt, None, kw)
    dt = time.perf_counter() - t0
    done = passes * len(ds)
    return done / dt, dt, done, build_s, passes * per_pass


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arms may use all the host threads there are."""
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    return torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port -- the reference is pure Python,
    nothing compiles into oracle/_ref) on all host threads, SAME config as the GPU arm: B = args.batch windows per step.
    Each step is a bounded sample: the oracle's epoch loop over 2 * B materialised windows (2 GB at B = 8192), repeated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    batch = args.batch
    steps = max(1, args.steps)
    wps, dt, n_done, build_s, steps_done = cpu_train_windows_per_sec(2 * batch, batch, steps=steps, warmup=min(args.warmup, 2))
    line = {"metric": METRIC, "value": wps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done, "warmup": min(args.warmup, 2),
            "ms_per_step": 1e3 * dt / steps_done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.batch, args.videos),
                       "note": "reference CPU path (oracle port of train_single_epoch, torch CPU fp32 + sklearn); each step is "
                               f"one batch of {batch} windows of the same workload drawn from a bounded sample of {2 * batch} "
                               "materialised windows"},
            "cpu_baseline": {"value": wps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_done} windows, B={batch}, {steps_done} steps over a {2 * batch}-window sample, "
                                       f"window build {build_s:.2f}s excluded"},
            "e2e": {"value": wps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ======================================================================================== GPU arm
def build_gpu_job(args, rank, device):
    from multimodal_error_detection_b200 import synthetic
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import CustomWindowDataset
    from multimodal_error_detection_b200.dataset.dataset_utils import dataset_from_index
    from multimodal_error_detection_b200.table import FrameTable
    g, e5, offsets = synthetic.label_tracks(1000 + rank, args.videos, 300, 900)
    N = len(g)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    image = torch.empty(N, IMAGE_DIM, device=device)
    for lo in range(0, N, 1 << 16):                       # chunked: no 2x transient of the 10 GB table
        image[lo:lo + (1 << 16)] = torch.randn(min(1 << 16, N - lo), IMAGE_DIM, device=device, generator=gen).clamp_min_(0)
    kin = torch.randn(N, KIN_DIM, device=device, generator=gen)
    names = np.repeat(np.arange(args.videos), np.diff(offsets))
    table = FrameTable(image, kin, torch.from_numpy(g), torch.from_numpy(e5), names, device=device)
    index = table.window_index(W, S)
    stats = {"image": {"mean": image[: 1 << 16].mean(0).cpu(), "std": (image[: 1 << 16].std(0) + 1e-3).cpu()},
             "kinematics": {"mean": kin.mean(0).cpu(), "std": (kin.std(0) + 1e-3).cpu()}}
    ds = dataset_from_index(index, True, stats)
    return ds, N


def timed_steps(ds, kw, device, batch, steps, warmup=3, graph=True, prefetch=False):
    """A bounded device-timed measurement of the window train step on `ds` for another configuration of the same workload
    (fp32 parity mode, strong-scaling batch): fresh model objects, K graph replays (or eager steps), CUDA events,
    max over ranks -> (ms per step, launch note, b200med launches per step)."""
    from multimodal_error_detection_b200 import _lib, parallel
    from multimodal_error_detection_b200.engine import WindowTrainStep
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, device,
                                                          ds.binary_error_distribution, W)
    n = len(ds)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(43))
    need = (steps + warmup + 1) * batch
    idx_all = perm.repeat((need + n - 1) // n)[:need].reshape(-1, batch).to(device)
    stepper = WindowTrainStep(ds, fe, model, crit, opt, kw, batch, prefetch=prefetch)
    note = "eager"
    mu._set_train(model, fe, kw, True)
    if graph:
        try:
            stepper.load(idx_all[0])
            stepper.capture()
            note = "cuda_graph"
        except Exception as e:
            stepper.graphs = [None, None]
            note = f"eager (graph capture failed: {type(e).__name__}: {e})"

    def step(i):
        if stepper.prefetch:
            if not stepper._primed:
                stepper.load(idx_all[i])
            stepper.run(idx_all[i + 1])
        else:
            stepper.load(idx_all[i])
            stepper.run()

    for i in range(warmup):
        step(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); parallel.barrier()
    n0 = _lib.launch_count()
    a.record()
    for i in range(steps):
        step(warmup + i)
    b.record()
    torch.cuda.synchronize(); parallel.barrier()
    ms = parallel.max_over_ranks(a.elapsed_time(b) / steps, device)
    launches = stepper.launches_per_step if stepper.graph is not None else (_lib.launch_count() - n0) // steps
    stepper.graphs = [None, None]
    return ms, note, launches


def other_configs(args, ds, device):
    """Auxiliary, bounded measurements of the BASELINE configs the headline line does not cover (they are parity-test
    cases, not bench lines), N = 1: the fp32 parity mode of the SAME train step (1e-5 arithmetic: six-product split-bf16 tcgen05 GEMMs for
    the large products, fp32 FMA kernels and an exact-math recurrence for the rest), config 0 (train_frame: FE + TeCNo, one video per step, eager / CUDA graph / stock torch layers) and config 5
    (ensemble inference: frame model + window model + device-side fusion) on ONE rank's shard of the 100 k-video job
    (12 500 videos, ~62 GB table).  Failures are reported, never fatal."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    out = {}
    try:
        Bf = 2048
        ms, note, launches = timed_steps(ds, exp_kwargs(Bf, "fp32"), device, Bf, steps=5, warmup=3)
        flops = Bf * W * (5_029_888 + 3 * 2 * 714_752)        # FE train + LSTM fwd/bwd per (window, step), SURVEY section 8d
        out["train_window_fp32"] = {"value": Bf / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "batch": Bf, "launch": note,
                                    "gpu_launches_per_step": launches, "dtype": "f32",
                                    "bound": "large products: tcgen05 on exactly split operands (x = h + m + l in bf16, six products, fp32 partial "
                                             "sums every 12 k-blocks: closer to fp64 than the fp32 FMA chain); per-step recurrent products and "
                                             "cells: fp32 FMA issue",
                                    "achieved_tflops": flops / (ms * 1e-3) / 1e12,
                                    "note": "same workload and step as the headline, exp_kwargs['precision'] = 'fp32'"}
    except Exception as e:
        out["train_window_fp32"] = {"error": f"{type(e).__name__}: {e}"}
    try:
        import bench_frame
        r = bench_frame.measure(frames=600, videos=32, steps=20)
        out["train_frame"] = {"unit": "frames/s", "frames_per_video": 600,
                              "cuda_graph": r["b200_graph"]["train_frames_per_s"], "eager": r["b200"]["train_frames_per_s"],
                              "cuda_graph_bf16_mode": r.get("b200_graph_bf16_fe", {}).get("train_frames_per_s"),
                              "ms_per_video_cuda_graph_bf16_mode": r.get("b200_graph_bf16_fe", {}).get("train_ms_per_video"),
                              "stock_torch_layers": r["torch_layers"]["train_frames_per_s"],
                              "ms_per_video_cuda_graph": r["b200_graph"]["train_ms_per_video"],
                              "inference_ragged_frames_per_s": r["head_inference"]["ragged_frames_per_s"], "config": r["config"]}
    except Exception as e:
        out["train_frame"] = {"error": f"{type(e).__name__}: {e}"}
    return out


ENSEMBLE_KEYS = ("value", "unit", "frames_per_s", "n_gpus", "videos_per_gpu", "frames_per_gpu", "ms_total", "ms_frame_model",
                 "ms_window_model", "ms_vote_fusion_counts", "config")


def ensemble_config(world):
    """BASELINE configs[4]: ensemble inference over a 100 k-video eval set, sharded by video.  Every rank owns 100 000 / 8 =
    12 500 videos (~7.5 M frames, ~62 GB fp32 table): at N = 8 that is the whole job; at smaller N the same per-GPU shard
    (weak scaling of the same job).  All ranks call this (one count all-reduce inside)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    try:
        import bench_ensemble
        torch.cuda.empty_cache()
        r = bench_ensemble.measure(videos=12500, reps=2, dist_init=world > 1)
        if r is None:
            return None
        res = {k: r[k] for k in ENSEMBLE_KEYS}
        res["job"] = f"{world * 12500} of the 100 000 videos of BASELINE configs[4] ({world} rank(s) x 12 500 videos)"
        return res
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def run_gpu(args):
    from multimodal_error_detection_b200 import _lib, ops, parallel
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader
    from multimodal_error_detection_b200.engine import WindowTrainStep
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200med path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, local_rank, world = parallel.init_from_env()
    device = torch.device("cuda", local_rank)
    B, K, Wm = args.batch, args.steps, max(3, args.warmup)
    kw = exp_kwargs(B, args.precision)
    kw["prefetch_sms"] = args.prefetch_sms
    ds, n_frames = build_gpu_job(args, rank, device)
    n_windows = len(ds)
    fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, device,
                                                          ds.binary_error_distribution, W)
    n_params = opt.n_params

    # ---- device-timed steps: the K + warmup batches of window indices are staged in HBM beforehand
    gen = torch.Generator().manual_seed(42)
    perm = torch.randperm(n_windows, generator=gen)
    need = (K + Wm) * B
    perm = perm.repeat((need + n_windows - 1) // n_windows)[:need].reshape(K + Wm, B)
    idx_all = perm.to(device).contiguous()                                # [K+Wm, B] int64 window indices, resident in HBM
    starts_all = ds._starts[idx_all].contiguous()
    # prefetch: K1 of step k+1 is issued inside step k (side stream, under the LSTM recurrence); every step still gathers
    # exactly one batch.  The K1 roofline below is timed with the gather at the START of the step (prefetch off).
    prefetch = bool(args.prefetch and args.precision == "bf16")
    stepper = WindowTrainStep(ds, fe, model, crit, opt, kw, B, gather_variant=args.gather_variant, prefetch=prefetch)
    graph_note = "eager"
    if args.graph:
        try:
            stepper.load(idx_all[0])
            stepper.capture()
            graph_note = "cuda_graph"
        except Exception as e:  # capture is an optimisation, never a correctness dependency
            stepper.graphs = [None, None]
            graph_note = f"eager (graph capture failed: {type(e).__name__})"
    from multimodal_error_detection_b200.modeling import modeling_utils as _mu
    _mu._set_train(model, fe, kw, True)

    def step(i):
        """One train step on batch i of idx_all (prefetch mode: batch i+1 is gathered inside it)."""
        if stepper.prefetch:
            if not stepper._primed:
                stepper.load(idx_all[i])
            stepper.run(idx_all[i + 1] if i + 1 < idx_all.shape[0] else None)
        else:
            stepper.load(idx_all[i])
            stepper.run()

    for i in range(Wm):
        step(i)
    # NVML init + the priming query cost ~10 ms on a multi-GPU box: they happen BEFORE the barrier, or rank 0 would enter the
    # timed region late and every other rank's first step would wait for it inside NCCL (measured: +7.6 ms on 20 steps)
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("B200MED_NO_CLOCKS")) else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    gather_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(K)]      # per-step marks (diagnostics: spread of the steps)
    torch.cuda.synchronize()
    parallel.barrier()
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    ev[0].record()
    for i in range(K):
        if stepper.graph is None and not stepper.prefetch:
            stepper.gather_events = gather_ev[i]
        step(Wm + i)
        step_ev[i].record()
    ev[1].record()
    torch.cuda.synchronize()
    parallel.barrier()
    step_ms = parallel.max_over_ranks(ev[0].elapsed_time(ev[1]) / K, device)
    marks = [ev[0]] + step_ev
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(K))
    step_spread = {"min": per_step[0], "median": per_step[K // 2], "max": per_step[-1]}
    stepper.gather_events = None
    launches = (_lib.launch_count() - launches0) if stepper.graph is None else stepper.launches_per_step * K
    clocks = sampler.stop() if sampler else None
    final_loss = float(stepper.loss.item())
    value = world * B / (step_ms * 1e-3)

    # ---- roofline of the dominant HBM kernel (K1), timed with CUDA events on the launch stream
    out_es = 2 if args.precision == "bf16" else 4
    fused = bool(getattr(stepper, "fused", False))
    k1_unfused_bytes = B * W * (IMAGE_DIM * (4 + out_es) + KIN_DIM * 8)
    # fused gather + first FeatureExtractor layer (csrc/gather_gemm.cu): fp32 table rows in, the bf16 batch (weight-gradient
    # operand of the backward) and the layer's bf16 output out; the events bracket that kernel alone
    k1_bytes = B * W * (IMAGE_DIM * (4 + 2) + 512 * 2) if fused else k1_unfused_bytes
    if stepper.graph is None and not stepper.prefetch:
        k1_ms = statistics.mean(a.elapsed_time(b) for a, b in gather_ev)
        k1_how = "CUDA events around the K1 launch inside each of the K timed steps"
    else:
        # the timed steps were graph replays (no events inside) and / or had K1 of the next step running on a side stream
        # under the LSTM kernels; time K1 inside the same train step launched eagerly with the gather at its start
        g, pf, stepper.graphs, stepper.prefetch = stepper.graphs, stepper.prefetch, [None, None], False
        n_ev = min(K, 8)
        k1_ms = None
        if g[0] is not None:
            # preferred: the same step captured as a CUDA graph with EXTERNAL event records around K1, so that K1 is timed
            # inside back-to-back replays (an eagerly launched step leaves the GPU idle between launches)
            try:
                ext = (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
                stepper.gather_events = ext
                stepper.load(idx_all[Wm])
                stepper.capture()
                ts = []
                for i in range(n_ev + 2):
                    stepper.load(idx_all[Wm + (i % K)])
                    stepper.run()
                    torch.cuda.synchronize()
                    ts.append(ext[0].elapsed_time(ext[1]))
                k1_ms = statistics.mean(ts[2:])
                k1_how = ("external CUDA events recorded around the K1 launches inside the replayed CUDA graph of the same train step "
                          "with the gather at the start of the step" +
                          (" (the timed steps prefetch K1 of step k+1 under step k's LSTM kernels)" if pf else ""))
            except Exception as e:      # external events in graphs are a torch >= 2.4 feature: fall back to eager timing
                k1_ms = None
                print(f"bench: in-graph K1 timing unavailable ({type(e).__name__}: {e}); timing K1 in eager steps", file=sys.stderr)
            stepper.graphs, stepper.gather_events = [None, None], None
        if k1_ms is None:
            for i in range(n_ev):
                stepper.load(idx_all[Wm + i])
                stepper.gather_events = gather_ev[i]
                stepper.run()
            torch.cuda.synchronize()
            k1_ms = statistics.mean(a.elapsed_time(b) for a, b in gather_ev[:n_ev])
            k1_how = ("CUDA events around the K1 launch inside the same train step launched eagerly with the gather at the start "
                      "of the step (the timed steps are graph replays" + (" with K1 of step k+1 prefetched under step k's LSTM kernels)" if pf else ")"))
        stepper.graphs, stepper.prefetch, stepper.gather_events, stepper._primed = g, pf, None, False
    iso = []
    for i in range(min(K, 10) + 3):      # isolated launches; every launch touches a fresh 1.1 GB slice (> L2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st_i = starts_all[i % (K + Wm)]
        a.record()
        ds.gather_batch(None, image_out=stepper.images, kin_out=stepper.kin, starts=st_i,
                        exact=args.precision != "bf16", variant=args.gather_variant)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            iso.append(a.elapsed_time(b))
    k1_iso_ms = statistics.mean(iso)
    hbm_peak, tf_peak, peak_src = measured_peaks()
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    # DRAM traffic of K1 per step: NOT measurable without a profiler attached -- taken from the committed `ncu --set full`
    # capture of this configuration (profiles/k1_traffic.json records the command, the config it was captured on and the two
    # dram__bytes counters); null when this run's configuration is not the captured one.
    traffic, traffic_src = k1_traffic(args.precision, B, W, args.gather_variant, fused)
    k1_name = ("gather_gemm_kernel (K1 fused into K2: window gather + standardise of the image stream as the A-operand producer of the "
               "FeatureExtractor's first Linear + ReLU on tcgen05; HBM-bound: 170 flop per byte, below the ridge)") if fused else \
        "gather_norm_tma_kernel + gather_norm_kernel (K1: window gather + standardise + concat, image + kinematics streams)"
    roofline = {"kernel": k1_name,
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                "share_of_step": k1_ms / step_ms, "how": k1_how, "peak_source": peak_src,
                "bytes_per_window": k1_bytes // B,
                "tflops_in_its_shadow": (2.0 * B * W * 512 * IMAGE_DIM / (k1_ms * 1e-3) / 1e12) if fused else None,
                # the standalone K1 kernel (every other loop of the package gathers through it): isolated launches, fresh slices
                "k1_standalone": {"kernel": "gather_norm_tma_kernel<bf16, 0, 8, 3, 256> + gather_norm_kernel (image + kinematics)",
                                  "bytes_per_launch": k1_unfused_bytes, "ms_per_launch_isolated": k1_iso_ms,
                                  "achieved": k1_unfused_bytes / (k1_iso_ms * 1e-3) / 1e9,
                                  "frac": k1_unfused_bytes / (k1_iso_ms * 1e-3) / 1e9 / hbm_peak}}

    # ---- tensor-pipe evidence for K2: FE layer-1 forward GEMM alone
    gemm = None
    if args.precision == "bf16" and ops.has_tcgen05():
        M = B * W
        x = stepper.images.reshape(M, IMAGE_DIM)
        w1 = ops.to_bf16(fe.linear.linear_0.weight.detach().contiguous())
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.gemm_bf16(x, w1, M, 512, IMAGE_DIM, True, True, bias=fe.linear.linear_0.bias, relu=True)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        tf = 2.0 * M * 512 * IMAGE_DIM / (statistics.mean(ts) * 1e-3) / 1e12
        gemm = {"kernel": "gemm_bf16_tcgen05_kernel<256, row-major epilogue, CTA pair / cta_group::2> (K2: FE layer-1 forward, M=B*W, N=512, K=2048)", "bound": "tensor",
                "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak, "ms_per_launch": statistics.mean(ts)}

    # ---- e2e: the public train_single_epoch over K steps; pinned-host index batches in, loss out EVERY step
    e2e = None
    if not args.no_e2e:
        kw_e2e = dict(kw, host_sync="step", prefetch_gather=prefetch)
        loader = DeviceWindowLoader(ds, B, shuffle=True, generator=torch.Generator().manual_seed(42), rank=0, world_size=1)
        full = max(1, n_windows // B)                            # full batches per pass: every timed step is a B-window graph replay
        loader.max_batches = min(Wm, full)                       # untimed warm-up pass (also captures the step's CUDA graph)
        mu.train_single_epoch(model, fe, loader, crit, opt, None, device, kw_e2e)
        torch.cuda.synchronize(); parallel.barrier()
        t0 = time.perf_counter()
        left, res = K, None
        while left > 0:                                          # K steps = as many passes over the loader as it takes
            loader.max_batches = min(left, full)
            res = mu.train_single_epoch(model, fe, loader, crit, opt, None, device, kw_e2e)
            left -= loader.max_batches
        torch.cuda.synchronize()
        dt = parallel.max_over_ranks(time.perf_counter() - t0, device)
        steps_e2e = K
        e2e = {"value": world * steps_e2e * B / dt, "unit": UNIT, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": 4,
               "steps": steps_e2e, "api": "modeling_utils.train_single_epoch(DeviceWindowLoader)", "loss": res[0],
               "table_upload_bytes_once": int(n_frames * (IMAGE_DIM + KIN_DIM + 6) * 4)}

    # ---- auxiliary configurations (all ranks take part: the strong-scaling steps and the ensemble job hold collectives)
    aux = None
    if not args.no_aux and args.precision == "bf16":
        aux = {}
        if world > 1 and B % world == 0:
            try:      # strong scaling: the SAME global batch of B windows split over the ranks
                ms, note, launches = timed_steps(ds, dict(kw), device, B // world, steps=K, warmup=3, prefetch=prefetch)
                aux["strong_scaling"] = {"value": B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "global_batch": B,
                                         "batch_per_gpu": B // world, "n_gpus": world, "scaling": "strong", "launch": note}
            except Exception as e:
                aux["strong_scaling"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            aux.update(other_configs(args, ds, device))
        if world in (1, 8) and not args.no_ensemble:
            del stepper
            opt._b200_stepper = None
            aux["ensemble_inference"] = ensemble_config(world)
    if rank == 0:
        cpu = None
        if world == 1:      # N = 1 only: at N > 1 the other ranks would spin in the final barrier and starve this rank's threads
            try:
                cores = host_threads()
                wps, dt, n_done, build_s, steps_done = cpu_train_windows_per_sec(2 * B, B, steps=args.cpu_steps)
                cpu = {"value": wps, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"{n_done} windows (W={W}) of the same workload, B={B}, {steps_done} steps of the oracle train loop "
                                 f"over a {2 * B}-window sample = {dt:.1f}s (host window build {build_s:.2f}s not included)"}
            except Exception as e:
                cpu = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {e}"}
        other = aux
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": workload_name(B, args.videos), "global_batch": world * B, "window": W, "stride": S,
                           "frames_per_gpu": n_frames, "windows_per_gpu": n_windows, "params": n_params,
                           "parallelism": f"dp{world}", "launch": graph_note,
                           "gather_prefetch": bool(prefetch), "gather_fused_into_layer1": fused,
                           "gradient_exchange": ("none (one rank)" if world == 1 else
                                                 "one kernel over NVLink peer memory per step (b200med_peer_allreduce_f32: reduce-scatter + "
                                                 "all-gather by direct peer loads / stores)" if getattr(opt, "_peer", None) is not None
                                                 else "NCCL all-reduce (peer-memory mappings unavailable or switched off)"),
                           "l2": "every step gathers a fresh ~1.1 GB slice of a ~10 GB table (inputs larger than the 126 MB L2)",
                           "gather_variant": args.gather_variant},
                "roofline": roofline, "roofline_gemm": gemm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "final_loss": final_loss, "ms_per_step_spread_rank0": step_spread,
                "other_configs": other}
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL work must be gone before the communicator is torn down; a wedged teardown must not
        # turn a finished run into a hang, so the process leaves through os._exit after a final barrier
        try:
            stepper.graphs = [None, None]
        except NameError:
            pass
        opt._b200_stepper = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        parallel.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--videos", type=int, default=2048,
                    help="videos per GPU in the synthetic table (2048 -> ~1.2 M frames, 10 GB, ~21 full batches per pass: the K = 20 "
                         "steps of the e2e measurement fit in one pass of the loader, like an epoch of a real fold does)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--gather-variant", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary measurements of the other BASELINE configs "
                    "(frame path, ensemble inference; N = 1 only, reported under 'other_configs')")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the 12 500-videos-per-GPU ensemble job (62 GB table per GPU)")
    ap.add_argument("--prefetch-sms", type=int, default=56, help="SMs the prefetching gather may occupy (side stream)")
    ap.add_argument("--prefetch", dest="prefetch", action="store_true",
                    help="gather the batch of step k+1 on a side stream inside step k (round 1's default: it hid K1 under LSTM kernels "
                         "that left 84 SMs idle; the generation-2 recurrence fills 128 SMs and the overlap no longer pays -- "
                         "profiles/r2_step_breakdown_in_graph_gen2.txt)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="steps of the in-run CPU baseline (N = 1 only; B windows each)")
    ap.add_argument("--cpu-windows", type=int, default=0, help="(ignored, kept for old command lines)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
