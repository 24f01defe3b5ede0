"""Host-side profile (cProfile) of the e2e measurement of bench.py: K steps of modeling_utils.train_single_epoch over the
DeviceWindowLoader, after a warm-up pass that captures the step's CUDA graph."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader  # noqa: E402
from multimodal_error_detection_b200.modeling import modeling_utils as mu  # noqa: E402


class A:
    videos, batch, precision, gather_variant = 2048, 8192, "bf16", 0


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
kw = dict(bench.exp_kwargs(A.batch, A.precision), host_sync="step")
ds, _ = bench.build_gpu_job(A, 0, dev)
fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, dev, ds.binary_error_distribution, bench.W)
loader = DeviceWindowLoader(ds, A.batch, shuffle=True, generator=torch.Generator().manual_seed(42), rank=0, world_size=1)
loader.max_batches = 3
mu.train_single_epoch(model, fe, loader, crit, opt, None, dev, kw)
torch.cuda.synchronize()
loader.max_batches = 20
for rep in range(2):
    t0 = time.perf_counter()
    mu.train_single_epoch(model, fe, loader, crit, opt, None, dev, kw)
    torch.cuda.synchronize()
    print(f"epoch of 20 steps: {1e3 * (time.perf_counter() - t0):.2f} ms")
pr = cProfile.Profile()
pr.enable()
mu.train_single_epoch(model, fe, loader, crit, opt, None, dev, kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
