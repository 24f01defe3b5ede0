"""Kernel timeline of ONE replayed train step (CUPTI through torch.profiler: start, duration, stream, grid of every kernel
node of the replayed CUDA graph).  The profiler adds overhead per node, so absolute times are longer than the bench's;
what this shows is WHICH kernels run side by side and where the main stream waits.  Writes gpurun_out/step_timeline.tsv."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from multimodal_error_detection_b200.engine import WindowTrainStep
from multimodal_error_detection_b200.modeling import modeling_utils as mu


class A:
    videos, gather_variant = 1024, 0
    batch = int(os.environ.get("BATCH", "8192"))
    precision = os.environ.get("PRECISION", "bf16")


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
kw = bench.exp_kwargs(A.batch, A.precision)
ds, _ = bench.build_gpu_job(A, 0, dev)
fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, dev,
                                                      ds.binary_error_distribution, bench.W)
idx = torch.randperm(len(ds), generator=torch.Generator().manual_seed(0)).repeat(2)[: 12 * A.batch].reshape(12, A.batch).to(dev)
st = WindowTrainStep(ds, fe, model, crit, opt, kw, A.batch, prefetch=False)
mu._set_train(model, fe, kw, True)
st.load(idx[0])
st.capture()
for i in range(4):
    st.load(idx[i]); st.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(4, 7):
        st.load(idx[i]); st.run()
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(out, exist_ok=True)
path = os.path.join(out, "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
with open(os.path.join(out, "step_timeline.tsv"), "w") as f:
    f.write("ts_us\tdur_us\tstream\tgrid\tblock\tname\n")
    t0 = ev[0]["ts"] if ev else 0
    for e in ev:
        a = e.get("args", {})
        f.write(f"{e['ts'] - t0:.2f}\t{e['dur']:.2f}\t{a.get('stream')}\t{a.get('grid')}\t{a.get('block')}\t{e['name'][:90]}\n")
os.remove(path)
print(len(ev), "device activities")
