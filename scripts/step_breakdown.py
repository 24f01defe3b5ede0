"""Phase breakdown of the replayed train step, measured with EXTERNAL CUDA events recorded inside the CUDA graph
(no profiler, no eager launch gaps).  Prints milliseconds per phase for the prefetching and the non-prefetching step."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodal_error_detection_b200.engine import WindowTrainStep
from multimodal_error_detection_b200.modeling import modeling_utils as mu


class A:
    videos, batch, precision, gather_variant = 1024, 8192, "bf16", 0


dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
kw = bench.exp_kwargs(A.batch, A.precision)
ds, _ = bench.build_gpu_job(A, 0, dev)
fe, model, crit, opt, sched = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, dev,
                                                      ds.binary_error_distribution, bench.W)
idx = torch.randperm(len(ds), generator=torch.Generator().manual_seed(0)).repeat(2)[: 12 * A.batch].reshape(12, A.batch).to(dev)
names = ["index+labels, K1 (if not prefetched)", "FE forward (+cat)", "head forward (pack, x-part GEMMs, LSTM recurrence, MLP)",
         "loss", "head backward (MLP, LSTM recurrence, dX GEMMs)", "FE backward (+ joined LSTM weight gradients)",
         "gradient collect + Adam", "join of the prefetch stream"]
for prefetch in ((False,) if os.environ.get("ONLY_NO_PREFETCH") else (False, True)):
    st = WindowTrainStep(ds, fe, model, crit, opt, kw, A.batch, prefetch=prefetch)
    st.phase_events = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(9)]
    mu._set_train(model, fe, kw, True)
    st.load(idx[0])
    st.capture()
    rows = []
    for i in range(10):
        if prefetch:
            if not st._primed:
                st.load(idx[i])
            st.run(idx[i + 1])
        else:
            st.load(idx[i])
            st.run()
        torch.cuda.synchronize()
        ev = st.phase_events
        order = [0, 1, 2, 3, 4, 5, 6, 7, 8]
        rows.append([ev[a].elapsed_time(ev[b]) for a, b in zip(order[:-1], order[1:])] + [ev[0].elapsed_time(ev[8])])
    med = [statistics.median(r[k] for r in rows[3:]) for k in range(9)]
    print(f"--- prefetch={prefetch}: step {med[8]:.3f} ms")
    for n, v in zip(names, med[:8]):
        print(f"  {v:7.3f} ms  {n}")
