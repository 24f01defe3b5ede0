"""LSTM head recurrence microbenchmark: forward / backward of the 3-layer stack at the headline shape (B = 8192, W = 16, H = 128),
generation 1 (lstm_rec.cu behind x-part / dX GEMMs) against generation 2 (lstm_rec2.cu, 2-CTA clusters, x-part and dX fused).
CUDA-event timing after warm-up; prints one JSON line.

    python scripts/bench_lstm.py [--batch 8192] [--window 16] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_error_detection_b200 import lstm_stack  # noqa: E402


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--window", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    torch.manual_seed(0)
    dev = "cuda"
    B, W, F, H = a.batch, a.window, 58, 128
    lstm = torch.nn.LSTM(F, H, num_layers=3, batch_first=True, dropout=0.2).to(dev)
    x = torch.randn(B, W, F, device=dev)
    gh = torch.randn(B, H, device=dev)
    seed = torch.tensor([3], dtype=torch.int32, device=dev)
    res = {"B": B, "W": W}
    outs = {}
    for gen in (1, 2):
        lstm_stack.REC_GEN = gen

        def fwd_only():
            with torch.no_grad():
                return lstm_stack.lstm_last_hidden(x.permute(0, 2, 1), lstm, training=True, seed_dev=seed)

        def fwd_bwd():
            lstm.zero_grad(set_to_none=True)
            xo = x.clone().requires_grad_(True)
            h = lstm_stack.lstm_last_hidden(xo.permute(0, 2, 1), lstm, training=True, seed_dev=seed)
            h.backward(gh)
            return h.detach(), xo.grad

        h, dx = fwd_bwd()
        outs[gen] = (h.clone(), dx.clone(), [p.grad.clone() for p in lstm.parameters()])
        res[f"gen{gen}"] = {"fwd_nograd_ms": timed(fwd_only, a.reps), "fwd_bwd_ms": timed(fwd_bwd, a.reps)}
    nrel = lambda u, v: float((u - v).norm() / v.norm().clamp_min(1e-12))
    res["gen2_vs_gen1"] = {"h": nrel(outs[2][0], outs[1][0]), "dx": nrel(outs[2][1], outs[1][1]),
                           "max_param_grad": max(nrel(u, v) for u, v in zip(outs[2][2], outs[1][2]))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
