"""Time the fused gather + first-Linear kernel alone (CUDA events, fresh windows every launch so the table reads miss L2).

Experiment knobs are environment variables read once by the library: B200MED_GG_RINGS (ring depths fp32-staging / A / B as
three digits: 323, 422, 224, 233, 332) and B200MED_GG_DEBUG (bits: 1 no Xb store, 2 no Y store, 4 no MMAs, 8 no conversion,
16 no W1 loads, 32 no evict-first hints, 64 one MMA per k-step).  `--sweep` re-runs itself once per combination
(GG_RINGS / GG_DEBUGS: comma-separated lists).
"""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(args):
    import torch
    from multimodal_error_detection_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(0)
    rows, K, W, B = args.rows, 2048, args.window, args.batch
    table = torch.randn(rows, K, device=dev, generator=g)
    mean = torch.randn(K, device=dev, generator=g) * 0.1
    std = torch.rand(K, device=dev, generator=g) + 0.5
    w = (torch.randn(512, K, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    bias = torch.randn(512, device=dev, generator=g) * 0.1
    starts = [torch.randint(0, rows - W, (B,), device=dev, generator=g, dtype=torch.int32) for _ in range(args.iters + 3)]
    for i in range(3):
        ops.gather_linear_bf16(table, mean, std, starts[i], W, w, bias, True)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for i in range(args.iters):
        ops.gather_linear_bf16(table, mean, std, starts[3 + i], W, w, bias, True, events=ev[i])
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = ms[len(ms) // 2]
    alg = B * W * (K * 4 + K * 2 + 512 * 2)
    print(json.dumps({"rings": os.environ.get("B200MED_GG_RINGS", "323"), "debug": os.environ.get("B200MED_GG_DEBUG", "0"), "B": B, "W": W,
                      "ms_median": round(med, 4), "ms_min": round(ms[0], 4), "GBps_algorithmic": round(alg / med / 1e6, 1)}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--window", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    if args.sweep:
        for rings in os.environ.get("GG_RINGS", "323,422").split(","):
            for dbg in os.environ.get("GG_DEBUGS", "0,1,2,3").split(","):
                env = dict(os.environ, B200MED_GG_RINGS=rings, B200MED_GG_DEBUG=dbg)
                subprocess.run([sys.executable, __file__, "--rows", str(args.rows), "--batch", str(args.batch), "--window", str(args.window),
                                "--iters", str(args.iters)], env=env, check=False)
    else:
        one(args)
