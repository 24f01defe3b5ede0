"""Micro-benchmark of K1 (gather + standardise): GB/s per variant / output dtype / batch, CUDA-event timed."""
import sys, os, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_error_detection_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
N, W = 600_000, 16
image = torch.randn(N, 2048, device=dev, generator=g)
kin = torch.randn(N, 26, device=dev, generator=g)
mi, si = torch.randn(2048, device=dev), torch.rand(2048, device=dev) + 0.5
mk, sk = torch.randn(26, device=dev), torch.rand(26, device=dev) + 0.5
peak = 6547.2
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2".split(","))]
for B in (512, 2048, 8192):
    for odt, oname, es in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
        for variant in variants:
            for exact in ((True, False) if odt == torch.float32 else (False,)):
                out = torch.empty(B, W, 2048, device=dev, dtype=odt)
                kout = torch.empty(B, W, 26, device=dev)
                ts = []
                for it in range(13):
                    starts = torch.randint(0, N - W, (B,), device=dev, generator=g).to(torch.int32)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    try:
                        ops.gather_norm([ops.GatherStream(image, mi, si, out, 0, exact), ops.GatherStream(kin, mk, sk, kout, 0, True)], starts, W, variant)
                    except Exception as e:
                        print("variant", variant, "failed:", e); break
                    b.record(); torch.cuda.synchronize()
                    if it >= 3: ts.append(a.elapsed_time(b))
                if not ts: continue
                ms = statistics.median(ts)
                nbytes = B * W * (2048 * (4 + es) + 26 * 8)
                print(json.dumps({"B": B, "out": oname, "variant": variant, "exact": exact, "ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3)}))
