"""A chain of N small dependent launches of the library (to_bf16 / to_f32 ping-pong on a 64 K-element buffer) replayed as a
CUDA graph, with the programmatic dependent launches (b200med_set_pdl) on and off: what one launch on the critical path costs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_error_detection_b200 import ops, _lib

dev = torch.device("cuda:0")
x = torch.randn(1 << 16, device=dev)
xb = torch.empty(1 << 16, device=dev, dtype=torch.bfloat16)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100


def chain():
    for _ in range(N // 2):
        ops.to_bf16(x, out=xb)
        x.copy_(ops.to_f32(xb)) if False else _lib.call("b200med_cast_bf16_to_f32", ops._ptr(xb), ops._ptr(x), x.numel(), ops._stream())


for pdl in (1, 0, 1, 0):
    _lib.load().b200med_set_pdl(pdl)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            chain()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
        tg = e0.elapsed_time(e1) / 20
        e0.record(s)
        for _ in range(20):
            chain()
        e1.record(s)
        torch.cuda.synchronize()
        te = e0.elapsed_time(e1) / 20
    print(f"pdl={pdl}: graph replay {tg * 1e3 / N:.2f} us per launch, eager {te * 1e3 / N:.2f} us per launch ({N} launches)")
