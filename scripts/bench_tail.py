"""In-graph cost of the LSTM head's Linear / ReLU / BatchNorm tail (heads.mlp_tail) at the bench batch: forward + backward
captured in a CUDA graph and replayed (CUDA events), fp32 and bf16 (split-bf16 tensor-core) modes."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_error_detection_b200 import _lib  # noqa: E402
from multimodal_error_detection_b200.heads import mlp_tail  # noqa: E402


def main(B=8192):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Flatten(), nn.Linear(128, 256), nn.ReLU(), nn.BatchNorm1d(256), nn.Linear(256, 64), nn.ReLU(), nn.BatchNorm1d(64),
                        nn.Linear(64, 1)).to(dev).train()
    h = torch.randn(B, 128, device=dev, requires_grad=True)
    dy = torch.randn(B, 1, device=dev)
    out = {}
    for precision in ("fp32", "bf16"):
        def body():
            for p in seq.parameters():
                p.grad = None
            h.grad = None
            y = mlp_tail(h, seq, relu_in=True, training=True, precision=precision)
            y.backward(dy)

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        launches = _lib.launch_count() - n0
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        out[precision] = {"ms_fwd_bwd_in_graph": a.elapsed_time(b) / 50, "b200med_launches": launches}
    print(json.dumps({"B": B, **out}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8192)
