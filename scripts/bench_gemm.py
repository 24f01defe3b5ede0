"""Time the GEMM shapes of the train step (CUDA events), or run once per shape for ncu (--once)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_error_detection_b200 import ops

once = "--once" in sys.argv
dev = "cuda"
M = 131072
shapes = [  # name, M, N, K, a_k, b_k, out dtype, rbi, mask, split
    ("FE L1 fwd", M, 512, 2048, True, True, torch.bfloat16, False, False, 1),
    ("FE L2 fwd", M, 256, 512, True, True, torch.bfloat16, False, False, 1),
    ("XG K=64", M, 512, 64, True, True, torch.float16, True, False, 1),
    ("XG K=128", M, 512, 128, True, True, torch.float16, True, False, 1),
    ("dgrad L2 (mask)", M, 512, 256, True, False, torch.bfloat16, False, True, 1),
    ("dX K=512 N=128 rbi f32", M, 128, 512, True, False, torch.float32, True, False, 1),
    ("wgrad L1", 512, 2048, M, False, False, torch.float32, False, False, 4),
    ("wgrad LSTM", 512, 256, M, False, False, torch.float32, False, False, 16),
]
for name, m, n, k, ak, bk, dt, rbi, mask, split in shapes:
    A = torch.randn((m, k) if ak else (k, m), device=dev).to(torch.bfloat16)
    B = torch.randn((n, k) if bk else (k, n), device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev) if split == 1 else None
    mk = (torch.randn(m, n, device=dev) > 0).to(torch.bfloat16) if mask else None
    out = torch.empty(m, n, device=dev, dtype=dt)
    ts = []
    for it in range(1 if once else 8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.gemm_bf16(A, B, m, n, k, ak, bk, bias=bias, mask=mk, out=out, out_dtype=dt, split_k=split, rbi=rbi)
        b.record(); torch.cuda.synchronize()
        if it >= 3 or once: ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    byts = (m * k + n * k) * 2 + m * n * out.element_size() + (m * n * 2 if mask else 0)
    print(f"{name:26s} {ms*1e3:8.1f} us  {2*m*n*k/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s")
