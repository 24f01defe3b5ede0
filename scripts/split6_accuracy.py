"""Accuracy of an fp32 product emulated by bf16 tensor-core products of 3-way split operands (x = h + m + l, six products
hh hm mh hl lh mm folded into one reduction of 6K) against fp64, next to the fp32 SIMT GEMM and the 2-way split (3 products)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_error_detection_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def split3(x):
    h = x.to(torch.bfloat16)
    r = x - h.float()
    m = r.to(torch.bfloat16)
    l = (r - m.float()).to(torch.bfloat16)
    return h, m, l


for (M, N, K, relu_in) in [(32768, 512, 2048, True), (32768, 256, 512, False), (4096, 512, 256, False)]:
    x = torch.randn(M, K, device=dev)
    if relu_in:
        x = x.clamp_min(0)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    ref = (x.double() @ w.double().T)
    xh, xm, xl = split3(x)
    wh, wm, wl = split3(w)
    A6 = torch.cat([xh, xh, xm, xh, xl, xm], dim=1).contiguous()
    B6 = torch.cat([wh, wm, wh, wl, wh, wm], dim=1).contiguous()
    y6 = ops.gemm_bf16(A6, B6, M, N, 6 * K, True, True, out_dtype=torch.float32)
    A3 = torch.cat([xh, xm, xh], dim=1).contiguous()
    B3 = torch.cat([wh, wh, wm], dim=1).contiguous()
    y3 = ops.gemm_bf16(A3, B3, M, N, 3 * K, True, True, out_dtype=torch.float32)
    # ordering variant: small terms first, the hh block last
    A6b = torch.cat([xm, xl, xh, xm, xh, xh], dim=1).contiguous()
    B6b = torch.cat([wm, wh, wl, wh, wm, wh], dim=1).contiguous()
    y6b = ops.gemm_bf16(A6b, B6b, M, N, 6 * K, True, True, out_dtype=torch.float32)
    for sk in (2, 4, 8, 16, 32):
        yk = ops.gemm_bf16(A6b, B6b, M, N, 6 * K, True, True, out_dtype=torch.float32, split_k=sk)
        print(f"   split6(small first) split_k={sk}: {float((yk.double() - ref).norm() / ref.norm()):.3e}")
    old = ops.FP32_TC_MIN_FLOP
    ops.FP32_TC_MIN_FLOP = 0.0
    ys = ops.linear_f32(x, w)
    ops.FP32_TC_MIN_FLOP = old
    yt = x @ w.T
    def rel(y):
        return float((y.double() - ref).norm() / ref.norm()), float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"M={M} N={N} K={K}: split6 {rel(y6)}  split6(small first) {rel(y6b)}  split3 {rel(y3)}  simt fp32 {rel(ys)}  torch fp32 {rel(yt)}")
    print("   mean signed err / |ref| mean: split6", float(((y6.double() - ref) * ref.sign()).mean() / ref.abs().mean()), " simt", float(((ys.double() - ref) * ref.sign()).mean() / ref.abs().mean()))
