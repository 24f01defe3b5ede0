"""Tensor-level wrappers over the C ABI.  torch is used only for device memory and streams.

Every function takes CUDA tensors, launches on ``torch.cuda.current_stream()`` and raises if it is
handed CPU tensors -- the product path has no CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import BF16, F16, F32, StreamDesc, call


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t: torch.Tensor, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"b200med: {name} must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"b200med: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"b200med: {name} must be contiguous")
    return t


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"b200med: unsupported dtype {t.dtype}")


_ws_cache = {}


class sm_limit:
    """``with ops.sm_limit(n):`` -- persistent kernels launched inside fill at most n SMs (side-stream work that runs next to a
    kernel which must keep its SMs)."""

    def __init__(self, sms: int):
        self.sms = int(sms)

    def __enter__(self):
        self.prev = _lib.load().b200med_set_sm_limit(self.sms)
        return self

    def __exit__(self, *exc):
        _lib.load().b200med_set_sm_limit(self.prev)
        return False


def workspace(nbytes: int, device, tag: str = "default", zero: bool = False) -> torch.Tensor:
    """Grow-only scratch buffers keyed by (device, stream, tag); stream-ordered reuse is safe.  Under stream capture the
    buffer is an ordinary temporary of the graph being recorded (allocated from ITS pool, never cached): a cached buffer
    would tie later graphs to memory of a pool whose graph may be gone by the time they replay."""
    if torch.cuda.is_current_stream_capturing():
        n = max(int(nbytes), 256)
        return torch.zeros(n, dtype=torch.uint8, device=device) if zero else torch.empty(n, dtype=torch.uint8, device=device)
    key = (str(device), torch.cuda.current_stream().cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        n = max(int(nbytes), 256)
        buf = torch.zeros(n, dtype=torch.uint8, device=device) if zero else torch.empty(n, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------------- K0
def window_index(g: torch.Tensor, subj_offsets: torch.Tensor, W: int, S: int, e5: Optional[torch.Tensor] = None):
    """Window start rows (int32) of a flat table with contiguous subjects, plus the first-frame
    labels.  Returns dict(starts, g_win, e5_win, subj_win, win_offsets).  Raises IndexError when a
    subject has no non-zero gesture, as the reference does (dataset_utils.py:211-212)."""
    g = _need(g.reshape(-1), torch.float32, "g")
    off = _need(subj_offsets, torch.int64, "subj_offsets")
    n_subj = off.numel() - 1
    dev = g.device
    if g.numel() == 0 or n_subj <= 0:
        z = lambda dt, *shape: torch.zeros(*shape, dtype=dt, device=dev)
        return dict(starts=z(torch.int32, 0), g_win=z(torch.float32, 0), e5_win=None if e5 is None else z(torch.float32, 0, 5),
                    subj_win=z(torch.int32, 0), win_offsets=z(torch.int64, max(n_subj, 0) + 1))
    win_off = torch.empty(n_subj + 1, dtype=torch.int64, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    call("b200med_window_count", _ptr(g), _ptr(off), n_subj, W, S, _ptr(win_off), _ptr(status), _stream())
    host = torch.stack([win_off[-1], status[0].to(torch.int64)]).cpu()  # one sync: index build is one-off
    total, bad = int(host[0]), int(host[1])
    if bad >= 0:
        raise IndexError(f"index 0 is out of bounds: subject #{bad} has no non-zero gesture")
    starts = torch.empty(total, dtype=torch.int32, device=dev)
    g_win = torch.empty(total, dtype=torch.float32, device=dev)
    subj_win = torch.empty(total, dtype=torch.int32, device=dev)
    e5_win = None
    if e5 is not None:
        e5 = _need(e5, torch.float32, "e5")
        e5_win = torch.empty(total, 5, dtype=torch.float32, device=dev)
    call("b200med_window_fill", _ptr(g), _ptr(off), _ptr(win_off), n_subj, W, S, _ptr(e5), _ptr(starts), _ptr(g_win),
         _ptr(e5_win), _ptr(subj_win), _stream())
    return dict(starts=starts, g_win=g_win, e5_win=e5_win, subj_win=subj_win, win_offsets=win_off)


def powerset(e5: torch.Tensor, delete_nd: bool = True):
    e5 = _need(e5, torch.float32, "e5")
    n = e5.shape[0]
    e7 = torch.empty(n, 7, dtype=torch.int32, device=e5.device)
    mask = torch.empty(n, dtype=torch.uint8, device=e5.device)
    call("b200med_powerset", _ptr(e5), n, int(bool(delete_nd)), _ptr(e7), _ptr(mask), _stream())
    return e7, mask.bool()


# ------------------------------------------------------------------------------------------- K1
class GatherStream:
    """One modality stream for :func:`gather_norm` (see b200med_stream_desc)."""

    def __init__(self, table, mean=None, std=None, out=None, out_col=0, exact_div=True):
        self.table, self.mean, self.std, self.out, self.out_col, self.exact_div = table, mean, std, out, out_col, exact_div


def gather_norm(streams: Sequence[GatherStream], starts: torch.Tensor, W: int, variant: int = 0):
    """out_s[b*W+t, col_s:col_s+D_s] = (table_s[starts[b]+t] - mean_s) / std_s for every stream."""
    starts = _need(starts, torch.int32, "starts")
    B = starts.numel()
    arr = (StreamDesc * len(streams))()
    keep = []
    for i, s in enumerate(streams):
        table = _need(s.table, None, "table")
        out = _need(s.out, None, "out")
        D = table.shape[-1]
        out2 = out.view(-1, out.shape[-1])
        if out2.shape[0] != B * W:
            raise ValueError(f"out has {out2.shape[0]} rows, expected B*W = {B * W}")
        mean = std = None
        stat_rows = 1
        if s.mean is not None:
            mean = _need(s.mean.reshape(-1, D), torch.float32, "mean")
            std = _need(s.std.reshape(-1, D), torch.float32, "std")
            stat_rows = mean.shape[0]
            keep += [mean, std]
        arr[i] = StreamDesc(table.data_ptr(), 0 if mean is None else mean.data_ptr(), 0 if std is None else std.data_ptr(),
                            out.data_ptr(), D, _dt(table), _dt(out), out2.shape[1], s.out_col, stat_rows,
                            int(bool(s.exact_div)), min(table.numel() // D, 2 ** 31 - 1))
    call("b200med_gather_norm", arr, len(streams), _ptr(starts), B, W, variant, _stream())


def gather_last_variant() -> int:
    """Device path of this thread's last gather_norm call (1 = LDG kernel, 2.. = TMA staging ring shapes)."""
    return int(_lib.load().b200med_gather_last_variant())


def gather_linear_supported(table, W: int, n_out: int, stat_rows: int = 1) -> bool:
    """Shapes the fused gather + first-layer kernel (b200med_gather_linear_bf16) serves."""
    return (table.dtype == torch.float32 and table.is_contiguous() and table.dim() == 2 and table.shape[1] % 64 == 0
            and W in (16, 32, 64, 128) and n_out == 512 and stat_rows == 1 and has_tcgen05())


def gather_linear_bf16(table, mean, std, starts, W: int, w_bf16, bias, relu: bool = True, events=None, want_xb: bool = True):
    """(xb [B*W, K] bf16, y [B*W, 512] bf16): standardised bf16 windows of `table` and relu(xb w^T + bias) in ONE kernel.
    want_xb=False (inference: no backward will read the batch): xb is not written and None is returned in its place."""
    table = _need(table, torch.float32, "table")
    starts = _need(starts, torch.int32, "starts")
    w_bf16 = _need(w_bf16, torch.bfloat16, "w")
    B, K, N = starts.numel(), table.shape[1], w_bf16.shape[0]
    if mean is not None:
        mean = _need(mean.reshape(-1), torch.float32, "mean"); std = _need(std.reshape(-1), torch.float32, "std")
        if mean.numel() != K or std.numel() != K:
            raise ValueError("the fused gather takes one mean / std per column")
    xb = torch.empty(B * W, K, dtype=torch.bfloat16, device=table.device) if want_xb else None
    y = torch.empty(B * W, N, dtype=torch.bfloat16, device=table.device)
    if events is not None:
        events[0].record()
    call("b200med_gather_linear_bf16", _ptr(table), table.shape[0], _ptr(mean), _ptr(std), _ptr(starts), B, W, _ptr(w_bf16),
         _ptr(None if bias is None else _need(bias, torch.float32, "bias")), int(bool(relu)), _ptr(xb), _ptr(y), N, K, _stream())
    if events is not None:
        events[1].record()
    return xb, y


def expand_stat(stat, D: int, W: int, device) -> torch.Tensor:
    """Bring a standardisation statistic of any shape broadcastable against [W, D] (scalar, [D],
    [1, D], [W, D]; the on-disk format is unpinned by the reference) to [1|W, D]."""
    t = torch.as_tensor(stat, dtype=torch.float32)
    if t.dim() == 2 and t.shape[0] == W and W > 1:
        t = t.expand(W, D)
    else:
        t = torch.broadcast_to(t.reshape(-1) if t.numel() in (1, D) else t, (D,)).reshape(1, D)
    return t.contiguous().to(device)


def standardise_rows(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, out=None, out_col=0):
    x = _need(x, torch.float32, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    call("b200med_standardise_rows", _ptr(x), _ptr(_need(mean.reshape(-1), torch.float32)), _ptr(_need(std.reshape(-1), torch.float32)),
         _ptr(out), rows, D, out.shape[-1], out_col, _stream())
    return out


# ------------------------------------------------------------------------------------------- K2 fp32
def linear_fwd_f32(x, w, b, relu: bool, out=None):
    x = _need(x, torch.float32, "x"); w = _need(w, torch.float32, "w")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"linear_fwd_f32: x has {K} columns, w has {w.shape[1]}")
    y = torch.empty(M, N, dtype=torch.float32, device=x.device) if out is None else _need(out, torch.float32, "out")
    if _fp32_tc(M, N, K, 6 * K) and y.shape[-1] == N:
        return gemm_bf16(split_bf16x6(x, 0)[0], split_bf16x6(w, 1)[0], M, N, 6 * K, True, True, bias=b, relu=relu, out=y,
                         split_k=_fp32_tc_split(6 * K))
    call("b200med_linear_fwd_f32", _ptr(x), _ptr(w), _ptr(b), _ptr(y), M, N, K, int(relu), _stream())
    return y


def linear_bwd_data_f32(dy, w, relu_out=None):
    dy = _need(dy, torch.float32, "dy"); w = _need(w, torch.float32, "w")
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
    if _fp32_tc(M, K, N, 6 * N, K):
        # dx [M, K] = dy [M, 6N] w6 [6N, K] (right operand stacked: reduction over its rows); ReLU mask = bf16(relu_out) > 0
        mask = to_bf16(_need(relu_out, torch.float32, "relu_out")) if relu_out is not None else None
        return gemm_bf16(split_bf16x6(dy, 0)[0], split_bf16x6(w, 1, False, True)[1], M, K, 6 * N, True, False, mask=mask, out=dx,
                         split_k=_fp32_tc_split(6 * N))
    call("b200med_linear_bwd_data_f32", _ptr(dy), _ptr(w), _ptr(relu_out), _ptr(dx), M, N, K, _stream())
    return dx


def _wgrad_fp32_tc(dy, x, relu_x: bool, dw):
    """dW [N, K] = dy [M, N]^T x [M, K] as six stacked bf16 products (both operands MN-major, reduction over 6M rows,
    deterministic split-K)."""
    M, N = dy.shape
    K = x.shape[1]
    a = split_bf16x6(dy, 0, False, True)[1]
    b = split_bf16x6(x, 1, False, True, relu=relu_x)[1]
    return gemm_bf16(a, b, N, K, 6 * M, a_kmajor=False, b_kmajor=False, split_k=max(gemm_split_k(N, K, 6 * M), _fp32_tc_split(6 * M)), out=dw)


def linear_bwd_weight_f32(dy, x, want_bias=True):
    dy = _need(dy, torch.float32, "dy"); x = _need(x, torch.float32, "x")
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
    db = torch.empty(N, dtype=torch.float32, device=dy.device) if want_bias else None
    if _fp32_tc(N, K, M, N, K):
        dw = _wgrad_fp32_tc(dy, x, False, dw)
        return dw, (colsum(dy) if want_bias else None)
    ws = workspace(_lib.load().b200med_linear_bwd_weight_ws_bytes(M, N, K), dy.device, "wgrad_f32")
    call("b200med_linear_bwd_weight_f32", _ptr(dy), _ptr(x), _ptr(dw), _ptr(db), M, N, K, 0, _ptr(ws), _stream())
    return dw, db


GEMM_RELU, GEMM_ACCUM, GEMM_RELU_A, GEMM_RELU_B, GEMM_SPLIT = 1, 2, 4, 8, 16


def gemm_f32(A, B, C_out, I, J, R, a_rs, a_cs, b_rs, b_cs, ldc=None, bias=None, mask=None, ld_mask=0, flags=0):
    """General strided fp32 product (b200med_gemm_f32): C[i,j] = epi(sum_r A[i*a_rs + r*a_cs] * B[j*b_rs + r*b_cs]).
    A / B may be tensors or (tensor, element offset) pairs; strides are in elements."""
    def base(t):
        if isinstance(t, tuple):
            return C.c_void_p(_need(t[0], torch.float32, "operand").data_ptr() + 4 * int(t[1]))
        return _ptr(_need(t, torch.float32, "operand"))
    Cn = _need(C_out, torch.float32, "C")
    ws = None
    if flags & GEMM_SPLIT:
        ws = workspace(_lib.load().b200med_gemm_f32_ws_bytes(I, J, R), Cn.device, "gemm_f32_split")
    call("b200med_gemm_f32", base(A), base(B), _ptr(Cn), I, J, R, a_rs, a_cs, b_rs, b_cs, J if ldc is None else ldc,
         _ptr(bias), _ptr(mask), ld_mask, flags, _ptr(ws), _stream())
    return Cn


def linear_f32(x, w, b=None, flags=0, out=None):
    """y [M, N] = x [M, K] w[N, K]^T (+ b) with B200MED_GEMM_* flags (RELU, ACCUM into `out`, RELU_A on x)."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device) if out is None else out
    if not (flags & ~(GEMM_RELU | GEMM_RELU_A)) and _fp32_tc(M, N, K, 6 * K) and y.is_contiguous() and y.shape[-1] == N:
        return gemm_bf16(split_bf16x6(_need(x, torch.float32, "x"), 0, relu=bool(flags & GEMM_RELU_A))[0],
                         split_bf16x6(_need(w, torch.float32, "w"), 1)[0], M, N, 6 * K, True, True, bias=b,
                         relu=bool(flags & GEMM_RELU), out=y.view(M, N), split_k=_fp32_tc_split(6 * K))
    return gemm_f32(x, w, y, M, N, K, K, 1, K, 1, N, bias=b, flags=flags)


def linear_dgrad_f32(dy, w, mask=None, out=None):
    """dx [M, K] = dy [M, N] w [N, K], zeroed where mask [M, K] <= 0."""
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=torch.float32, device=dy.device) if out is None else out
    if _fp32_tc(M, K, N, 6 * N, K) and dx.is_contiguous() and dx.shape[-1] == K:
        mk = to_bf16(_need(mask, torch.float32, "mask")) if mask is not None else None
        return gemm_bf16(split_bf16x6(_need(dy, torch.float32, "dy"), 0)[0], split_bf16x6(_need(w, torch.float32, "w"), 1, False, True)[1],
                         M, K, 6 * N, True, False, mask=mk, out=dx.view(M, K), split_k=_fp32_tc_split(6 * N))
    return gemm_f32(dy, w, dx, M, K, N, N, 1, 1, K, K, mask=mask, ld_mask=K)


def linear_wgrad_f32(dy, x, relu_x=False):
    """dW [N, K] = dy [M, N]^T x [M, K] (x clamped at zero when relu_x), deterministic split reduction."""
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
    if _fp32_tc(N, K, M, N, K):
        return _wgrad_fp32_tc(_need(dy, torch.float32, "dy"), _need(x, torch.float32, "x"), relu_x, dw)
    return gemm_f32(dy, x, dw, N, K, M, 1, N, 1, K, K, flags=GEMM_SPLIT | (GEMM_RELU_B if relu_x else 0))


def bn_fwd(x, gamma, beta, eps, momentum, training, running_mean=None, running_var=None, num_batches=None):
    """nn.BatchNorm1d over the rows of x [M, C] -> (y, save_mean, save_rstd) (the last two None in inference)."""
    x = _need(x, torch.float32, "x")
    M, Cn = x.shape
    y = torch.empty_like(x)
    sm = sr = ws = None
    if training:
        sm = torch.empty(Cn, dtype=torch.float32, device=x.device)
        sr = torch.empty(Cn, dtype=torch.float32, device=x.device)
        ws = workspace(_lib.load().b200med_bn_ws_bytes(M, Cn), x.device, "bn")
    call("b200med_bn_fwd", _ptr(x), M, Cn, _ptr(gamma), _ptr(beta), float(eps), float(momentum), int(bool(training)),
         _ptr(running_mean), _ptr(running_var), _ptr(num_batches), _ptr(y), _ptr(sm), _ptr(sr), _ptr(ws), _stream())
    return y, sm, sr


def bn_bwd(dy, x, gamma, save_mean, save_rstd, relu_mask=False, want_affine=True):
    dy = _need(dy, torch.float32, "dy"); x = _need(x, torch.float32, "x")
    M, Cn = x.shape
    dx = torch.empty_like(x)
    dg = torch.empty(Cn, dtype=torch.float32, device=x.device) if want_affine else None
    db = torch.empty(Cn, dtype=torch.float32, device=x.device) if want_affine else None
    ws = workspace(_lib.load().b200med_bn_ws_bytes(M, Cn), x.device, "bn")
    call("b200med_bn_bwd", _ptr(dy), _ptr(x), M, Cn, _ptr(gamma), _ptr(save_mean), _ptr(save_rstd), int(bool(relu_mask)), _ptr(dx),
         _ptr(dg), _ptr(db), _ptr(ws), _stream())
    return dx, dg, db


# ---- the heads' Linear / ReLU / BatchNorm tail as three kernels per direction (csrc/mlp_tail.cu)
def tail_supported(K: int, N: int, kind: int) -> bool:
    return bool(_lib.load().b200med_tail_supported(int(K), int(N), int(kind)))


def _tail_part(M: int, rows: int, width: int, device) -> torch.Tensor:
    return torch.empty(int(_lib.load().b200med_tail_slabs(M)) * rows * width, dtype=torch.float32, device=device)


def _tail_bn_args(bn, training: bool):
    """bn = None (no BatchNorm in front) or (part, gamma, beta, eps, momentum, running_mean, running_var, num_batches)."""
    if bn is None:
        return 0, None, None, None, 0.0, 0.0, None, None, None
    part, gamma, beta, eps, mom, rm, rv, nbt = bn
    return (1 if training else 2), part, gamma, beta, float(eps), float(mom or 0.0), rm, rv, nbt


def tail_fwd_hidden(x, W, b, relu_in=False, bn=None, training=True, want_part=True):
    """One hidden layer of the tail: a = relu(in' W^T + b); returns (a, part, y, save_mean, save_rstd) -- y = in' = BatchNorm(x)
    and the saved statistics when ``bn`` is given (training), part = the statistics partials of a for the next call."""
    x = _need(x, torch.float32, "x")
    M, K = x.shape
    N = W.shape[0]
    dev = x.device
    a = torch.empty(M, N, dtype=torch.float32, device=dev)
    part = _tail_part(M, 3, N, dev) if want_part else None
    mode, bpart, gamma, beta, eps, mom, rm, rv, nbt = _tail_bn_args(bn, training)
    y = torch.empty_like(x) if (mode and training) else None
    sm = torch.empty(K, dtype=torch.float32, device=dev) if mode == 1 else None
    sr = torch.empty(K, dtype=torch.float32, device=dev) if mode == 1 else None
    call("b200med_tail_fwd_hidden", _ptr(x), M, K, int(bool(relu_in)), mode, _ptr(bpart), _ptr(gamma), _ptr(beta), eps, mom,
         _ptr(rm), _ptr(rv), _ptr(nbt), _ptr(sm), _ptr(sr), _ptr(y), _ptr(_need(W, torch.float32, "W")), _ptr(b), N, _ptr(a),
         _ptr(part), _stream())
    return a, part, y, sm, sr


def tail_fwd_out(x, W, b, relu_in=False, bn=None, training=True):
    """The output layer: out = in' W^T + b (C <= 8 columns); returns (out, y, save_mean, save_rstd)."""
    x = _need(x, torch.float32, "x")
    M, K = x.shape
    Cn = W.shape[0]
    dev = x.device
    out = torch.empty(M, Cn, dtype=torch.float32, device=dev)
    mode, bpart, gamma, beta, eps, mom, rm, rv, nbt = _tail_bn_args(bn, training)
    y = torch.empty_like(x) if (mode and training) else None
    sm = torch.empty(K, dtype=torch.float32, device=dev) if mode == 1 else None
    sr = torch.empty(K, dtype=torch.float32, device=dev) if mode == 1 else None
    call("b200med_tail_fwd_out", _ptr(x), M, K, int(bool(relu_in)), mode, _ptr(bpart), _ptr(gamma), _ptr(beta), eps, mom,
         _ptr(rm), _ptr(rv), _ptr(nbt), _ptr(sm), _ptr(sr), _ptr(y), _ptr(_need(W, torch.float32, "W")), _ptr(b), Cn, _ptr(out),
         _stream())
    return out, y, sm, sr


def tail_bwd_out(g, Wl, a, save_mean, save_rstd):
    """Partials (sum g2, sum g2 xhat) of the last BatchNorm's backward, g2 = g Wl."""
    g = _need(g, torch.float32, "g"); a = _need(a, torch.float32, "a")
    M, N = a.shape
    part = _tail_part(M, 2, N, a.device)
    call("b200med_tail_bwd_out", _ptr(g), g.shape[1], _ptr(_need(Wl, torch.float32, "Wl")), _ptr(a), M, N, _ptr(save_mean),
         _ptr(save_rstd), _ptr(part), _stream())
    return part


def tail_bwd_hidden(dy, g, Wl, a, part, gamma, save_mean, save_rstd, W, prev=None, relu_mask=None):
    """Backward of one hidden layer (see b200med_tail_bwd_hidden): returns (dz, dx, dgamma, dbeta, part_prev).
    dy = None for the last hidden layer (then g, Wl).  prev = (a_prev, mean_prev, rstd_prev) of the BatchNorm in front."""
    a = _need(a, torch.float32, "a")
    M, K = a.shape
    N = W.shape[1]
    dev = a.device
    dz = torch.empty(M, K, dtype=torch.float32, device=dev)
    dx = torch.empty(M, N, dtype=torch.float32, device=dev)
    dgamma = torch.empty(K, dtype=torch.float32, device=dev)
    dbeta = torch.empty(K, dtype=torch.float32, device=dev)
    part_prev = _tail_part(M, 2, N, dev) if prev is not None else None
    a_prev, mean_prev, rstd_prev = prev if prev is not None else (None, None, None)
    call("b200med_tail_bwd_hidden", _ptr(dy), _ptr(g), 0 if g is None else g.shape[1], _ptr(Wl), _ptr(a), M, K, _ptr(part),
         _ptr(gamma), _ptr(save_mean), _ptr(save_rstd), _ptr(dz), _ptr(dgamma), _ptr(dbeta), _ptr(_need(W, torch.float32, "W")),
         N, _ptr(dx), _ptr(a_prev), _ptr(mean_prev), _ptr(rstd_prev), _ptr(part_prev), _ptr(relu_mask), _stream())
    return dz, dx, dgamma, dbeta, part_prev


def pool_drop_fwd(z, B, L, Lc, Cn, drop_p=0.0, seed_dev=None, drop_base=0):
    p = torch.empty(B * (Lc // 2), Cn, dtype=torch.float32, device=z.device)
    call("b200med_pool_drop_fwd", _ptr(_need(z, torch.float32, "z")), _ptr(p), B, L, Lc, Cn, float(drop_p),
         _ptr(seed_dev), C.c_uint64(int(drop_base)), _stream())
    return p


def pool_drop_bwd(dp, z, dz, B, L, Lc, Cn, drop_p=0.0, seed_dev=None, drop_base=0):
    call("b200med_pool_drop_bwd", _ptr(_need(dp, torch.float32, "dp")), _ptr(z), _ptr(dz), B, L, Lc, Cn, float(drop_p),
         _ptr(seed_dev), C.c_uint64(int(drop_base)), _stream())
    return dz


def conv_pack(w, want_fwd=True, want_bwd=True):
    w = _need(w, torch.float32, "w")
    Cout, Cin, k = w.shape
    if k != 3:
        raise ValueError("the native convolution is built for kernel_size = 3")
    fwd = torch.empty(Cout, 3 * Cin, dtype=torch.float32, device=w.device) if want_fwd else None
    bwd = torch.empty(Cin, 3 * Cout, dtype=torch.float32, device=w.device) if want_bwd else None
    call("b200med_conv_pack", _ptr(w), _ptr(fwd), _ptr(bwd), Cout, Cin, _stream())
    return fwd, bwd


def conv_unpack_grad(dfwd, Cout, Cin):
    dw = torch.empty(Cout, Cin, 3, dtype=torch.float32, device=dfwd.device)
    call("b200med_conv_unpack_grad", _ptr(_need(dfwd, torch.float32, "dfwd")), _ptr(dw), Cout, Cin, _stream())
    return dw


def concat2(a, b):
    """[M, Ca] | [M, Cb] -> [M, Ca + Cb] (fp32)."""
    a = _need(a, torch.float32, "a"); b = _need(b, torch.float32, "b")
    M, Ca = a.shape
    Cb = b.shape[1]
    out = torch.empty(M, Ca + Cb, dtype=torch.float32, device=a.device)
    call("b200med_concat2", _ptr(a), _ptr(b), _ptr(out), M, Ca, Cb, _stream())
    return out


def slice_cols(x, col0, Cn):
    x = _need(x, torch.float32, "x")
    M, ld = x.shape
    out = torch.empty(M, Cn, dtype=torch.float32, device=x.device)
    call("b200med_slice_cols", _ptr(x), _ptr(out), M, ld, col0, Cn, _stream())
    return out


def take_rows(src, idx, out=None):
    """out[i] = src[idx[i]] for a 4-byte dtype (f32 / i32) and int64 indices."""
    if src.element_size() != 4:
        raise TypeError("take_rows serves 4-byte element types")
    src = _need(src, None, "src")
    idx = _need(idx, torch.int64, "idx")
    Cn = 1 if src.dim() == 1 else int(src[0].numel())
    if out is None:
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    call("b200med_take_rows", _ptr(src), _ptr(idx), _ptr(out), idx.numel(), Cn, _stream())
    return out


def transpose_last2(x, B, R, Cn):
    """[B, R, C] -> [B, C, R]"""
    y = torch.empty(B, Cn, R, dtype=torch.float32, device=x.device)
    call("b200med_transpose_last2", _ptr(_need(x, torch.float32, "x")), _ptr(y), B, R, Cn, _stream())
    return y


# ------------------------------------------------------------------------------------------- K2 bf16 (tcgen05)
def has_tcgen05() -> bool:
    return bool(_lib.load().b200med_has_tcgen05())


def gemm_bf16(A, B, M, N, K, a_kmajor=True, b_kmajor=True, bias=None, mask=None, relu=False,
              out_dtype=torch.bfloat16, split_k=1, out=None, rbi=False):
    """D[M,N] = A[M,K] B[N,K]^T on tcgen05.  K-major operand = stored [rows, K]; MN-major = stored [K, rows].
    rbi=True: D is written row-block-interleaved ([M/32][N/V][32][V], see b200med.h) -- M must be a multiple of 32."""
    A = _need(A, torch.bfloat16, "A"); Bm = _need(B, torch.bfloat16, "B")
    lda, ldb = A.shape[-1], Bm.shape[-1]
    D = out if out is not None else torch.empty(M, N, dtype=out_dtype, device=A.device)
    ws = None
    if split_k > 1:
        ws = workspace(_lib.load().b200med_gemm_bf16_ws_bytes(M, N, K, split_k), A.device, "splitk")
    call("b200med_gemm_bf16", _ptr(A), _ptr(Bm), _ptr(D), _ptr(bias), _ptr(mask), M, N, K, lda, ldb, D.shape[-1],
         int(a_kmajor), int(b_kmajor), _dt(D), int(relu), split_k, int(bool(rbi)), _ptr(ws), _stream())
    return D


def gemm_split_k(M: int, N: int, K: int, b_kmajor: bool = False) -> int:
    """split_k of a weight-gradient GEMM that fills the SMs this thread may use in whole waves (csrc/gemm_tcgen05.cu)."""
    return int(_lib.load().b200med_gemm_bf16_pick_split(M, N, K, int(b_kmajor)))


def colsum(dy: torch.Tensor) -> torch.Tensor:
    dy = _need(dy, None, "dy")
    M, N = dy.shape
    db = torch.empty(N, dtype=torch.float32, device=dy.device)
    ws = workspace(_lib.load().b200med_colsum_ws_bytes(M, N), dy.device, "colsum")
    call("b200med_colsum", _ptr(dy), _dt(dy), _ptr(db), M, N, N, _ptr(ws), _stream())
    return db


def to_bf16(x: torch.Tensor, out=None, relu: bool = False) -> torch.Tensor:
    x = _need(x, torch.float32, "x")
    y = out if out is not None else torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    call("b200med_relu_cast_f32_to_bf16" if relu else "b200med_cast_f32_to_bf16", _ptr(x), _ptr(y), x.numel(), _stream())
    return y


def split_bf16x3(x: torch.Tensor, row_order=None, stack_order=None, relu: bool = False):
    """(row3 [R, 3C], stack3 [3R, C]) split-bf16 copies of x [R, C] f32 (b200med_split_bf16x3); order 0 = left operand
    (hi, lo, hi), 1 = right operand (hi, hi, lo), None = that layout is not produced."""
    x = _need(x, torch.float32, "x")
    R, Cn = x.shape
    row3 = torch.empty(R, 3 * Cn, dtype=torch.bfloat16, device=x.device) if row_order is not None else None
    stack3 = torch.empty(3 * R, Cn, dtype=torch.bfloat16, device=x.device) if stack_order is not None else None
    call("b200med_split_bf16x3", _ptr(x), _ptr(row3), _ptr(stack3), R, Cn, int(row_order or 0), int(stack_order or 0), int(bool(relu)),
         _stream())
    return row3, stack3


def split_bf16x6(x: torch.Tensor, role: int, want_row: bool = True, want_stack: bool = False, relu: bool = False):
    """(row6 [R, 6C], stack6 [6R, C]) of the exact three-way bf16 split of x [R, C] f32 (b200med_split_bf16x6); role 0 = left
    operand, 1 = right operand of the product."""
    x = _need(x, torch.float32, "x")
    R, Cn = x.shape
    row6 = torch.empty(R, 6 * Cn, dtype=torch.bfloat16, device=x.device) if want_row else None
    stack6 = torch.empty(6 * R, Cn, dtype=torch.bfloat16, device=x.device) if want_stack else None
    call("b200med_split_bf16x6", _ptr(x), _ptr(row6), _ptr(stack6), R, Cn, int(role), int(bool(relu)), _stream())
    return row6, stack6


# The fp32 mode's LARGE products run on the bf16 tensor cores as six products of exactly split operands (fp32-like: 1.4e-6
# against fp64 at K = 2048, scripts/split6_accuracy.py); below this many flop the fp32 FMA kernels of csrc/gemm_f32.cu are as
# fast as split + GEMM.  0 switches the route off (B200MED_FP32_TC=0).
FP32_TC_MIN_FLOP = 0.0 if os.environ.get("B200MED_FP32_TC", "1") == "0" else 1.0e9


FP32_TC_KB_PER_SPLIT = 12      # k-blocks (of 64) one accumulator takes before it is rounded into the fp32 partial sum


def _fp32_tc_split(k6: int) -> int:
    """split_k of a six-product GEMM over a reduction of k6: the tensor core TRUNCATES when it adds into its fp32 accumulator,
    so a long reduction drifts towards zero (1.4e-6 at 6 x 2048; the fp32 FMA chain 5.7e-7); partial sums of <= 12 k-blocks
    added in fp32 by the fixed-order split-K reduction bring it to 4e-7 (scripts/split6_accuracy.py)."""
    nkb = (k6 + 63) // 64
    return max(1, min(256, (nkb + FP32_TC_KB_PER_SPLIT - 1) // FP32_TC_KB_PER_SPLIT))


def _fp32_tc(M: int, N: int, K: int, *lds: int) -> bool:
    return (FP32_TC_MIN_FLOP > 0 and 2.0 * M * N * K >= FP32_TC_MIN_FLOP and all(ld % 8 == 0 for ld in lds)
            and torch.cuda.is_available() and has_tcgen05())


def to_f32(x: torch.Tensor) -> torch.Tensor:
    x = _need(x, torch.bfloat16, "x")
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call("b200med_cast_bf16_to_f32", _ptr(x), _ptr(y), x.numel(), _stream())
    return y


# ------------------------------------------------------------------------------------------- K3
def _loss_ws(device):
    return workspace(_lib.load().b200med_loss_ws_bytes(0), device, "loss", zero=True)


def bce_logits(logits, labels, pos_weight=1.0, grad_scale=1.0, want_grad=True, want_probs=False, want_preds=True,
               counts=None, accumulate=False, out=None):
    """Fused BCE-with-logits: returns dict(loss[1], dlogits, probs, preds, counts[4]=(tn,fp,fn,tp)).  ``out``: optional dict of
    preallocated destination tensors (loss / probs / preds / counts) -- a captured train step lets the kernel write its
    persistent result buffers directly."""
    logits = _need(logits.reshape(-1), torch.float32, "logits")
    labels = _need(labels.reshape(-1), torch.float32, "labels")
    B, dev = logits.numel(), logits.device
    out = out or {}
    loss = out.get("loss") if out.get("loss") is not None else torch.empty(1, dtype=torch.float32, device=dev)
    dl = torch.empty(B, dtype=torch.float32, device=dev) if want_grad else None
    pr = (out.get("probs") if out.get("probs") is not None else torch.empty(B, dtype=torch.float32, device=dev)) if want_probs else None
    pd = (out.get("preds") if out.get("preds") is not None else torch.empty(B, dtype=torch.float32, device=dev)) if want_preds else None
    if counts is None:
        counts = out.get("counts")
    if counts is None:
        counts = torch.zeros(4, dtype=torch.int64, device=dev)
    call("b200med_bce_logits", _ptr(logits), _ptr(labels), B, float(pos_weight), float(grad_scale), _ptr(loss), _ptr(dl),
         _ptr(pr), _ptr(pd), _ptr(counts), int(accumulate), _ptr(_loss_ws(dev)), _stream())
    return dict(loss=loss, dlogits=dl, probs=pr, preds=pd, counts=counts)


def ce_logits(logits, target, class_weight=None, mask=None, target_shift=0, reduction=0, grad_scale=1.0,
              want_grad=True, want_probs=False, pred_shift=0, pred_mask_mode=0, cm=None, cm_classes=None,
              accumulate=False):
    logits = _need(logits, torch.float32, "logits")
    target = _need(target.reshape(-1), torch.int32, "target")
    B, Cn = logits.shape
    dev = logits.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dl = torch.empty(B, Cn, dtype=torch.float32, device=dev) if want_grad else None
    pr = torch.empty(B, Cn, dtype=torch.float32, device=dev) if want_probs else None
    pd = torch.empty(B, dtype=torch.int32, device=dev)
    cm_classes = cm_classes or Cn
    if cm is None:
        cm = torch.zeros(cm_classes, cm_classes, dtype=torch.int64, device=dev)
    call("b200med_ce_logits", _ptr(logits), _ptr(target), _ptr(class_weight), _ptr(mask), B, Cn, target_shift, reduction,
         float(grad_scale), _ptr(loss), _ptr(dl), _ptr(pr), _ptr(pd), pred_shift, pred_mask_mode, _ptr(cm), cm_classes,
         int(accumulate), _ptr(_loss_ws(dev)), _stream())
    return dict(loss=loss, dlogits=dl, probs=pr, preds=pd, cm=cm)


def ce_frame(logits, e, grad_scale=1.0, want_grad=True, counts=None, accumulate=False):
    """logits [stages, 1, 2, T] (or [stages, 2, T]); e [T] soft error labels."""
    logits = _need(logits, torch.float32, "logits")
    stages, T = logits.shape[0], logits.shape[-1]
    if logits.numel() != stages * 2 * T:
        raise ValueError("frame loss expects 2 classes and batch size 1")
    e = _need(e.reshape(-1), torch.float32, "e")
    dev = logits.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dl = torch.empty_like(logits) if want_grad else None
    pd = torch.empty(T, dtype=torch.float32, device=dev)
    if counts is None:
        counts = torch.zeros(4, dtype=torch.int64, device=dev)
    call("b200med_ce_frame", _ptr(logits), _ptr(e), stages, T, float(grad_scale), _ptr(loss), _ptr(dl), _ptr(pd),
         _ptr(counts), int(accumulate), _ptr(_loss_ws(dev)), _stream())
    return dict(loss=loss, dlogits=dl, preds=pd, counts=counts)


# ------------------------------------------------------------------------------------------- optimiser
def adam_advance(state: torch.Tensor, beta1: float, beta2: float):
    call("b200med_adam_advance", _ptr(_need(state, torch.float32, "state")), beta1, beta2, _stream())


def multi_copy(dst, src, adam=None):
    """dst[i] <- src[i] for lists of equally sized contiguous tensors in one launch per 48 pairs (b200med_multi_copy_f32);
    src f32, dst f32 or bf16 (all alike).  adam = (state, beta1, beta2): the launch also advances the Adam step scalars."""
    k = len(dst)
    if k == 0 and adam is None:
        return
    for d, s_ in zip(dst, src):
        _need(s_, torch.float32, "src"); _need(d, None, "dst")
        if d.numel() != s_.numel() or d.dtype != dst[0].dtype:
            raise ValueError("b200med multi_copy: sizes / dtypes of the pairs do not match")
    dt = _dt(dst[0]) if k else F32
    if dt not in (F32, BF16):
        raise TypeError("b200med multi_copy: dst must be float32 or bfloat16")
    srcs = (C.c_void_p * max(k, 1))(*[s_.data_ptr() for s_ in src])
    dsts = (C.c_void_p * max(k, 1))(*[d.data_ptr() for d in dst])
    ns = (C.c_int64 * max(k, 1))(*[d.numel() for d in dst])
    state, b1, b2 = adam if adam is not None else (None, 0.0, 0.0)
    call("b200med_multi_copy_f32", srcs, dsts, ns, k, dt, _ptr(state), float(b1), float(b2), _stream())


def adam_step(p, g, m, v, state, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    call("b200med_adam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(state), beta1, beta2, eps, weight_decay,
         grad_scale, _stream())


# ------------------------------------------------------------------------------------------- post-processing
def window_vote(frame_preds, starts, W: int, binary: bool = True):
    fp = _need(frame_preds.reshape(-1), torch.float32, "frame_preds")
    starts = _need(starts, torch.int32, "starts")
    out = torch.empty(starts.numel(), dtype=torch.float32, device=fp.device)
    call("b200med_window_vote", _ptr(fp), _ptr(starts), starts.numel(), W, int(binary), _ptr(out), _stream())
    return out


def soft_vote(pa, pb, labels=None):
    pa = _need(pa.reshape(-1), torch.float32, "pa"); pb = _need(pb.reshape(-1), torch.float32, "pb")
    n = pa.numel()
    preds = torch.empty(n, dtype=torch.float32, device=pa.device)
    counts = torch.zeros(4, dtype=torch.int64, device=pa.device)
    lab = None if labels is None else _need(labels.reshape(-1), torch.float32, "labels")
    call("b200med_soft_vote", _ptr(pa), _ptr(pb), _ptr(lab), n, _ptr(preds), _ptr(counts), 0, _ptr(None), _stream())
    return preds, counts


def cascade(binary, multiclass):
    b = _need(binary.reshape(-1), torch.int32, "binary"); m = _need(multiclass.reshape(-1), torch.int32, "multiclass")
    out = torch.empty_like(m)
    call("b200med_cascade", _ptr(b), _ptr(m), b.numel(), _ptr(out), _stream())
    return out


def confusion(target, pred, n_classes: int, cm=None, accumulate=False):
    t = _need(target.reshape(-1), torch.int32, "target"); p = _need(pred.reshape(-1), torch.int32, "pred")
    if cm is None:
        cm = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=t.device)
    call("b200med_confusion", _ptr(t), _ptr(p), t.numel(), n_classes, _ptr(cm), int(accumulate), _stream())
    return cm


def roc_auc(scores, labels):
    """Area under the ROC curve on the device -> (auc f64[1], stats i64[4] = n_pos, n_neg, 2*less+equal, 0)."""
    sc = _need(scores.reshape(-1), torch.float32, "scores"); lb = _need(labels.reshape(-1), torch.float32, "labels")
    if sc.numel() != lb.numel():
        raise ValueError("scores and labels differ in length")
    n, dev = sc.numel(), sc.device
    auc = torch.empty(1, dtype=torch.float64, device=dev)
    stats = torch.empty(4, dtype=torch.int64, device=dev)
    ws = workspace(_lib.load().b200med_roc_auc_ws_bytes(n), dev, "roc_auc")
    call("b200med_roc_auc", _ptr(sc), _ptr(lb), n, _ptr(auc), _ptr(stats), _ptr(ws), _stream())
    return auc, stats


# ------------------------------------------------------------------------------------------- TeCNo frame head
TCN_PACK_FLOATS, TCN_GRAD_FLOATS, TCN_MAPS = 32896, 16512, 64


def _geom(tloc, trem):
    return (_ptr(None if tloc is None else _need(tloc, torch.int32, "tloc")),
            _ptr(None if trem is None else _need(trem, torch.int32, "trem")))


def tcn_slots(T: int) -> int:
    return int(_lib.load().b200med_tcn_slots(int(T)))


def tcn_pack(ptr_table: torch.Tensor, n_layers: int, out=None) -> torch.Tensor:
    """ptr_table: int64 CUDA tensor [n_layers, 4] of parameter addresses (see b200med_tcn_pack)."""
    ptr_table = _need(ptr_table, torch.int64, "ptr_table")
    if ptr_table.numel() != n_layers * 4:
        raise ValueError("ptr_table must hold 4 addresses per layer")
    if out is None:
        out = torch.empty(n_layers, TCN_PACK_FLOATS, dtype=torch.float32, device=ptr_table.device)
    call("b200med_tcn_pack", _ptr(ptr_table), n_layers, _ptr(out), _stream())
    return out


def _seed_dev(seed_dev):
    return _ptr(None if seed_dev is None else _need(seed_dev, torch.int64, "seed_dev"))


def tcn_layer_fwd(x, pack, out, y_save, dilation: int, causal: bool, drop_p=0.0, seed=0, drop_base=0, tloc=None, trem=None,
                  seed_dev=None):
    x = _need(x, torch.float32, "x")
    T = x.shape[0]
    tl, tr = _geom(tloc, trem)
    call("b200med_tcn_layer_fwd", _ptr(x), _ptr(pack), _ptr(out), _ptr(y_save), T, int(dilation), int(bool(causal)), tl, tr,
         float(drop_p), C.c_uint64(int(seed)), _seed_dev(seed_dev), C.c_uint64(int(drop_base)), _stream())
    return out


def tcn_layer_bwd_hidden(dout, x, y, pack, dpre, partials, n_slots, dilation, causal, drop_p=0.0, seed=0, drop_base=0,
                         tloc=None, trem=None, seed_dev=None):
    dout = _need(dout, torch.float32, "dout")
    tl, tr = _geom(tloc, trem)
    call("b200med_tcn_layer_bwd_hidden", _ptr(dout), _ptr(x), _ptr(y), _ptr(pack), _ptr(dpre), _ptr(partials), int(n_slots),
         dout.shape[0], int(dilation), int(bool(causal)), tl, tr, float(drop_p), C.c_uint64(int(seed)), _seed_dev(seed_dev),
         C.c_uint64(int(drop_base)), _stream())
    return dpre


def tcn_layer_bwd_input(dpre, dout, pack, dx, dilation, causal, tloc=None, trem=None):
    dpre = _need(dpre, torch.float32, "dpre")
    tl, tr = _geom(tloc, trem)
    call("b200med_tcn_layer_bwd_input", _ptr(dpre), _ptr(dout), _ptr(pack), _ptr(dx), dpre.shape[0], int(dilation),
         int(bool(causal)), tl, tr, _stream())
    return dx


def tcn_reduce_grads(partials, n_layers: int, n_slots: int) -> torch.Tensor:
    grads = torch.empty(n_layers, TCN_GRAD_FLOATS, dtype=torch.float32, device=partials.device)
    call("b200med_tcn_reduce_grads", _ptr(partials), n_layers, n_slots, _ptr(grads), _stream())
    return grads


def tcn_out_fwd(x, w, b) -> torch.Tensor:
    x = _need(x, torch.float32, "x"); w = _need(w, torch.float32, "w"); b = _need(b, torch.float32, "b")
    T, Cn = x.shape[0], w.shape[0]
    logits = torch.empty(Cn, T, dtype=torch.float32, device=x.device)
    call("b200med_tcn_out_fwd", _ptr(x), _ptr(w), _ptr(b), _ptr(logits), T, Cn, _stream())
    return logits


def tcn_out_bwd(dlogits, w):
    dl = _need(dlogits, torch.float32, "dlogits"); w = _need(w, torch.float32, "w")
    Cn, T = dl.shape
    dx = torch.empty(T, TCN_MAPS, dtype=torch.float32, device=dl.device)
    dl_t = torch.empty(T, Cn, dtype=torch.float32, device=dl.device)
    call("b200med_tcn_out_bwd", _ptr(dl), _ptr(w), _ptr(dx), _ptr(dl_t), T, Cn, _stream())
    return dx, dl_t


def tcn_softmax_fwd(logits) -> torch.Tensor:
    logits = _need(logits, torch.float32, "logits")
    Cn, T = logits.shape
    p = torch.empty(T, Cn, dtype=torch.float32, device=logits.device)
    call("b200med_tcn_softmax_fwd", _ptr(logits), _ptr(p), T, Cn, _stream())
    return p


def tcn_softmax_bwd(p, dp) -> torch.Tensor:
    p = _need(p, torch.float32, "p"); dp = _need(dp, torch.float32, "dp")
    T, Cn = p.shape
    dl = torch.empty(Cn, T, dtype=torch.float32, device=p.device)
    call("b200med_tcn_softmax_bwd", _ptr(p), _ptr(dp), _ptr(dl), T, Cn, _stream())
    return dl


def _f32_array(values):
    return None if values is None else (C.c_float * len(values))(*[float(v) for v in values])


def tcn_stage_fwd(x, softmax_in, in_w, in_b, ptr_table, n_layers, out_w, out_b, causal, drop_p=None, seed=0, layer_base=0,
                  keep=True, tloc=None, trem=None, seed_dev=None):
    """Whole SingleStageModel forward in one C call -> dict(logits [C,T], xin, acts, ys, pack)."""
    x = _need(x, torch.float32, "x")
    in_w = _need(in_w, torch.float32, "in_w"); out_w = _need(out_w, torch.float32, "out_w")
    n_cls, in_dim = out_w.shape[0], in_w.shape[1]
    T = x.shape[1] if softmax_in else x.shape[0]
    if (x.shape[0] if softmax_in else x.shape[1]) != in_dim:
        raise ValueError(f"stage input has {x.shape[0] if softmax_in else x.shape[1]} features, conv_1x1 expects {in_dim}")
    dev = x.device
    p_in = torch.empty(T, n_cls, dtype=torch.float32, device=dev) if softmax_in else None
    acts = torch.empty((n_layers + 1) if keep else 2, T, TCN_MAPS, dtype=torch.float32, device=dev)
    ys = torch.empty(n_layers, T, TCN_MAPS, dtype=torch.float32, device=dev) if keep else None
    pack = torch.empty(n_layers, TCN_PACK_FLOATS, dtype=torch.float32, device=dev)
    logits = torch.empty(n_cls, T, dtype=torch.float32, device=dev)
    tl, tr = _geom(tloc, trem)
    if T:
        call("b200med_tcn_stage_fwd", _ptr(x), in_dim, int(bool(softmax_in)), _ptr(in_w), _ptr(_need(in_b, torch.float32, "in_b")),
             _ptr(_need(ptr_table, torch.int64, "ptr_table")), n_layers, _ptr(out_w), _ptr(_need(out_b, torch.float32, "out_b")),
             n_cls, T, int(bool(causal)), tl, tr, _f32_array(drop_p), C.c_uint64(int(seed)), _seed_dev(seed_dev),
             C.c_uint64(int(layer_base)), int(bool(keep)), _ptr(p_in), _ptr(acts), _ptr(ys), _ptr(pack), _ptr(logits), _stream())
    return dict(logits=logits, xin=p_in if softmax_in else x, acts=acts, ys=ys, pack=pack)


def tcn_stage_bwd(dlogits, xin, softmax_in, in_w, out_w, n_layers, causal, acts, ys, pack, want_dx, drop_p=None, seed=0,
                  layer_base=0, tloc=None, trem=None, seed_dev=None):
    """Whole stage backward in one C call -> dict(dx, d_in_w [64,in_dim], d_in_b, layer_grads [L,GRAD], d_out_w [C,64], d_out_b)."""
    dl = _need(dlogits, torch.float32, "dlogits")
    n_cls, T = dl.shape
    in_dim = in_w.shape[1]
    dev = dl.device
    d_in_w = torch.empty(TCN_MAPS, in_dim, dtype=torch.float32, device=dev)
    d_in_b = torch.empty(TCN_MAPS, dtype=torch.float32, device=dev)
    layer_grads = torch.empty(n_layers, TCN_GRAD_FLOATS, dtype=torch.float32, device=dev)
    d_out_w = torch.empty(n_cls, TCN_MAPS, dtype=torch.float32, device=dev)
    d_out_b = torch.empty(n_cls, dtype=torch.float32, device=dev)
    dx = None
    if want_dx:
        dx = torch.empty((n_cls, T) if softmax_in else (T, in_dim), dtype=torch.float32, device=dev)
    ws = workspace(_lib.load().b200med_tcn_stage_bwd_ws_bytes(T, in_dim, n_cls, n_layers), dev, "tcn_bwd")
    tl, tr = _geom(tloc, trem)
    call("b200med_tcn_stage_bwd", _ptr(dl), _ptr(_need(xin, torch.float32, "xin")), in_dim, int(bool(softmax_in)),
         _ptr(_need(in_w, torch.float32, "in_w")), _ptr(_need(out_w, torch.float32, "out_w")), n_cls, n_layers, T,
         int(bool(causal)), tl, tr, _f32_array(drop_p), C.c_uint64(int(seed)), _seed_dev(seed_dev),
         C.c_uint64(int(layer_base)), _ptr(acts), _ptr(ys), _ptr(pack), _ptr(ws), _ptr(d_in_w), _ptr(d_in_b), _ptr(layer_grads),
         _ptr(d_out_w), _ptr(d_out_b), _ptr(dx), _stream())
    return dict(dx=dx, d_in_w=d_in_w, d_in_b=d_in_b, layer_grads=layer_grads, d_out_w=d_out_w, d_out_b=d_out_b)


def tcn_stage_fwd_bf16(x, softmax_in, in_w, in_b, ptr_table, n_layers, out_w, out_b, causal, tloc=None, trem=None) -> torch.Tensor:
    """Inference-only stage forward on the bf16 tcgen05 layer kernel -> logits [C, T]."""
    x = _need(x, torch.float32, "x")
    in_w = _need(in_w, torch.float32, "in_w"); out_w = _need(out_w, torch.float32, "out_w")
    n_cls, in_dim = out_w.shape[0], in_w.shape[1]
    T = x.shape[1] if softmax_in else x.shape[0]
    dev = x.device
    logits = torch.empty(n_cls, T, dtype=torch.float32, device=dev)
    if T == 0:
        return logits
    p_in = torch.empty(T, n_cls, dtype=torch.float32, device=dev) if softmax_in else None
    res = torch.empty(2, T, TCN_MAPS, dtype=torch.float32, device=dev)
    opn = torch.empty(2, T, TCN_MAPS, dtype=torch.bfloat16, device=dev)
    pack = torch.empty(n_layers, TCN_PACK_FLOATS, dtype=torch.float32, device=dev)
    wb16 = torch.empty(n_layers, 4 * TCN_MAPS, TCN_MAPS, dtype=torch.bfloat16, device=dev)
    tl, tr = _geom(tloc, trem)
    call("b200med_tcn_stage_fwd_bf16", _ptr(x), in_dim, int(bool(softmax_in)), _ptr(in_w), _ptr(_need(in_b, torch.float32, "in_b")),
         _ptr(_need(ptr_table, torch.int64, "ptr_table")), n_layers, _ptr(out_w), _ptr(_need(out_b, torch.float32, "out_b")), n_cls, T,
         int(bool(causal)), tl, tr, _ptr(p_in), _ptr(res), _ptr(opn), _ptr(pack), _ptr(wb16), _ptr(logits), _stream())
    return logits
