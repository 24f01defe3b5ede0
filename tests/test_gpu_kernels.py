"""GPU parity tests of the individual kernels, called through the C ABI (ctypes), against the oracle
and the committed golden fixtures.  Bit-exact for integer / index / fp32-elementwise work; stated
tolerances for floating-point reductions."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
from multimodal_error_detection_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    from multimodal_error_detection_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def c_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "c", "libmed_oracle.so"))
    lib.med_oracle_window_starts.restype = ctypes.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dev(a, dtype=None):
    t = torch.as_tensor(a)
    return t.to("cuda", dtype) if dtype is not None else t.to("cuda")


# ------------------------------------------------------------------------------------------- K0
@pytest.mark.parametrize("case", cases.WINDOW_CASES, ids=[c[0] for c in cases.WINDOW_CASES])
def test_window_index_golden(case, ops, golden_dir):
    name, seed, nv, lo, hi, W, S = case
    gold = np.load(os.path.join(golden_dir, "window_index.npz"))
    g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
    r = ops.window_index(dev(g), dev(offsets), W, S, dev(e5))
    assert np.array_equal(r["starts"].cpu().numpy(), gold[f"{name}/starts"])
    assert np.array_equal(r["g_win"].cpu().numpy().reshape(-1, 1), gold[f"{name}/g_win"])
    assert np.array_equal(r["e5_win"].cpu().numpy(), gold[f"{name}/e_win"])
    subj = [synthetic.trial_name(i) for i in r["subj_win"].cpu().tolist()]
    assert subj == gold[f"{name}/subj_win"].tolist()


def test_window_index_large_vs_c_oracle(ops, c_oracle):
    """BASELINE-scale table (4096 videos, ~2.4 M frames): identical to the C oracle, plus the
    size-independent properties of the walk."""
    W, S = 16, 4
    g, e5, offsets = synthetic.label_tracks(11, 4096, 300, 900)
    r = ops.window_index(dev(g), dev(offsets), W, S)
    starts = r["starts"].cpu().numpy().astype(np.int64)
    want = np.zeros(len(g), dtype=np.int64)
    n = c_oracle.med_oracle_window_starts(_p(g), _p(offsets), ctypes.c_int64(len(offsets) - 1), ctypes.c_int64(W),
                                          ctypes.c_int64(S), _p(want), ctypes.c_int64(len(want)))
    assert n == len(starts) and np.array_equal(starts, want[:n])
    assert np.all(np.diff(starts) > 0)                                  # globally sorted
    assert np.all(g[starts] == g[starts + W - 1])                       # end points agree
    subj = np.searchsorted(offsets, starts, side="right") - 1
    assert np.all(starts + W < offsets[subj + 1])                       # strict bound: last legal window never emitted
    same = subj[1:] == subj[:-1]
    assert np.all(np.diff(starts)[same] >= min(S, 1))


def test_window_index_edge_cases(ops):
    # subject with no non-zero gesture -> IndexError like the reference
    g = np.zeros(50, dtype=np.float32)
    with pytest.raises(IndexError):
        ops.window_index(dev(g), dev(np.asarray([0, 50], dtype=np.int64)), 10, 6)
    # subjects shorter than a window, and n == W + 1 exactly (one window at most)
    g = np.ones(40, dtype=np.float32)
    off = np.asarray([0, 5, 16, 40], dtype=np.int64)     # lengths 5, 11, 24
    r = ops.window_index(dev(g), dev(off), 10, 6)
    assert r["starts"].cpu().tolist() == [5, 16, 22, 28]
    assert r["win_offsets"].cpu().tolist() == [0, 0, 1, 4]
    # empty table
    r = ops.window_index(dev(np.zeros(0, dtype=np.float32)), dev(np.asarray([0], dtype=np.int64)), 10, 6)
    assert r["starts"].numel() == 0
    # NaN gesture ids never match (Python float semantics of the reference)
    g = np.ones(30, dtype=np.float32); g[12] = np.nan
    r = ops.window_index(dev(g), dev(np.asarray([0, 30], dtype=np.int64)), 5, 5)
    s = r["starts"].cpu().numpy()
    assert not np.any((s == 12) | (s + 4 == 12))


@pytest.mark.parametrize("delete_nd", [True, False])
def test_powerset_golden(delete_nd, ops, golden_dir):
    gold = np.load(os.path.join(golden_dir, "powerset.npz"))
    e7, mask = ops.powerset(dev(cases.all_label_rows()), delete_nd)
    assert e7.dtype == torch.int32 and mask.dtype == torch.bool
    assert np.array_equal(e7.cpu().numpy(), gold[f"e7_{int(delete_nd)}"])
    assert np.array_equal(mask.cpu().numpy(), gold[f"mask_{int(delete_nd)}"])
    e7, mask = ops.powerset(dev(np.zeros((0, 5), dtype=np.float32)), delete_nd)
    assert e7.shape == (0, 7)


def test_powerset_large_vs_c_oracle(ops, c_oracle):
    _, e5, _ = synthetic.label_tracks(5, 2048, 300, 900)
    e7, mask = ops.powerset(dev(e5), True)
    want = np.zeros((len(e5), 7), dtype=np.int32); wm = np.zeros(len(e5), dtype=np.uint8)
    c_oracle.med_oracle_powerset(_p(e5), ctypes.c_int64(len(e5)), ctypes.c_int(1), _p(want), _p(wm))
    assert np.array_equal(e7.cpu().numpy(), want) and np.array_equal(mask.cpu().numpy(), wm.astype(bool))
    assert np.all(e7.cpu().numpy()[:, :6].sum(1) <= 1)


# ------------------------------------------------------------------------------------------- K1
def _oracle_gather(c_oracle, table, mean, std, starts, W):
    out = np.zeros((len(starts), W, table.shape[1]), dtype=np.float32)
    c_oracle.med_oracle_gather_norm(_p(table), ctypes.c_int64(table.shape[1]), _p(mean), _p(std),
                                    _p(starts.astype(np.int64)), ctypes.c_int64(len(starts)), ctypes.c_int64(W), _p(out))
    return out


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("W", [10, 16, 3])
def test_gather_norm_bit_exact(variant, W, ops, c_oracle):
    rng = np.random.Generator(np.random.PCG64(W))
    N, B = 700, 67
    image = np.maximum(rng.standard_normal((N, 2048), dtype=np.float32), 0)
    kin = rng.standard_normal((N, 26), dtype=np.float32)
    mi, si = rng.standard_normal(2048, dtype=np.float32), rng.random(2048, dtype=np.float32) + 0.25
    mk, sk = rng.standard_normal(26, dtype=np.float32), rng.random(26, dtype=np.float32) + 0.25
    starts = rng.integers(0, N - W, B).astype(np.int32)
    img_out = torch.empty(B, W, 2048, device="cuda")
    kin_out = torch.empty(B, W, 26, device="cuda")
    ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), img_out), ops.GatherStream(dev(kin), dev(mk), dev(sk), kin_out)],
                    dev(starts), W, variant)
    assert np.array_equal(img_out.cpu().numpy(), _oracle_gather(c_oracle, image, mi, si, starts, W))
    assert np.array_equal(kin_out.cpu().numpy(), _oracle_gather(c_oracle, kin, mk, sk, starts, W))


def test_gather_norm_concat_layout_and_bf16(ops, c_oracle):
    """Streams written side by side into one [B*W, 2048+26+6] matrix; bf16 image output = RN(fp32 result)."""
    rng = np.random.Generator(np.random.PCG64(5))
    N, B, W = 300, 33, 10
    image = rng.standard_normal((N, 2048), dtype=np.float32)
    kin = rng.standard_normal((N, 26), dtype=np.float32)
    mi, si = rng.standard_normal(2048, dtype=np.float32), rng.random(2048, dtype=np.float32) + 0.5
    mk, sk = rng.standard_normal(26, dtype=np.float32), rng.random(26, dtype=np.float32) + 0.5
    starts = rng.integers(0, N - W, B).astype(np.int32)
    cat = torch.full((B * W, 2080), -7.0, device="cuda")
    ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), cat, 0), ops.GatherStream(dev(kin), dev(mk), dev(sk), cat, 2048)],
                    dev(starts), W)
    got = cat.cpu().numpy().reshape(B, W, 2080)
    assert np.array_equal(got[:, :, :2048], _oracle_gather(c_oracle, image, mi, si, starts, W))
    assert np.array_equal(got[:, :, 2048:2074], _oracle_gather(c_oracle, kin, mk, sk, starts, W))
    assert np.all(got[:, :, 2074:] == -7.0)
    want32 = torch.from_numpy(_oracle_gather(c_oracle, image, mi, si, starts, W))
    for variant in (1, 2):
        out16 = torch.empty(B, W, 2048, device="cuda", dtype=torch.bfloat16)
        ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), out16, 0, exact_div=True)], dev(starts), W, variant)
        assert torch.equal(out16.cpu(), want32.to(torch.bfloat16))
        ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), out16, 0, exact_div=False)], dev(starts), W, variant)
        assert torch.allclose(out16.cpu().float(), want32, rtol=8e-3, atol=1e-6)     # 1 bf16 ulp


def test_gather_norm_per_step_stats_no_stats_and_extra_streams(ops):
    rng = np.random.Generator(np.random.PCG64(9))
    N, B, W = 200, 19, 8
    starts = rng.integers(0, N - W, B).astype(np.int32)
    rows = torch.from_numpy(starts.astype(np.int64))[:, None] + torch.arange(W)[None]
    for D in (2048, 512, 128, 26, 7):
        table = torch.from_numpy(rng.standard_normal((N, D), dtype=np.float32))
        mean = torch.from_numpy(rng.standard_normal((W, D), dtype=np.float32))
        std = torch.from_numpy(rng.random((W, D), dtype=np.float32) + 0.5)
        out = torch.empty(B, W, D, device="cuda")
        ops.gather_norm([ops.GatherStream(dev(table), dev(mean), dev(std), out)], dev(starts), W)
        assert torch.equal(out.cpu(), (table[rows] - mean) / std), D
        ops.gather_norm([ops.GatherStream(dev(table), None, None, out)], dev(starts), W)
        assert torch.equal(out.cpu(), table[rows]), D
    # empty batch is a no-op
    ops.gather_norm([ops.GatherStream(dev(table), None, None, torch.empty(0, W, 7, device="cuda"))],
                    torch.empty(0, dtype=torch.int32, device="cuda"), W)


def test_gather_norm_full_size_property(ops):
    """BASELINE config T at full batch (8192 windows x 16 x 2074): device-side identity check against
    plain indexing, and linearity in (mean, std)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    N, B, W = 200_000, 8192, 16
    image = torch.randn(N, 2048, device="cuda", generator=g)
    mean, std = torch.randn(2048, device="cuda", generator=g), torch.rand(2048, device="cuda", generator=g) + 0.5
    starts = torch.randint(0, N - W, (B,), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    out = torch.empty(B, W, 2048, device="cuda")
    for variant in (1, 2):
        out.zero_()
        ops.gather_norm([ops.GatherStream(image, mean, std, out)], starts, W, variant)
        rows = starts.long()[:, None] + torch.arange(W, device="cuda")[None]
        for lo in range(0, B, 1024):
            assert torch.equal(out[lo:lo + 1024], (image[rows[lo:lo + 1024]] - mean) / std)


@pytest.mark.parametrize("B,W,want_variant", [(8192, 16, 5), (512, 10, 4), (4096, 8, 5), (600, 7, 4)])
@pytest.mark.parametrize("exact", [True, False])
def test_gather_norm_benchmarked_instantiation_vs_oracle(B, W, want_variant, exact, ops, c_oracle):
    """The instantiation bench.py times -- bf16 output, variant 0 (auto) at B*W >= 4096: the TMA staging-ring kernel
    gather_norm_tma_kernel<bf16, exact, R, S, 256> (8-row stages when W % 8 == 0, else 2-row stages) -- against the C oracle
    at the headline size (8192 x 16 x 2048) and at the parity size (512 x 10).  exact division: bit-identical to
    RN_bf16(oracle fp32); reciprocal multiply: within one bf16 ulp.  The kinematics stream rides along in the same call
    (fp32, bit-exact) exactly as CustomWindowDataset.gather_batch issues it.  b200med_gather_last_variant() proves which
    device path ran."""
    rng = np.random.Generator(np.random.PCG64(B + W))
    N = 40_000
    image = np.maximum(rng.standard_normal((N, 2048), dtype=np.float32), 0)
    kin = rng.standard_normal((N, 26), dtype=np.float32)
    mi, si = rng.standard_normal(2048, dtype=np.float32) * 0.3, rng.random(2048, dtype=np.float32) + 0.25
    mk, sk = rng.standard_normal(26, dtype=np.float32), rng.random(26, dtype=np.float32) + 0.25
    starts = rng.integers(0, N - W, B).astype(np.int32)
    starts[:3] = [0, N - W, N - W - 1]                       # first and last legal rows
    img_out = torch.full((B, W, 2048), 7.0, device="cuda", dtype=torch.bfloat16)
    kin_out = torch.full((B, W, 26), 7.0, device="cuda")
    ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), img_out, 0, exact_div=exact),
                     ops.GatherStream(dev(kin), dev(mk), dev(sk), kin_out, 0, exact_div=True)], dev(starts), W, 0)
    assert ops.gather_last_variant() == want_variant, ops.gather_last_variant()
    want32 = torch.from_numpy(_oracle_gather(c_oracle, image, mi, si, starts, W))
    got = img_out.cpu()
    if exact:
        assert torch.equal(got, want32.to(torch.bfloat16))
    else:
        w16 = want32.to(torch.bfloat16)
        # one bf16 ulp: the two neighbours of RN_bf16(oracle) in the bf16 grid
        ulp = (w16.view(torch.int16).int() - got.view(torch.int16).int()).abs()
        assert int(ulp.max()) <= 1, int(ulp.max())
        assert float((ulp > 0).float().mean()) < 0.2
    assert np.array_equal(kin_out.cpu().numpy(), _oracle_gather(c_oracle, kin, mk, sk, starts, W))
    # SM cap (the prefetching launch of the train step): same bits on 56 SMs
    img2 = torch.empty_like(img_out)
    ops.gather_norm([ops.GatherStream(dev(image), dev(mi), dev(si), img2, 0, exact_div=exact)], dev(starts), W, 56 << 8)
    assert ops.gather_last_variant() == want_variant and torch.equal(img2, img_out)


@pytest.mark.parametrize("B,W,K", [(8192, 16, 2048), (100, 16, 2048), (33, 32, 2048), (7, 128, 2048), (1, 64, 2048),
                                   (300, 16, 64), (300, 16, 128), (300, 16, 192), (1000, 16, 256)])
def test_gather_linear_fused_vs_k1_then_gemm(B, W, K, ops):
    """b200med_gather_linear_bf16 (K1 fused into the FeatureExtractor's first layer): the bf16 batch it emits is BIT-IDENTICAL to
    the standalone K1 kernel's (which the tests above pin to the C oracle), and its output equals the tcgen05 GEMM on that batch
    (same operands, same fp32 accumulation: at most an ulp of bf16 apart) -- at the headline size, on ragged tile tails and for
    short reductions (K of one to four k-blocks: the operand rings then hold a whole tile and the tile hand-over changes)."""
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    g = torch.Generator(device="cuda").manual_seed(B + W + K)
    N = 30_000
    table = torch.randn(N, K, device="cuda", generator=g).clamp_min_(0)
    mean, std = torch.randn(K, device="cuda", generator=g) * 0.3, torch.rand(K, device="cuda", generator=g) + 0.25
    starts = torch.randint(0, N - W, (B,), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
    starts[0] = N - W                                     # last legal window
    w = (torch.randn(512, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(512, device="cuda", generator=g)
    xb, y = ops.gather_linear_bf16(table, mean, std, starts, W, w, bias, relu=True)
    img = torch.empty(B, W, K, device="cuda", dtype=torch.bfloat16)
    ops.gather_norm([ops.GatherStream(table, mean, std, img, 0, exact_div=False)], starts, W, 1)
    assert torch.equal(xb.view(B, W, K), img)
    y_ref = ops.gemm_bf16(img.view(B * W, K), w, B * W, 512, K, True, True, bias=bias, relu=True)
    ulp = (y.view(torch.int16).int() - y_ref.view(torch.int16).int()).abs()
    assert int(ulp.max()) <= 1 and float((ulp > 0).float().mean()) < 1e-2, (int(ulp.max()), float((ulp > 0).float().mean()))
    want = torch.relu(img.view(B * W, K)[:4096].double() @ w.double().T + bias.double())
    assert float((y[:4096].double() - want).abs().max() / want.abs().max()) < 2e-2
    # no statistics: plain bf16 copy of the rows
    xb0, _ = ops.gather_linear_bf16(table, None, None, starts, W, w, None, relu=False)
    rows = starts.long()[:, None] + torch.arange(W, device="cuda")[None]
    assert torch.equal(xb0.view(B, W, K), table[rows].to(torch.bfloat16))


# ------------------------------------------------------------------------------------------- K2 fp32
@pytest.mark.parametrize("shape", [(5120, 512, 2048), (1234, 256, 512), (77, 32, 256), (1, 6, 64), (130, 65, 33)])
def test_linear_f32(shape, ops):
    """Relative tolerance 1e-5 of the matrix max (north_star fp32 bar), against an fp64 CPU product."""
    Mr, N, K = shape
    g = torch.Generator().manual_seed(Mr)
    x, w, b = torch.randn(Mr, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    dy = torch.randn(Mr, N, generator=g)
    y = ops.linear_fwd_f32(dev(x), dev(w), dev(b), relu=True).cpu()
    y64 = torch.relu(x.double() @ w.double().T + b.double())
    assert (y - y64).abs().max() <= 1e-5 * y64.abs().max()
    m2 = (torch.randn(Mr, K, generator=g) > 0).float()
    dx = ops.linear_bwd_data_f32(dev(dy), dev(w), relu_out=dev(m2)).cpu()
    dx64 = (dy.double() @ w.double()) * m2.double()
    assert (dx - dx64).abs().max() <= 1e-5 * dx64.abs().max() + 1e-12
    dw, db = ops.linear_bwd_weight_f32(dev(dy), dev(x))
    dw64, db64 = dy.double().T @ x.double(), dy.double().sum(0)
    assert (dw.cpu() - dw64).abs().max() <= 1e-5 * dw64.abs().max()
    assert (db.cpu() - db64).abs().max() <= 1e-5 * db64.abs().max() + 1e-6
    # determinism: fixed reduction order -> identical bits run to run
    dw2, db2 = ops.linear_bwd_weight_f32(dev(dy), dev(x))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


# ------------------------------------------------------------------------------------------- K2 bf16 tcgen05
GEMM_CASES = [
    # M, N, K, a_kmajor, b_kmajor, split_k      (what it stands for)
    (5120, 512, 2048, True, True, 1),      # FE layer 1 forward
    (5120, 256, 512, True, True, 1),       # layer 2 forward
    (5120, 32, 256, True, True, 1),        # layer 3 forward (BLOCK_N = 32)
    (512, 2048, 5120, False, False, 4),    # layer 1 weight gradient (both MN-major, split-K)
    (256, 512, 5120, False, False, 8),     # layer 2 weight gradient
    (32, 256, 5120, False, False, 8),      # layer 3 weight gradient (M tile mostly padding)
    (5120, 256, 32, True, False, 1),       # layer 3 data gradient (K = 32 < BLOCK_K)
    (5120, 512, 256, True, False, 1),      # layer 2 data gradient
    (300, 200, 136, True, True, 1),        # ragged M / N / K tails
    (136, 72, 200, False, True, 1),        # MN-major A, K-major B
    (1000, 128, 1000, True, False, 3),     # K tail under split-K
    (512, 192, 8192, False, False, 8),     # LSTM layer-0 weight gradient: N = 192 on the 256-column CTA-pair tile (B rows >= 192 zero-filled by TMA)
    (512, 192, 1024, True, True, 1),       # the same width with K-major operands, one split (pair tile, row-major / staged epilogues)
    (384, 448, 2048, False, True, 2),      # a last pair tile three quarters full behind a whole one, ragged M tile
    (40000, 512, 256, True, False, 1),     # layer 2 data gradient, several tiles per CTA (double-buffered staging: 626 tiles on 148 CTAs)
    (40000, 256, 32, True, False, 1),      # layer 3 data gradient, several tiles per CTA
    (37000, 304, 200, True, True, 1),      # the same flow with ragged M / N / K (N = 304: a 48-column last tile, rows TMA can address)
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=[f"{c[0]}x{c[1]}x{c[2]}_{'k' if c[3] else 'm'}{'k' if c[4] else 'm'}_s{c[5]}" for c in GEMM_CASES])
def test_gemm_bf16_tcgen05(case, ops):
    """bf16 operands, fp32 accumulate: compare with the same bf16-rounded operands multiplied in fp64.
    Tolerance 2e-2 relative to the matrix max is the north_star bar; the fp32-output case is held to 1e-3."""
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    Mr, N, K, ak, bk, split = case
    g = torch.Generator().manual_seed(Mr + N + K)
    A = (torch.randn(Mr, K, generator=g) / K ** 0.25).to(torch.bfloat16)
    B = (torch.randn(N, K, generator=g) / K ** 0.25).to(torch.bfloat16)
    want = A.double() @ B.double().T
    Ad = dev(A if ak else A.T.contiguous())
    Bd = dev(B if bk else B.T.contiguous())
    out = ops.gemm_bf16(Ad, Bd, Mr, N, K, ak, bk, out_dtype=torch.float32, split_k=split).cpu().double()
    err = (out - want).abs().max() / want.abs().max()
    assert err < 1e-3, f"fp32-out rel err {err:.3e}"
    bias = torch.randn(N, generator=g)
    mask = (torch.randn(Mr, N, generator=g) > 0).to(torch.bfloat16)
    out = ops.gemm_bf16(Ad, Bd, Mr, N, K, ak, bk, bias=dev(bias), relu=True, mask=dev(mask), split_k=split).cpu().double()
    want2 = torch.relu(want + bias.double()) * mask.double()
    err = (out - want2).abs().max() / want2.abs().max()
    assert err < 2e-2, f"bf16-out rel err {err:.3e}"
    # bf16 output without a mask (the staged epilogue's other flow: buffers recycled on the TMA stores' read completion)
    out = ops.gemm_bf16(Ad, Bd, Mr, N, K, ak, bk, bias=dev(bias), relu=True, split_k=split).cpu().double()
    want3 = torch.relu(want + bias.double())
    err = (out - want3).abs().max() / want3.abs().max()
    assert err < 2e-2, f"bf16-out (no mask) rel err {err:.3e}"


@pytest.mark.parametrize("M,N,K,f32", [(4096, 512, 64, False), (4096, 512, 128, False), (2048, 128, 512, True), (96, 64, 136, True)])
def test_gemm_bf16_tcgen05_interleaved_output(M, N, K, f32, ops):
    """out_layout = RBI32: the same product, written row-block-interleaved ([M/32][N/V][32][V], V = elements per 16
    bytes) -- bit-identical to the row-major result after undoing the permutation."""
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    g = torch.Generator().manual_seed(M + N + K)
    A = dev((torch.randn(M, K, generator=g) / K ** 0.25).to(torch.bfloat16))
    B = dev((torch.randn(N, K, generator=g) / K ** 0.25).to(torch.bfloat16))
    bias = dev(torch.randn(N, generator=g))
    dt = torch.float32 if f32 else torch.bfloat16
    plain = ops.gemm_bf16(A, B, M, N, K, True, True, bias=bias, out_dtype=dt)
    tiled = ops.gemm_bf16(A, B, M, N, K, True, True, bias=bias, out_dtype=dt, rbi=True)
    V = 4 if f32 else 8
    back = tiled.view(M // 32, N // V, 32, V).permute(0, 2, 1, 3).reshape(M, N)
    assert torch.equal(back, plain)


# ------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("B", [1, 12, 512, 8192, 20000])
@pytest.mark.parametrize("pw", [1.0, 0.6666])
def test_bce_logits(B, pw, ops):
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, generator=g) * 3
    y = (torch.rand(B, generator=g) > 0.4).float()
    xr = x.clone().requires_grad_(True)
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(pw)) if pw != 1.0 else torch.nn.BCEWithLogitsLoss()
    loss = crit(xr, y); loss.backward()
    r = ops.bce_logits(dev(x), dev(y), pw, want_probs=True)
    assert abs(r["loss"].item() - loss.item()) <= 1e-6 * max(1.0, abs(loss.item()))
    assert torch.allclose(r["dlogits"].cpu(), xr.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(r["probs"].cpu(), torch.sigmoid(x), rtol=1e-6, atol=1e-7)
    pred = (torch.sigmoid(x) > 0.5).float()
    assert torch.equal(r["preds"].cpu(), pred)
    tn, fp, fn, tp = r["counts"].cpu().tolist()
    assert (tn, fp, fn, tp) == (int(((y == 0) & (pred == 0)).sum()), int(((y == 0) & (pred == 1)).sum()),
                                int(((y == 1) & (pred == 0)).sum()), int(((y == 1) & (pred == 1)).sum()))
    r2 = ops.bce_logits(dev(x), dev(y), pw)
    assert r2["loss"].item() == r["loss"].item()       # deterministic


@pytest.mark.parametrize("C", [5, 6])
@pytest.mark.parametrize("B", [12, 777, 9000])
def test_ce_logits(B, C, ops):
    g = torch.Generator().manual_seed(B + C)
    x = torch.randn(B, C, generator=g) * 2
    t = torch.randint(0, C, (B,), generator=g)
    w = torch.rand(C, generator=g) + 0.5
    for weight in (None, w):
        xr = x.clone().requires_grad_(True)
        loss = torch.nn.CrossEntropyLoss(weight=weight)(xr, t); loss.backward()
        r = ops.ce_logits(dev(x), dev(t.int()), None if weight is None else dev(weight), want_probs=True)
        assert abs(r["loss"].item() - loss.item()) <= 2e-6 * max(1.0, abs(loss.item()))
        assert torch.allclose(r["dlogits"].cpu(), xr.grad, rtol=1e-5, atol=1e-9)
        assert torch.allclose(r["probs"].cpu(), torch.softmax(x, 1), rtol=1e-5, atol=1e-7)
        assert torch.equal(r["preds"].cpu().long(), torch.argmax(x, 1))
        cm = torch.zeros(C, C, dtype=torch.long)
        for a, b in zip(t.tolist(), torch.argmax(x, 1).tolist()):
            cm[a, b] += 1
        assert torch.equal(r["cm"].cpu(), cm)
    # cascade train reduction: CE(label-1) masked to label != 0, preds forced to 0 where label == 0
    t6 = torch.randint(0, C + 1, (B,), generator=g)
    mask = (t6 != 0).float()
    xr = x.clone().requires_grad_(True)
    per = torch.nn.CrossEntropyLoss(reduction="none")(xr, (t6 - 1).clamp(min=0)) * mask
    loss = per.sum() / mask.sum(); loss.backward()
    r = ops.ce_logits(dev(x), dev(t6.int()), None, dev(mask), target_shift=-1, reduction=1, pred_shift=1, pred_mask_mode=1,
                      cm_classes=C + 1)
    assert abs(r["loss"].item() - loss.item()) <= 2e-6 * max(1.0, abs(loss.item()))
    assert torch.allclose(r["dlogits"].cpu(), xr.grad, rtol=1e-5, atol=1e-9)
    want_pred = torch.where(t6 == 0, torch.zeros_like(t6), torch.argmax(x, 1) + 1)
    assert torch.equal(r["preds"].cpu().long(), want_pred)
    # cascade validation quirk: sum of ALL losses if any window fired, else the mean
    fired = (torch.rand(B, generator=g) > 0.5).float()
    per = torch.nn.CrossEntropyLoss(reduction="none")(x, (t6 - 1).clamp(min=0))
    r = ops.ce_logits(dev(x), dev(t6.int()), None, dev(fired), target_shift=-1, reduction=3, want_grad=False, pred_shift=1,
                      pred_mask_mode=2, cm_classes=C + 1)
    assert abs(r["loss"].item() - per.sum().item()) <= 2e-6 * per.sum().item()
    r0 = ops.ce_logits(dev(x), dev(t6.int()), None, dev(torch.zeros(B)), target_shift=-1, reduction=3, want_grad=False)
    assert abs(r0["loss"].item() - per.mean().item()) <= 2e-6 * per.mean().item()


@pytest.mark.parametrize("T", [1, 150, 5000])
def test_ce_frame(T, ops):
    g = torch.Generator().manual_seed(T)
    S = 2
    x = torch.randn(S, 1, 2, T, generator=g)
    e = (torch.rand(1, T, generator=g) > 0.5).float()
    xr = x.clone().requires_grad_(True)
    target = torch.cat((1 - e, e), dim=0).transpose(1, 0)
    crit = torch.nn.CrossEntropyLoss()
    loss = sum(crit(xr[j].squeeze(0).transpose(1, 0), target) for j in range(S)) / S
    loss.backward()
    r = ops.ce_frame(dev(x), dev(e))
    assert abs(r["loss"].item() - loss.item()) <= 2e-6 * max(1.0, abs(loss.item()))
    assert torch.allclose(r["dlogits"].cpu(), xr.grad, rtol=1e-5, atol=1e-9)
    assert torch.equal(r["preds"].cpu().long(), torch.max(x[-1].squeeze(0).transpose(1, 0), 1)[1])


# ------------------------------------------------------------------------------------------- Adam
def test_adam_matches_torch(ops):
    g = torch.Generator().manual_seed(0)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, weight_decay=1e-4)
    p, m, v = dev(p0.clone()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pad = (n + 3) // 4 * 4
    buf = [torch.zeros(pad, device="cuda") for _ in range(4)]
    buf[0][:n] = p
    state = torch.zeros(4, device="cuda"); state[1] = 3e-4
    for step in range(5):
        grad = torch.randn(n, generator=g)
        ref.grad = grad.clone(); opt.step()
        buf[1][:n] = dev(grad)
        ops.adam_advance(state, 0.9, 0.999)
        ops.adam_step(buf[0], buf[1], buf[2], buf[3], state, 0.9, 0.999, 1e-8, 1e-4)
        assert torch.allclose(buf[0][:n].cpu(), ref.data, rtol=1e-6, atol=1e-7), step
    assert state[0].item() == 5.0


# ------------------------------------------------------------------------------------------- post-processing
@pytest.mark.parametrize("case", cases.WINDOW_CASES[:3], ids=[c[0] for c in cases.WINDOW_CASES[:3]])
def test_window_predictions_golden(case, golden_dir):
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    name, seed, nv, lo, hi, W, S = case
    gold = np.load(os.path.join(golden_dir, "window_predictions.npz"))
    g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
    names = np.concatenate([[synthetic.trial_name(nv - 1 - i)] * int(offsets[i + 1] - offsets[i]) for i in range(nv)])
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    pb = (rng.random(len(g)) > 0.5).astype(np.float64)
    pm = rng.integers(0, 6, len(g)).astype(np.float64)
    for tag, p, binary in (("bin", pb, True), ("multi", pm, False)):
        pw, ew, gw, sw = mu.window_predictions(p, e5[:, 4].astype(np.float64), g.astype(np.float64), names, W, S, binary)
        assert np.array_equal(pw.numpy().reshape(-1), gold[f"{name}/{tag}/preds"])
        assert np.array_equal(ew.numpy().reshape(-1), gold[f"{name}/{tag}/labels"])
        assert np.array_equal(gw.numpy().reshape(-1), gold[f"{name}/{tag}/gest"])
        assert sw["subject"].tolist() == gold[f"{name}/{tag}/subj"].tolist()


def test_ensemble_fusion(ops):
    from oracle import window_index as O
    rng = np.random.Generator(np.random.PCG64(1))
    n = 100_000
    pa, pb = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    pa[:4], pb[:4] = [0.5, 0.25, 0.75, 0.0], [0.5, 0.75, 0.25, 1.0]   # exact ties at 0.5 count as errors (>=)
    lab = (rng.random(n) > 0.5).astype(np.float32)
    preds, counts = ops.soft_vote(dev(pa), dev(pb), dev(lab))
    want = O.soft_vote(pa, pb)
    assert np.array_equal(preds.cpu().numpy().astype(np.int64), want)
    cm = [[int(((lab == a) & (want == b)).sum()) for b in (0, 1)] for a in (0, 1)]
    assert counts.cpu().numpy().reshape(2, 2).tolist() == cm
    b = rng.integers(0, 2, n).astype(np.int32); m = rng.integers(0, 6, n).astype(np.int32)
    assert np.array_equal(ops.cascade(dev(b), dev(m)).cpu().numpy(), O.cascade(b, m))
    cmk = ops.confusion(dev(m), dev(ops.cascade(dev(b), dev(m)).cpu().numpy()), 6).cpu().numpy()
    assert cmk.sum() == n and np.array_equal(np.diag(cmk) + 0, [int(((m == c) & (O.cascade(b, m) == c)).sum()) for c in range(6)])


DEV = "cuda"


@pytest.mark.parametrize("world,n", [(2, 4096), (4, 10_001), (8, 1_599_265)])
def test_peer_allreduce_kernel_ranks_emulated_on_one_gpu(world, n):
    """b200med_peer_allreduce_f32 (the data-parallel gradient exchange, csrc/peer_exchange.cu) with the `world` ranks played by
    `world` streams of ONE GPU: every "rank" owns a buffer and a flag block, sees the others through the same pointer arrays a
    real rank gets from CUDA IPC, and launches the kernel once per exchange.  After each of three exchanges every buffer must
    hold the sum in rank order, bit for bit the same in all of them (the real multi-GPU run is scripts/dp_window_check.py)."""
    import ctypes as C
    from multimodal_error_detection_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(world * 7 + 1)
    bufs = [torch.empty(n, device=DEV) for _ in range(world)]
    flags = [torch.zeros(int(lib.b200med_peer_flag_bytes()) // 4, dtype=torch.int32, device=DEV) for _ in range(world)]
    bufs_dev = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    flags_dev = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=DEV)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for rep in range(3):
        src = [torch.randn(n, device=DEV, generator=g) * (q + 1) for q in range(world)]
        want = src[0].clone()
        for q in range(1, world):
            want += src[q]                      # rank order, fp32: what the kernel computes
        for q in range(world):
            bufs[q].copy_(src[q])
        torch.cuda.synchronize()
        # at most ceil(n / world / 2048) CTAs per "rank": all of them fit the GPU together, as the flag barriers require
        with _lib_sm_limit(lib, max(1, 148 // world)):
            for q in range(world):
                with torch.cuda.stream(streams[q]):
                    _lib.call("b200med_peer_allreduce_f32", C.c_void_p(bufs_dev.data_ptr()), C.c_void_p(flags_dev.data_ptr()), q, world, n,
                              C.c_void_p(streams[q].cuda_stream))
        torch.cuda.synchronize()
        for q in range(world):
            assert torch.equal(bufs[q], want), (rep, q)
        assert all(int(f[32]) == rep + 1 and int(f[33]) == 0 and int(f[34]) == 0 for f in flags)


class _lib_sm_limit:
    def __init__(self, lib, sms):
        self.lib, self.sms = lib, sms

    def __enter__(self):
        self.prev = self.lib.b200med_set_sm_limit(self.sms)

    def __exit__(self, *exc):
        self.lib.b200med_set_sm_limit(self.prev)


@pytest.mark.parametrize("Ca,Cb,stat_rows,B,W", [(32, 26, 1, 77, 16), (32, 26, 16, 77, 16), (5, 3, 1, 40, 10), (31, 27, 1, 9, 7), (64, 0, 1, 33, 4)])
def test_lstm_pack_parts_kernel(Ca, Cb, stat_rows, B, W):
    """b200med_lstm_pack_parts (FeatureExtractor output + kinematics gathered / standardised from the frame table -> first operand
    of the LSTM recurrence, bf16, time-major) against its definition built with torch: columns [0, Ca) = bf16(feats),
    [Ca, Ca + Cb) = bf16((table[start + t] - mean) / std) (IEEE subtract / divide, as CustomWindowDataset.py:56-60), padding
    and the h_{-1} columns of step 0 zero, the h columns of the later steps untouched.  Even / odd widths, one statistics row or
    one per step: both the 64-column fast path and the generic kernel."""
    import ctypes as C
    from multimodal_error_detection_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(Ca * 100 + Cb)
    H, hoff = 128, 64
    Kp, Bp, N = hoff + H, (B + 31) // 32 * 32, 3000
    feats = torch.randn(B, W, Ca, device=DEV, generator=g)
    table = torch.randn(N, max(Cb, 1), device=DEV, generator=g)[:, :Cb].contiguous() if Cb else torch.zeros(N, 0, device=DEV)
    mean = torch.randn(stat_rows, max(Cb, 1), device=DEV, generator=g)[:, :Cb].contiguous()
    std = (torch.rand(stat_rows, max(Cb, 1), device=DEV, generator=g) + 0.5)[:, :Cb].contiguous()
    starts = torch.randint(0, N - W, (B,), device=DEV, generator=g, dtype=torch.int64).to(torch.int32)
    sentinel = 7.0
    A0 = torch.full((W, Bp, Kp), sentinel, dtype=torch.bfloat16, device=DEV)
    P = lambda t: C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)
    _lib.call("b200med_lstm_pack_parts", P(feats), Ca, P(table), N, Cb, P(mean if Cb else None), P(std if Cb else None), stat_rows,
              P(starts), P(A0), B, Bp, W, H, Kp, hoff, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rows = starts.long()[:, None] + torch.arange(W, device=DEV)[None]                  # [B, W]
    want = torch.full((W, Bp, Kp), sentinel, dtype=torch.float32, device=DEV)
    want[:, :B, :Ca] = feats.permute(1, 0, 2)
    if Cb:
        kin = table[rows]                                                              # [B, W, Cb]
        m = mean[0] if stat_rows == 1 else mean[None, :, :]
        s = std[0] if stat_rows == 1 else std[None, :, :]
        want[:, :B, Ca:Ca + Cb] = ((kin - m) / s).permute(1, 0, 2)
    want[:, :B, Ca + Cb:hoff] = 0.0
    want[0, :B, hoff:] = 0.0                                                           # h_{-1}; later steps keep the sentinel
    assert torch.equal(A0.view(torch.int16), want.to(torch.bfloat16).view(torch.int16))


# ------------------------------------------------------------------------------------------- fp32 products on the tensor cores
def test_split_bf16x6_is_exact_and_laid_out_for_both_roles(ops):
    """x = h + m + l exactly (three bf16 terms); blocks (m, l, h, m, h, h) for a left operand, (m, h, l, h, m, h) for a right
    one, side by side in every row (row6) or stacked (stack6); relu folds max(., 0) in front of the split."""
    g = torch.Generator().manual_seed(6)
    for R, Cn in ((37, 24), (5, 7)):                    # vectorised and scalar instantiation
        x = torch.randn(R, Cn, generator=g) * torch.logspace(-6, 6, Cn)[None, :]
        for relu in (False, True):
            xr = x.clamp_min(0) if relu else x
            h = xr.to(torch.bfloat16); m = (xr - h.float()).to(torch.bfloat16); l = (xr - h.float() - m.float()).to(torch.bfloat16)
            assert torch.equal(h.float() + m.float() + l.float(), xr)
            for role, order in ((0, (m, l, h, m, h, h)), (1, (m, h, l, h, m, h))):
                row6, stack6 = ops.split_bf16x6(dev(x), role, True, True, relu=relu)
                assert torch.equal(row6.cpu().view(torch.int16), torch.cat(order, dim=1).view(torch.int16)), (R, Cn, relu, role)
                assert torch.equal(stack6.cpu().view(torch.int16), torch.cat(order, dim=0).view(torch.int16)), (R, Cn, relu, role)


@pytest.mark.parametrize("M,N,K", [(4096, 512, 2048), (8192, 256, 512), (16384, 512, 128)])
def test_fp32_linear_ops_on_the_tensor_cores(M, N, K, ops):
    """The fp32 mode's large products (forward, data gradient with ReLU mask, weight + bias gradient) take the six-product
    bf16 route (ops._fp32_tc): held against fp64 at 5e-6 norm-wise -- the fp32 FMA kernels of the same ops (route switched off)
    sit at ~6e-7 -- and bit-reproducible run to run."""
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    assert ops._fp32_tc(M, N, K, 6 * K) and ops._fp32_tc(M, K, N, 6 * N, K) and ops._fp32_tc(N, K, M, N, K)
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).clamp_min(0)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    rel = lambda a, r: float((a.double().cpu() - r).norm() / r.norm())
    y64 = torch.relu(x.double() @ w.double().T + b.double())
    dx64 = (dy.double() @ w.double()) * (x.double() > 0)
    dw64, db64 = dy.double().T @ x.double(), dy.double().sum(0)
    res = {}
    for tc in (True, False):
        old = ops.FP32_TC_MIN_FLOP
        ops.FP32_TC_MIN_FLOP = old if tc else 0.0
        try:
            y = ops.linear_fwd_f32(dev(x), dev(w), dev(b), relu=True)
            y2 = ops.linear_f32(dev(x), dev(w), dev(b), ops.GEMM_RELU)
            dx = ops.linear_bwd_data_f32(dev(dy), dev(w), relu_out=dev(x))
            dx2 = ops.linear_dgrad_f32(dev(dy), dev(w), mask=dev(x))
            dw, db = ops.linear_bwd_weight_f32(dev(dy), dev(x))
            dw2 = ops.linear_wgrad_f32(dev(dy), dev(x), relu_x=True)
        finally:
            ops.FP32_TC_MIN_FLOP = old
        res[tc] = (y, dx, dw)
        bar = 5e-6 if tc else 2e-6
        assert rel(y, y64) < bar and rel(y2, y64) < bar, (tc, rel(y, y64))
        assert rel(dx, dx64) < bar and rel(dx2, dx64) < bar, (tc, rel(dx, dx64))
        assert rel(dw, dw64) < bar and rel(dw2, dw64) < bar, (tc, rel(dw, dw64))
        assert rel(db, db64) < 2e-6
        if tc:
            assert torch.equal(y, y2) and torch.equal(dx, dx2) and torch.equal(dw, dw2)
    dw_again, _ = ops.linear_bwd_weight_f32(dev(dy), dev(x))
    assert torch.equal(dw_again, res[True][2])
