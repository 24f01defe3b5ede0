"""The head kernels (csrc/head.cu, the strided fp32 GEMM, the exact-math fp32 LSTM cells) against torch CPU fp32/fp64
definitions of the reference layers (MED/modeling/models.py:49-131, 135-220): forward, every gradient, BatchNorm running
statistics.  1e-5 norm-wise relative is the north_star fp32 bar."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,C", [(12, 16), (8192, 256), (1000, 64), (33, 130)])
def test_batchnorm_kernels(M, C):
    from multimodal_error_detection_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x = torch.relu(torch.randn(M, C, generator=g) * 2 + 0.3)
    dy = torch.randn(M, C, generator=g)
    bn = nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5); bn.bias.copy_(torch.randn(C, generator=g))
        bn.running_mean.copy_(torch.randn(C, generator=g)); bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    rm, rv = bn.running_mean.clone().float().to(DEV), bn.running_var.clone().float().to(DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    xr = x.double().requires_grad_(True)
    bn.train()
    y = bn(xr)
    # the head applies ReLU in front of the BatchNorm; relu_mask folds its backward in: d(pre) = dx * (x > 0)
    y.backward(dy.double())
    gam, bet = bn.weight.detach().float().to(DEV), bn.bias.detach().float().to(DEV)
    yk, sm, sr = ops.bn_fwd(x.to(DEV), gam, bet, bn.eps, bn.momentum, True, rm, rv, nbt)
    assert nrel(yk, y) < 1e-5
    assert nrel(rm, bn.running_mean) < 1e-6 and nrel(rv, bn.running_var) < 1e-5 and int(nbt) == 1
    for mask in (False, True):
        dx, dg, db = ops.bn_bwd(dy.to(DEV), x.to(DEV), gam, sm, sr, relu_mask=mask)
        want = xr.grad * (x > 0).double() if mask else xr.grad
        assert nrel(dx, want) < 1e-5
        assert nrel(dg, bn.weight.grad) < 1e-5 and nrel(db, bn.bias.grad) < 1e-5
    y2, _, _ = ops.bn_fwd(x.to(DEV), gam, bet, bn.eps, bn.momentum, True, rm.clone(), rv.clone(), None)
    assert torch.equal(y2, yk)                                     # deterministic
    bn.eval()
    ye, _, _ = ops.bn_fwd(x.to(DEV), gam, bet, bn.eps, bn.momentum, False, rm, rv, None)
    assert nrel(ye, bn(x.double())) < 1e-5


@pytest.mark.parametrize("relu_in,dims,B", [(True, [128, 256, 64, 1], 700), (False, [128, 256, 32, 16, 6], 12), (True, [128, 256, 64, 6], 8192)])
def test_mlp_tail_vs_torch(relu_in, dims, B):
    """[ReLU ->] (Linear, ReLU, BatchNorm1d)*, Linear: logits, input gradient, every parameter gradient and the running
    statistics against the same nn.Sequential in fp64 on the CPU."""
    from multimodal_error_detection_b200.heads import mlp_tail
    torch.manual_seed(B)
    mods = []
    for i in range(len(dims) - 2):
        mods += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU(), nn.BatchNorm1d(dims[i + 1])]
    mods += [nn.Linear(dims[-2], dims[-1])]
    seq = nn.Sequential(*mods)
    for m in seq:
        if isinstance(m, nn.BatchNorm1d):
            nn.init.uniform_(m.weight, 0.5, 1.5); nn.init.normal_(m.bias)
    ref = nn.Sequential(*[type(m)(*([m.in_features, m.out_features] if isinstance(m, nn.Linear) else [m.num_features] if isinstance(m, nn.BatchNorm1d) else []))
                          for m in seq]).double()
    ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in seq.state_dict().items()})
    seq = seq.to(DEV).train(); ref.train()
    x = torch.randn(B, dims[0])
    dy = torch.randn(B, dims[-1])
    xr = x.double().requires_grad_(True)
    yr = ref(torch.relu(xr) if relu_in else xr)
    yr.backward(dy.double())
    xg = x.to(DEV).requires_grad_(True)
    yg = mlp_tail(xg, seq, relu_in=relu_in, training=True)
    yg.backward(dy.to(DEV))
    assert nrel(yg, yr) < 1e-5
    assert nrel(xg.grad, xr.grad) < 2e-5
    for (k, p), (_, q) in zip(seq.named_parameters(), ref.named_parameters()):
        scale = max(float(q.grad.abs().max()), 1e-4 * max(float(r.grad.abs().max()) for r in ref.parameters()))
        assert float((p.grad.double().cpu() - q.grad).abs().max()) <= 5e-5 * scale, k
    for (k, b), (_, c) in zip(seq.named_buffers(), ref.named_buffers()):
        assert nrel(b.double(), c.double()) < 1e-5, k
    seq.eval(); ref.eval()
    with torch.no_grad():
        assert nrel(mlp_tail(x.to(DEV), seq, relu_in=relu_in, training=False), ref(torch.relu(x.double()) if relu_in else x.double())) < 1e-5


@pytest.mark.parametrize("relu_in,dims,B", [(True, [128, 256, 64, 1], 8192), (True, [128, 256, 64, 1], 700), (False, [128, 256, 32, 16, 6], 12)])
def test_mlp_tail_bf16_mode_vs_torch(relu_in, dims, B):
    """The same head tail in the bf16 throughput mode: hidden-layer products with K % 64 == 0 run on the tcgen05 GEMM as split-bf16
    products (hi + lo operands, three terms folded into one GEMM; BatchNorm and the output layer fp32), against the exact
    nn.Sequential in fp64 on the CPU.  The products carry ~1e-5 relative error, so the logits are held to 1e-4; gradients to
    5e-3 NORM-wise: a pre-activation within ~1e-5 of zero may flip its ReLU, and each flipped unit (a handful in 2 M at
    B = 8192) moves a whole row of the input gradient."""
    from multimodal_error_detection_b200 import ops
    from multimodal_error_detection_b200.heads import mlp_tail
    if not ops.has_tcgen05():
        pytest.skip("needs a compute-capability 10.x device")
    torch.manual_seed(B + 1)
    mods = []
    for i in range(len(dims) - 2):
        mods += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU(), nn.BatchNorm1d(dims[i + 1])]
    mods += [nn.Linear(dims[-2], dims[-1])]
    seq = nn.Sequential(*mods)
    for m in seq:
        if isinstance(m, nn.BatchNorm1d):
            nn.init.uniform_(m.weight, 0.5, 1.5); nn.init.normal_(m.bias)
    ref = nn.Sequential(*[type(m)(*([m.in_features, m.out_features] if isinstance(m, nn.Linear) else [m.num_features] if isinstance(m, nn.BatchNorm1d) else []))
                          for m in seq]).double()
    ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in seq.state_dict().items()})
    seq = seq.to(DEV).train(); ref.train()
    x = torch.randn(B, dims[0])
    dy = torch.randn(B, dims[-1])
    xr = x.double().requires_grad_(True)
    yr = ref(torch.relu(xr) if relu_in else xr)
    yr.backward(dy.double())
    xg = x.to(DEV).requires_grad_(True)
    yg = mlp_tail(xg, seq, relu_in=relu_in, training=True, precision="bf16")
    yg.backward(dy.to(DEV))

    def frel(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert nrel(yg, yr) < 1e-4
    assert frel(xg.grad, xr.grad) < 5e-3
    gmax = max(float(r.grad.norm()) for r in ref.parameters())
    for (k, p), (_, q) in zip(seq.named_parameters(), ref.named_parameters()):
        scale = max(float(q.grad.norm()), 1e-3 * gmax)
        assert float((p.grad.double().cpu() - q.grad).norm()) <= 5e-3 * scale, k
    for (k, b), (_, c) in zip(seq.named_buffers(), ref.named_buffers()):
        assert nrel(b.double(), c.double()) < 1e-4, k
    seq.eval(); ref.eval()
    with torch.no_grad():
        ye = mlp_tail(x.to(DEV), seq, relu_in=relu_in, training=False, precision="bf16")
        assert nrel(ye, ref(torch.relu(x.double()) if relu_in else x.double())) < 1e-4


def _tail_seq(dims, seed):
    torch.manual_seed(seed)
    mods = []
    for i in range(len(dims) - 2):
        mods += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU(), nn.BatchNorm1d(dims[i + 1])]
    mods += [nn.Linear(dims[-2], dims[-1])]
    seq = nn.Sequential(*mods)
    for m in seq:
        if isinstance(m, nn.BatchNorm1d):
            nn.init.uniform_(m.weight, 0.5, 1.5); nn.init.normal_(m.bias)
            m.running_mean.normal_(); m.running_var.uniform_(0.5, 1.5)
    return seq.to(DEV)


@pytest.mark.parametrize("dims,B,relu_in", [([128, 256, 64, 1], 8192, True), ([128, 256, 64, 1], 70, True), ([128, 256, 64, 6], 64, True),
                                             ([64, 128, 5], 3, False), ([256, 128, 256, 64, 8], 129, True), ([128, 64, 1], 1000, False)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_tail_vs_layerwise(dims, B, relu_in, precision, monkeypatch):
    """The three-kernels-per-direction tail (csrc/mlp_tail.cu) against the layer-at-a-time kernels it replaces: logits, input
    gradient, every parameter gradient, running statistics; ragged last row tile, 1-3 hidden layers, 1-8 output columns;
    bit-identical when repeated; 2 n_hidden + 1 launches forward (a layer kernel and a statistics finalize per hidden layer, the output layer)."""
    import copy
    from multimodal_error_detection_b200 import _lib, heads
    seq_f = _tail_seq(dims, B).train()
    seq_l = copy.deepcopy(seq_f)
    x = torch.randn(B, dims[0], device=DEV)
    dy = torch.randn(B, dims[-1], device=DEV)
    n_hidden = len(dims) - 2
    assert heads._fused_ok(x, [p for m in seq_f if not isinstance(m, nn.ReLU) for p in
                               ([m.weight, m.bias] + ([m.running_mean, m.running_var, m.num_batches_tracked] if isinstance(m, nn.BatchNorm1d) else []))],
                           n_hidden)

    def run(seq):
        for p in seq.parameters():
            p.grad = None
        xg = x.clone().requires_grad_(True)
        n0 = _lib.launch_count()
        y = heads.mlp_tail(xg, seq, relu_in=relu_in, training=True, precision=precision)
        n1 = _lib.launch_count()
        y.backward(dy)
        from multimodal_error_detection_b200 import lstm_stack
        lstm_stack.join_pending()
        torch.cuda.synchronize()
        return y.detach(), xg.grad, [p.grad.clone() for p in seq.parameters()], [b.clone() for b in seq.buffers()], n1 - n0

    yf, gxf, gpf, bf, launches = run(seq_f)
    assert launches == 2 * n_hidden + 1
    monkeypatch.setattr(heads, "FUSED_TAIL", False)
    yl, gxl, gpl, bl, launches_l = run(seq_l)
    assert launches_l > launches
    monkeypatch.setattr(heads, "FUSED_TAIL", True)
    tol = 2e-5 if precision == "fp32" else 2e-4
    assert nrel(yf, yl) < tol
    if precision == "fp32":
        assert nrel(gxf, gxl) < 5 * tol
        gmax = max(float(g.abs().max()) for g in gpl)
        for (k, _), a, b in zip(seq_f.named_parameters(), gpf, gpl):
            scale = max(float(b.abs().max()), 1e-4 * gmax)
            assert float((a - b).abs().max()) <= 10 * tol * scale, k
    else:       # the layer-at-a-time bf16 mode flips the ReLU of pre-activations within ~1e-5 of zero (a handful per batch, each
        # moves a row of the input gradient): norm-wise, like test_mlp_tail_bf16_mode_vs_torch
        def frel(a, b):
            return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        assert frel(gxf, gxl) < 1e-2
        gmax = max(float(g.norm()) for g in gpl)
        for (k, _), a, b in zip(seq_f.named_parameters(), gpf, gpl):
            assert float((a.double() - b.double()).norm()) <= 1e-2 * max(float(b.norm()), 1e-3 * gmax), k
    for (k, _), a, b in zip(seq_f.named_buffers(), bf, bl):
        assert nrel(a.double(), b.double()) < 1e-5, k
    seq_r = _tail_seq(dims, B).train()
    yr, gxr, gpr, br, _ = run(seq_r)
    assert torch.equal(yr, yf) and torch.equal(gxr, gxf) and all(torch.equal(a, b) for a, b in zip(gpr, gpf))
    seq_f.eval(); seq_l.eval()
    with torch.no_grad():
        ye = heads.mlp_tail(x, seq_f, relu_in=relu_in, training=False, precision=precision)
        monkeypatch.setattr(heads, "FUSED_TAIL", False)
        yl = heads.mlp_tail(x, seq_l, relu_in=relu_in, training=False, precision=precision)
    assert nrel(ye, yl) < tol


def test_split_bf16x3_kernel():
    """b200med_split_bf16x3: hi = bf16(x), lo = bf16(x - hi) in both layouts and both operand orders, bit for bit."""
    from multimodal_error_detection_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(37, 24, generator=g) * 3).to(DEV)
    for relu in (False, True):
        v = torch.relu(x) if relu else x
        hi = v.to(torch.bfloat16)
        lo = (v - hi.float()).to(torch.bfloat16)
        for ro, so in ((0, 0), (1, 1), (0, 1), (1, 0)):
            row3, stack3 = ops.split_bf16x3(x, ro, so, relu=relu)
            want_row = torch.cat([hi, hi if ro else lo, lo if ro else hi], dim=1)
            want_stack = torch.cat([hi, hi if so else lo, lo if so else hi], dim=0)
            assert torch.equal(row3.view(torch.int16), want_row.view(torch.int16))
            assert torch.equal(stack3.view(torch.int16), want_stack.view(torch.int16))
    r, s = ops.split_bf16x3(x, None, 1)
    assert r is None and s.shape == (111, 24)


@pytest.mark.parametrize("W,B,view", [(10, 12, True), (10, 513, False), (30, 40, True), (30, 7, False)])
def test_cnn_head_vs_torch(W, B, view):
    """The whole CNN head (conv-as-GEMM over overlapping time-major rows, pool + dropout(0), BatchNorm, MLP tail) against the
    same modules in fp64 on the CPU: logits, input gradient, parameter gradients; both input layouts (a permuted view of
    [B, W, F] as define_inputs builds it, and a plain contiguous [B, F, W])."""
    from multimodal_error_detection_b200.modeling.models import CNN
    from oracle import nets
    torch.manual_seed(W + B)
    model = CNN(58, W, 1)
    ref = nets.OracleCNN(58, W, 1)
    ref.load_state_dict(model.state_dict())
    for m in list(model.modules()) + list(ref.modules()):
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    ref = ref.double().train()
    model = model.to(DEV).train()
    x = torch.randn(B, W, 58)
    dy = torch.randn(B, 1)
    xr = x.double().requires_grad_(True)
    yr = ref(xr.permute(0, 2, 1))
    yr.backward(dy.double())
    if view:
        xg = x.to(DEV).requires_grad_(True)
        yg = model(xg.permute(0, 2, 1))
        yg.backward(dy.to(DEV))
        gx = xg.grad
    else:
        xg = x.permute(0, 2, 1).contiguous().to(DEV).requires_grad_(True)
        yg = model(xg)
        yg.backward(dy.to(DEV))
        gx = xg.grad.permute(0, 2, 1)
    assert nrel(yg, yr) < 2e-5
    assert nrel(gx, xr.grad) < 5e-5
    gmax = max(float(q.grad.abs().max()) for q in ref.parameters())
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        # gradients that are zero in exact arithmetic (a bias in front of a BatchNorm) are fp32 round-off of sums of O(gmax)
        # terms on our side and fp64 round-off on the reference's: the comparison scale is floored at 1e-2 of the largest gradient
        scale = max(float(q.grad.abs().max()), 1e-2 * gmax)
        assert float((p.grad.double().cpu() - q.grad).abs().max()) <= 1e-4 * scale, k
    for (k, b), (k2, c) in zip([kv for kv in model.named_buffers() if "_drop_seed" not in kv[0]], ref.named_buffers()):
        assert k == k2 and nrel(b.double(), c.double()) < 1e-5, k


def test_cnn_head_dropout_mask_consistent():
    """Dropout(0.2) inside the conv blocks: forward and backward regenerate the same counter-based mask (finite-difference
    check of a random projection of the logits along the analytic input gradient), and a new seed draws a new mask."""
    from multimodal_error_detection_b200.modeling.models import CNN
    torch.manual_seed(0)
    model = CNN(58, 10, 3).to(DEV).train()
    x = torch.randn(256, 10, 58, device=DEV)
    r = torch.randn(256, 3, device=DEV)
    xg = x.clone().requires_grad_(True)
    y = model(xg.permute(0, 2, 1))
    (y * r).sum().backward()
    seed = model._drop_seed.clone()
    with torch.no_grad():
        v = xg.grad / xg.grad.norm()
        eps = 2e-2

        def f(inp):
            model._drop_seed.copy_(seed - 1)          # forward() advances the seed by one before use
            return (model(inp.permute(0, 2, 1)).double() * r.double()).sum()
        fd = float((f(x + eps * v) - f(x - eps * v)) / (2 * eps))
    # BatchNorm running stats move between the calls, but train-mode outputs use batch statistics only
    an = float((xg.grad * v).sum())
    assert abs(an) > 1e-2 and abs(fd - an) < 5e-2 * abs(an), (fd, an)
    y2 = model(x.permute(0, 2, 1))                    # next seed: another mask
    assert not torch.equal(y2, y.detach())
    # a wrong (unmasked) backward would differ: the p = 0 gradient is not the p = 0.2 gradient
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    x0 = x.clone().requires_grad_(True)
    (model(x0.permute(0, 2, 1)) * r).sum().backward()
    assert float((x0.grad - xg.grad).norm() / xg.grad.norm()) > 5e-2


@pytest.mark.parametrize("B,W,L,H", [(37, 10, 3, 128), (600, 16, 3, 128), (5, 1, 2, 64), (64, 6, 1, 96)])
def test_lstm_fp32_vs_torch_cpu(B, W, L, H):
    """fp32 parity mode of the LSTM head's recurrence (exact-math cell kernels + fp32 SIMT gate GEMMs) against nn.LSTM in
    fp64 on the CPU: last hidden state 1e-5, input gradient and every weight gradient 5e-5."""
    from multimodal_error_detection_b200.lstm_stack import lstm_last_hidden
    torch.manual_seed(B)
    F = 58
    lstm = nn.LSTM(F, H, num_layers=L, batch_first=True, dropout=0.0)
    ref = nn.LSTM(F, H, num_layers=L, batch_first=True, dropout=0.0).double()
    ref.load_state_dict({k: v.double() for k, v in lstm.state_dict().items()})
    lstm = lstm.to(DEV)
    x = torch.randn(B, W, F)
    gh = torch.randn(B, H)
    xr = x.double().requires_grad_(True)
    out, _ = ref(xr)
    out[:, -1, :].backward(gh.double())
    xg = x.to(DEV).requires_grad_(True)
    h = lstm_last_hidden(xg.permute(0, 2, 1), lstm, training=True, seed_dev=None, precision="fp32")
    h.backward(gh.to(DEV))
    assert nrel(h, out[:, -1, :]) < 1e-5
    assert nrel(xg.grad, xr.grad) < 5e-5
    for (k, p), (_, q) in zip(lstm.named_parameters(), ref.named_parameters()):
        if float(q.grad.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert nrel(p.grad, q.grad) < 5e-5, k
    # contiguous [B, F, W] input: same bits
    lstm.zero_grad()
    xc = x.permute(0, 2, 1).contiguous().to(DEV).requires_grad_(True)
    h2 = lstm_last_hidden(xc, lstm, training=True, seed_dev=None, precision="fp32")
    h2.backward(gh.to(DEV))
    assert torch.equal(h2, h) and torch.equal(xc.grad.permute(0, 2, 1), xg.grad)


def test_gemm_f32_flags():
    """b200med_gemm_f32 operand / epilogue flags against fp64: ReLU on A, ReLU on B (split reduction), accumulate, mask."""
    from multimodal_error_detection_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, N, K = 300, 70, 130
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    y0 = torch.randn(M, N, generator=g)
    xd, wd = x.double(), w.double()
    out = y0.clone().to(DEV)
    ops.linear_f32(x.to(DEV), w.to(DEV), b.to(DEV), ops.GEMM_RELU_A | ops.GEMM_ACCUM | ops.GEMM_RELU, out=out)
    assert nrel(out, torch.relu(torch.relu(xd) @ wd.T + b.double() + y0.double())) < 1e-5
    dy = torch.randn(M, N, generator=g)
    dw = ops.linear_wgrad_f32(dy.to(DEV), x.to(DEV), relu_x=True)
    assert nrel(dw, dy.double().T @ torch.relu(xd)) < 1e-5
    dx = ops.linear_dgrad_f32(dy.to(DEV), w.to(DEV), mask=x.to(DEV))
    assert nrel(dx, (dy.double() @ wd) * (xd > 0)) < 1e-5
