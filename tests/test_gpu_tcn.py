"""GPU parity tests of the TeCNo kernels (csrc/tcn.cu), called through the C ABI: every kernel against its torch
definition (tests/tcn_reference.py, computed in fp64 on the host), the whole MultiStageModel against the oracle TeCNo
(forward, input gradient, every parameter gradient), ragged batches against per-video runs, dropout mask consistency
between forward and backward.  fp32 FMA arithmetic: bars are 1e-5 relative to the largest magnitude of the compared
tensor (1e-4 for gradients reduced over T)."""
import pytest
import torch

import tcn_reference as ref
from oracle import nets

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from multimodal_error_detection_b200 import ops as _ops
    return _ops


def close(a, b, tol=1e-5):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    return err <= tol * (b.abs().max().item() + 1e-12), err


def _layer(seed=0):
    g = torch.Generator().manual_seed(seed)
    wd = torch.randn(64, 64, 3, generator=g) * 0.1
    bd = torch.randn(64, generator=g) * 0.1
    w1 = torch.randn(64, 64, 1, generator=g) * 0.1
    b1 = torch.randn(64, generator=g) * 0.1
    return wd, bd, w1, b1


def _pack_on_device(ops, layers):
    tensors = [t.to(DEV).contiguous() for layer in layers for t in layer]
    table = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64).to(DEV)
    pack = ops.tcn_pack(table, len(layers))
    torch.cuda.synchronize()
    return pack, tensors


def test_pack_layout(ops):
    layers = [_layer(1), _layer(2), _layer(3)]
    pack, _ = _pack_on_device(ops, layers)
    for l, layer in enumerate(layers):
        assert torch.equal(pack[l].cpu(), ref.pack_layer(*layer))


# T <= 1184: 4-frame CTAs (one frame per warp); <= 2368: 8-frame; longer: 16-frame.  The hidden backward: 8-frame tiles up to T = 640
GEOMS = [(77, 1, True), (77, 4, False), (600, 128, True), (333, 64, False), (16, 2, True), (5, 8, True), (1030, 1, True),
         (1500, 16, True), (2500, 32, False), (2500, 1, True)]


@pytest.mark.parametrize("T,dilation,causal", GEOMS)
def test_layer_forward_and_backward_kernels(ops, T, dilation, causal):
    layer = _layer(T + dilation)
    pack, _ = _pack_on_device(ops, [layer])
    pk = pack[0]
    g = torch.Generator().manual_seed(T)
    x = torch.randn(T, 64, generator=g)
    dout = torch.randn(T, 64, generator=g)
    pk64 = ref.pack_layer(*[t.double() for t in layer])
    out_r, y_r = ref.layer_fwd(x.double(), pk64, dilation, causal)
    out = torch.empty(T, 64, device=DEV)
    y = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_fwd(x.to(DEV), pk, out, y, dilation, causal)
    ok, err = close(out, out_r); assert ok, f"out {err}"
    ok, err = close(y, y_r); assert ok, f"y {err}"
    # eval form: no y_save
    out2 = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_fwd(x.to(DEV), pk, out2, None, dilation, causal)
    assert torch.equal(out, out2)
    # backward, part 1 (the reference is given the kernel's own y so the ReLU mask is identical)
    dpre_r, grad_r = ref.layer_bwd_hidden(dout.double(), x.double(), y.cpu().double(), pk64, dilation, causal)
    n_slots = ops.tcn_slots(T)
    assert 1 <= n_slots <= 80
    dpre = torch.empty(T, 64, device=DEV)
    partials = torch.full((1, n_slots, ops.TCN_GRAD_FLOATS), float("nan"), device=DEV)
    ops.tcn_layer_bwd_hidden(dout.to(DEV), x.to(DEV), y, pk, dpre, partials[0], n_slots, dilation, causal)
    ok, err = close(dpre, dpre_r); assert ok, f"dpre {err}"
    grads = ops.tcn_reduce_grads(partials, 1, n_slots)[0]
    for name, a, b in (("dWd", 0, ref.WD), ("dW1", ref.WD, ref.WD + ref.W1), ("dbd", ref.WD + ref.W1, ref.WD + ref.W1 + 64),
                       ("db1", ref.WD + ref.W1 + 64, ref.GRAD)):
        ok, err = close(grads[a:b], grad_r[a:b], 1e-4); assert ok, f"{name} {err}"
    # backward, part 2
    dx_r = ref.layer_bwd_input(dpre_r, dout.double(), pk64, dilation, causal)
    dx = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_bwd_input(dpre, dout.to(DEV), pk, dx, dilation, causal)
    ok, err = close(dx, dx_r); assert ok, f"dx {err}"
    # determinism of the partial sums
    partials2 = torch.empty_like(partials)
    ops.tcn_layer_bwd_hidden(dout.to(DEV), x.to(DEV), y, pk, dpre, partials2[0], n_slots, dilation, causal)
    assert torch.equal(ops.tcn_reduce_grads(partials2, 1, n_slots)[0], grads)


def test_layer_kernels_ragged(ops):
    from multimodal_error_detection_b200 import tcn
    lengths = [40, 9, 65, 16, 1]
    T = sum(lengths)
    tl, tr = tcn.ragged_geometry(lengths, DEV)
    layer = _layer(9)
    pack, _ = _pack_on_device(ops, [layer])
    pk64 = ref.pack_layer(*[t.double() for t in layer])
    g = torch.Generator().manual_seed(3)
    x, dout = torch.randn(T, 64, generator=g), torch.randn(T, 64, generator=g)
    for dilation, causal in ((4, True), (2, False), (32, True)):
        out_r, y_r = ref.layer_fwd(x.double(), pk64, dilation, causal, tl.cpu(), tr.cpu())
        out, y = torch.empty(T, 64, device=DEV), torch.empty(T, 64, device=DEV)
        ops.tcn_layer_fwd(x.to(DEV), pack[0], out, y, dilation, causal, tloc=tl, trem=tr)
        ok, err = close(out, out_r); assert ok, f"ragged out {err}"
        dpre_r, _ = ref.layer_bwd_hidden(dout.double(), x.double(), y.cpu().double(), pk64, dilation, causal, tl.cpu(), tr.cpu())
        dx_r = ref.layer_bwd_input(dpre_r, dout.double(), pk64, dilation, causal, tl.cpu(), tr.cpu())
        dx = torch.empty(T, 64, device=DEV)
        ops.tcn_layer_bwd_input(dpre_r.float().to(DEV), dout.to(DEV), pack[0], dx, dilation, causal, tloc=tl, trem=tr)
        ok, err = close(dx, dx_r); assert ok, f"ragged dx {err}"


def test_dropout_mask_is_shared_by_forward_and_backward(ops):
    T, dilation, p = 200, 2, 0.5
    layer = _layer(5)
    pack, _ = _pack_on_device(ops, [layer])
    g = torch.Generator().manual_seed(11)
    x = torch.randn(T, 64, generator=g).to(DEV)
    dout = torch.randn(T, 64, generator=g).to(DEV)
    plain, y = torch.empty(T, 64, device=DEV), torch.empty(T, 64, device=DEV)
    ops.tcn_layer_fwd(x, pack[0], plain, y, dilation, True)
    z = plain - x                                           # conv_1x1(y) + b_1
    dropped = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_fwd(x, pack[0], dropped, None, dilation, True, drop_p=p, seed=1234, drop_base=7 << 40)
    kept = (dropped - x).abs() > 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02, frac
    ok, err = close((dropped - x)[kept], (z / (1 - p))[kept], 1e-5); assert ok, err
    other = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_fwd(x, pack[0], other, None, dilation, True, drop_p=p, seed=1235, drop_base=7 << 40)
    assert ((other - x).abs() > 0).ne(kept).float().mean().item() > 0.3       # a different seed draws a different mask
    # backward with the same (seed, base): db_1 = sum_t dout * mask / (1 - p)
    n_slots = ops.tcn_slots(T)
    partials = torch.empty(1, n_slots, ops.TCN_GRAD_FLOATS, device=DEV)
    dpre = torch.empty(T, 64, device=DEV)
    ops.tcn_layer_bwd_hidden(dout, x, y, pack[0], dpre, partials[0], n_slots, dilation, True, drop_p=p, seed=1234,
                             drop_base=7 << 40)
    db1 = ops.tcn_reduce_grads(partials, 1, n_slots)[0][ref.WD + ref.W1 + 64:]
    ok, err = close(db1, (dout * kept / (1 - p)).sum(0), 1e-4); assert ok, err


@pytest.mark.parametrize("C", [2, 6])
def test_stage_end_kernels(ops, C):
    T = 301
    g = torch.Generator().manual_seed(C)
    x, w, b = torch.randn(T, 64, generator=g), torch.randn(C, 64, generator=g) * 0.2, torch.randn(C, generator=g)
    logits = ops.tcn_out_fwd(x.to(DEV), w.to(DEV), b.to(DEV))
    ok, err = close(logits, ref.out_fwd(x.double(), w.double(), b.double())); assert ok, err
    dl = torch.randn(C, T, generator=g)
    dx, dl_t = ops.tcn_out_bwd(dl.to(DEV), w.to(DEV))
    dx_r, dl_t_r = ref.out_bwd(dl.double(), w.double())
    ok, err = close(dx, dx_r); assert ok, err
    assert torch.equal(dl_t.cpu(), dl.t().contiguous())
    p = ops.tcn_softmax_fwd(logits)
    ok, err = close(p, ref.softmax_fwd(logits.cpu().double())); assert ok, err
    dp = torch.randn(T, C, generator=g)
    dlog = ops.tcn_softmax_bwd(p, dp.to(DEV))
    ok, err = close(dlog, ref.softmax_bwd(p.cpu().double(), dp.double())); assert ok, err


def _models(causal, f_dim, layers=8):
    from multimodal_error_detection_b200.modeling import models_TCN
    torch.manual_seed(0)
    o = nets.OracleTeCNo(2, layers, 64, f_dim, 2, causal).double()
    m = models_TCN.MultiStageModel(2, layers, 64, f_dim, 2, causal)
    m.load_state_dict({k: v.float() for k, v in o.state_dict().items()})
    nets.disable_dropout(o)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return o.train(), m.to(DEV).train()


@pytest.mark.parametrize("causal,f_dim,T", [(True, 58, 345), (False, 58, 150), (True, 2048, 410)])
def test_model_matches_oracle(causal, f_dim, T):
    o, m = _models(causal, f_dim)
    g = torch.Generator().manual_seed(T)
    base = torch.randn(1, T, f_dim, generator=g)
    xo = base.double().requires_grad_(True)
    xm = base.to(DEV).requires_grad_(True)
    out_o = o(xo.permute(0, 2, 1))
    out_m = m(xm.permute(0, 2, 1))
    assert m.impl == "b200" and tuple(out_m.shape) == tuple(out_o.shape) == (2, 1, 2, T)
    ok, err = close(out_m, out_o, 2e-5); assert ok, f"logits {err}"
    w = torch.randn(out_o.shape, generator=g)
    (out_o * w.double()).sum().backward()
    (out_m * w.to(DEV)).sum().backward()
    ok, err = close(xm.grad, xo.grad, 1e-4); assert ok, f"dx {err}"
    po, pm = dict(o.named_parameters()), dict(m.named_parameters())
    for k in po:
        ok, err = close(pm[k].grad, po[k].grad, 2e-4); assert ok, f"{k} {err}"
    # eval / no_grad form (ping-pong activations) gives the same logits
    m.eval()
    with torch.no_grad():
        out_e = m(base.to(DEV).permute(0, 2, 1))
    assert torch.equal(out_e, out_m.detach())


def test_ragged_inference_equals_per_video():
    _, m = _models(True, 58)
    m.eval()
    lengths = [300, 17, 451, 64, 5]
    g = torch.Generator().manual_seed(1)
    frames = torch.randn(sum(lengths), 58, generator=g).to(DEV)
    cat = m.forward_ragged(frames, lengths)
    parts, s = [], 0
    with torch.no_grad():
        for n in lengths:
            parts.append(m(frames[s:s + n].unsqueeze(0).permute(0, 2, 1))[:, 0])
            s += n
    ok, err = close(cat, torch.cat(parts, dim=2), 1e-6); assert ok, err


def test_train_mode_dropout_is_seeded():
    _, m = _models(True, 58, layers=4)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.5
    x = torch.randn(1, 58, 120, device=DEV)
    torch.manual_seed(7); a = m(x)
    torch.manual_seed(7); b = m(x)
    c = m(x)
    assert torch.equal(a, b) and not torch.equal(a, c)
    a.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters())


def test_frame_graph_epoch_matches_eager_epoch(fold_on_disk):
    """Frame path: train_single_epoch with every video's step replayed from its own CUDA graph (engine.FrameTrainStep) gives
    the same epochs as the eager loop -- capturing (three real steps per video, rolled back) must not disturb the training
    state -- and every video really ran from a graph."""
    import numpy as np
    import cases
    from multimodal_error_detection_b200.dataset.CustomFrameDataset import CustomFrameDataset, FrameLoader
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    path, fold = fold_on_disk
    res = {}
    for graph in (False, True):
        kw = dict(cases.FRAME_EPOCH_CASES["tecno_multimodal"], cuda_graph=graph, return_train_preds=True)
        ds = CustomFrameDataset(path, csv_filename="train.csv", delete_ND=kw["delete_ND"])
        tr = FrameLoader(ds, shuffle=True, generator=torch.Generator().manual_seed(42))
        fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device(DEV), (0.4, 0.6), 0)
        for mod in list(model.modules()) + list(fe.modules()):
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        res[graph] = [mu.train_single_epoch(model, fe, tr, crit, opt, sched, DEV, kw) for _ in range(3)]
        if graph:
            steps = opt._b200_frame_steps
            assert not steps["failed"] and len(steps["steps"]) == len(ds)
            assert all(s.graph is not None and s.launches_per_step > 50 for s in steps["steps"].values())
    for a, b in zip(res[False], res[True]):
        assert abs(a[0] - b[0]) < 1e-6 * max(1.0, abs(a[0])), (a[0], b[0])
        assert np.array_equal(a[5], b[5])
        assert a[7] == b[7] and a[8] == b[8] and a[9] == b[9]


def test_frame_graph_dropout_draws_a_fresh_mask_per_replay(fold_on_disk):
    """Dropout inside the captured step: the mask seed is a host constant plus a device counter advanced inside the graph,
    so two replays of the same video with frozen weights (lr = 0) give different losses."""
    import cases
    from multimodal_error_detection_b200.dataset.CustomFrameDataset import CustomFrameDataset, FrameLoader
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    path, fold = fold_on_disk
    kw = dict(cases.FRAME_EPOCH_CASES["tecno_multimodal"], cuda_graph=True, lr=0.0, lr_scheduler=False, weight_decay=0.0)
    ds = CustomFrameDataset(path, csv_filename="train.csv", delete_ND=kw["delete_ND"])
    tr = FrameLoader(ds, shuffle=False)
    fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device(DEV), (0.4, 0.6), 0)
    losses = [mu.train_single_epoch(model, fe, tr, crit, opt, sched, DEV, kw)[0] for _ in range(3)]
    assert not opt._b200_frame_steps["failed"]
    assert len({round(l, 7) for l in losses}) == 3, losses


@pytest.mark.parametrize("causal,lengths", [(True, [345]), (True, [300, 17, 451, 64, 5, 129]), (False, [200, 77, 130])])
def test_bf16_tcgen05_inference_matches_fp32(ops, causal, lengths):
    """Inference in the bf16 mode (dilated residual layers as tcgen05 MMAs over 128-frame tiles, taps as TMA box loads at row
    offsets, fp32 residual stream) against the fp32 SIMT kernels on the same weights: logits within the 2e-2 bar of the bf16
    mode, frame predictions equal except near ties; ragged batches never leak across video boundaries."""
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    _, m = _models(causal, 58)
    m.eval()
    g = torch.Generator().manual_seed(len(lengths))
    frames = torch.randn(sum(lengths), 58, generator=g).to(DEV)
    ref = m.forward_ragged(frames, lengths)
    m.precision = "bf16"
    out = m.forward_ragged(frames, lengths)
    m.precision = "fp32"
    assert out.shape == ref.shape and not torch.equal(out, ref)      # really another arithmetic
    for s in range(ref.shape[0]):
        ok, err = close(out[s], ref[s], 2e-2); assert ok, f"stage {s}: {err}"
    agree = (out[-1].argmax(0) == ref[-1].argmax(0)).float().mean().item()
    assert agree > 0.98, agree
    # a video's logits do not depend on its neighbours in the batch
    if len(lengths) > 1:
        m.precision = "bf16"
        alone = m.forward_ragged(frames[:lengths[0]], lengths[:1])
        m.precision = "fp32"
        ok, err = close(out[:, :, :lengths[0]], alone, 1e-6); assert ok, err


def test_frame_train_step_bf16_mode_vs_fp32_mode():
    """The frame path in the bf16 mode (exp_kwargs['precision'] = 'bf16'): the FeatureExtractor's products run on the tcgen05
    GEMMs, TeCNo trains on its fp32 kernels (one video per step is a latency chain: bf16 MMAs would not shorten it).  One
    train-mode forward / backward on the same seed-42 weights and the same video in both modes, dropout off: loss and logits
    within the bf16-mode bar (2e-2), every gradient on the fp32 gradient's side (cosine > 0.99)."""
    from multimodal_error_detection_b200 import ops
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    T = 437
    g = torch.Generator().manual_seed(9)
    images = torch.randn(1, T, 2048, generator=g).clamp_min(0).to(DEV)
    kin = torch.randn(1, T, 26, generator=g).to(DEV)
    y = (torch.rand(1, T, generator=g) > 0.5).float().to(DEV)
    out = {}
    for precision in ("fp32", "bf16"):
        kw = dict(dataset_type="frame", error_type="global", pos_weight=False, n_epochs=2, batch_size=1, lr=3e-4, lr_scheduler=True,
                  weight_decay=1e-4, num_layers=3, hidden_size=128, video_dims=32, data_type="multimodal", delete_ND=True,
                  return_train_preds=False, siamese=False, model_name="TeCNo", mstcn_stages=2, mstcn_layers=8, mstcn_f_maps=64,
                  mstcn_f_dim=58, out_features=2, mstcn_causal_conv=True, precision=precision)
        fe, model, crit, opt, _ = mu.define_model_objects(kw, {"multimodal": 58, "video": 32, "kinematics": 26}, torch.device(DEV), (0.4, 0.6))
        for mod in list(model.modules()) + list(fe.modules()):
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        model.train(); fe.train()
        logits = model(mu.define_inputs(images, kin, fe, kw, torch.device(DEV)))
        loss, _ = mu.compute_loss(logits, y, crit, "frame")
        loss.backward()
        grads = {f"fe.{k}": p.grad.detach().clone() for k, p in fe.named_parameters()}
        grads.update({f"model.{k}": p.grad.detach().clone() for k, p in model.named_parameters()})
        out[precision] = (float(loss), logits.detach().clone(), grads, getattr(fe, "precision", None))
    assert out["bf16"][3] == "bf16" and out["fp32"][3] == "fp32"
    l32, l16 = out["fp32"][0], out["bf16"][0]
    assert abs(l16 - l32) <= 2e-2 * abs(l32), (l16, l32)
    ok, err = close(out["bf16"][1], out["fp32"][1], 2e-2); assert ok, err
    assert not torch.equal(out["bf16"][1], out["fp32"][1])          # really another arithmetic
    for k, g32 in out["fp32"][2].items():
        g16 = out["bf16"][2][k]
        if float(g32.norm()) < 1e-12:
            continue
        cos = float((g16.double() * g32.double()).sum() / (g16.double().norm() * g32.double().norm()))
        assert cos > 0.99, (k, cos)
