"""Pin the oracle: every oracle function against the fixtures produced by executing the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
from hashing import digest, state_digest
from multimodal_error_detection_b200 import synthetic
from oracle import loops, nets, window_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def c_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "c", "libmed_oracle.so"))
    lib.med_oracle_window_starts.restype = ctypes.c_int64
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _names(offsets, reverse=False):
    nv = len(offsets) - 1
    return np.concatenate([[synthetic.trial_name(nv - 1 - i if reverse else i)] * int(offsets[i + 1] - offsets[i])
                           for i in range(nv)])


@pytest.mark.parametrize("case", cases.WINDOW_CASES, ids=[c[0] for c in cases.WINDOW_CASES])
def test_window_starts(case, golden_dir, c_oracle):
    name, seed, nv, lo, hi, W, S = case
    gold = np.load(os.path.join(golden_dir, "window_index.npz"))
    g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
    rows, subj = window_index.window_starts(g, _names(offsets), W, S)
    assert np.array_equal(rows[:, 0], gold[f"{name}/starts"])
    assert np.array_equal(rows, rows[:, :1] + np.arange(W)[None])
    assert subj == gold[f"{name}/subj_win"].tolist()
    _, _, gw, ew, _ = window_index.window_data(np.zeros((len(g), 1)), np.zeros((len(g), 1)), g, e5, _names(offsets), W, S)
    assert np.array_equal(gw, gold[f"{name}/g_win"]) and np.array_equal(ew, gold[f"{name}/e_win"])
    # C restatement
    cap = len(g)
    out = np.zeros(cap, dtype=np.int64)
    n = c_oracle.med_oracle_window_starts(_ptr(g), _ptr(offsets), ctypes.c_int64(nv), ctypes.c_int64(W),
                                          ctypes.c_int64(S), _ptr(out), ctypes.c_int64(cap))
    assert n == len(gold[f"{name}/starts"]) and np.array_equal(out[:n], gold[f"{name}/starts"])


@pytest.mark.parametrize("delete_nd", [True, False])
def test_powerset(delete_nd, golden_dir, c_oracle):
    gold = np.load(os.path.join(golden_dir, "powerset.npz"))
    rows = cases.all_label_rows()
    assert np.array_equal(rows, gold["rows"])
    e7, mask = window_index.powerset_error_labels(rows, delete_nd)
    assert e7.dtype == np.int32
    assert np.array_equal(e7, gold[f"e7_{int(delete_nd)}"]) and np.array_equal(mask, gold[f"mask_{int(delete_nd)}"])
    e7c = np.zeros((len(rows), 7), dtype=np.int32)
    mc = np.zeros(len(rows), dtype=np.uint8)
    c_oracle.med_oracle_powerset(_ptr(rows), ctypes.c_int64(len(rows)), ctypes.c_int(int(delete_nd)), _ptr(e7c), _ptr(mc))
    assert np.array_equal(e7c, e7) and np.array_equal(mc.astype(bool), mask)


@pytest.mark.parametrize("case", cases.WINDOW_CASES[:3], ids=[c[0] for c in cases.WINDOW_CASES[:3]])
def test_window_predictions(case, golden_dir):
    name, seed, nv, lo, hi, W, S = case
    gold = np.load(os.path.join(golden_dir, "window_predictions.npz"))
    g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
    names = _names(offsets, reverse=True)
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    pb = (rng.random(len(g)) > 0.5).astype(np.float64)
    pm = rng.integers(0, 6, len(g)).astype(np.float64)
    for tag, p, binary in (("bin", pb, True), ("multi", pm, False)):
        pw, ew, gw, sw = window_index.window_predictions(p, e5[:, 4].astype(np.float64), g.astype(np.float64), names, W, S, binary)
        assert np.array_equal(pw, gold[f"{name}/{tag}/preds"])
        assert np.array_equal(ew, gold[f"{name}/{tag}/labels"])
        assert np.array_equal(gw, gold[f"{name}/{tag}/gest"])
        assert sw == gold[f"{name}/{tag}/subj"].tolist()


@pytest.mark.parametrize("name", list(cases.MODEL_CASES))
def test_models(name, golden_dir):
    """Seed-42 construction gives the reference's weights bit-for-bit; logits, loss, gradients and one
    Adam + cosine step agree (same torch ops on the same CPU -> compared tightly)."""
    meta = json.load(open(os.path.join(golden_dir, "models.json")))[name]
    gold = np.load(os.path.join(golden_dir, "models.npz"))
    kw, W, counts = cases.MODEL_CASES[name]
    fe, model, crit, opt, sched = nets.build_objects(kw, cases.IN_FEATURES, counts, W)
    assert list(model.state_dict().keys()) == meta["model_keys"]
    assert list(fe.state_dict().keys()) == meta["fe_keys"]
    assert state_digest(model.state_dict()) == meta["model_sd"]
    assert state_digest(fe.state_dict()) == meta["fe_sd"]
    n_params = sum(p.numel() for p in model.parameters()) + sum(p.numel() for p in fe.parameters())
    assert n_params == meta["n_params"]
    images, kin, y = cases.model_inputs(name)
    model.eval(); fe.eval()
    with torch.no_grad():
        logits = model(loops.fuse_inputs(images, kin, fe, kw))
    np.testing.assert_allclose(logits.numpy(), gold[f"{name}/logits_eval"], rtol=1e-6, atol=1e-6)
    nets.disable_dropout(model, fe)
    model.train(); fe.train()
    out = model(loops.fuse_inputs(images, kin, fe, kw))
    if kw["dataset_type"] == "window" and kw["error_type"] == "all_errors":
        loss = crit(out, y.long())
    else:
        loss, _ = loops.loss_fn(out, y, crit, kw["dataset_type"])
    opt.zero_grad(); loss.backward()
    np.testing.assert_allclose(loss.item(), gold[f"{name}/loss"], rtol=1e-6)
    norms = []
    for prefix, mod in (("fe", fe), ("model", model)):
        for k, p in mod.named_parameters():
            norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
            if p.grad is not None:
                np.testing.assert_allclose(p.grad.reshape(-1)[:16].numpy(), gold[f"{name}/grad/{prefix}.{k}"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(norms, gold[f"{name}/grad_norms"], rtol=1e-5, atol=1e-9)
    opt.step(); sched.step()
    assert abs(opt.param_groups[0]["lr"] - meta["lr_after_sched"]) < 1e-12
    np.testing.assert_allclose(next(fe.parameters()).detach().reshape(-1)[:64].numpy(), gold[f"{name}/fe_w0_after_step"], rtol=1e-6, atol=1e-8)


def _oracle_window_loaders(fold, kw, W, S):
    stats = {"image": {"mean": torch.from_numpy(fold.mean_image), "std": torch.from_numpy(fold.std_image)},
             "kinematics": {"mean": torch.from_numpy(fold.mean_kin), "std": torch.from_numpy(fold.std_kin)}}
    dss = []
    for trials in (fold.train, fold.test):
        image, kin, g, e5, names, _ = synthetic.flat_tables(trials)
        iw, kw_, gw, ew, sw = window_index.window_data(image, kin, g, e5, names, W, S)
        e7, mask = window_index.powerset_error_labels(ew, kw["delete_ND"])
        if kw["delete_ND"]:
            keep = ~mask
            iw, kw_, gw, e7, sw = iw[keep], kw_[keep], gw[keep], e7[keep], [s for s, k in zip(sw, keep) if k]
        dss.append(loops.OracleWindowDataset(torch.from_numpy(iw), torch.from_numpy(kw_), torch.from_numpy(gw),
                                             torch.from_numpy(e7), sw, stats))
    return loops.make_loaders(dss[0], dss[1], kw["batch_size"])


@pytest.mark.parametrize("name", list(cases.EPOCH_CASES))
def test_window_epochs(name, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "epochs.json")))[name]
    kw, W, S = cases.EPOCH_CASES[name]
    fold = synthetic.make_fold(**cases.FOLD_ARGS)
    tr, te = _oracle_window_loaders(fold, kw, W, S)
    ds = tr.dataset
    assert len(ds) == gold["n_train"] and len(te.dataset) == gold["n_test"]
    np.testing.assert_allclose([float(v) for v in ds.binary_error_distribution], gold["binary_error_distribution"], rtol=1e-7)
    np.testing.assert_allclose(ds.specific_error_distribution, gold["specific_error_distribution"], rtol=1e-6)
    item = ds[3]
    assert digest(item[0]) == gold["item3_image_digest"] and digest(item[1]) == gold["item3_kin_digest"]
    assert item[3].tolist() == gold["item3_e7"] and item[4] == gold["item3_subject"]
    first = next(iter(torch.utils.data.DataLoader(ds, batch_size=kw["batch_size"], shuffle=True,
                                                  generator=torch.Generator().manual_seed(42))))
    assert digest(first[3]) == gold["first_batch_e7_digest"] and digest(first[0]) == gold["first_batch_image_digest"]
    fe, model, crit, opt, sched = nets.build_objects(kw, cases.IN_FEATURES, ds.binary_error_distribution, W)
    nets.disable_dropout(model, fe)
    for ep in range(kw["n_epochs"]):
        t = loops.train_epoch(model, fe, tr, crit, opt, sched, kw)
        v = loops.validate_epoch(model, fe, te, crit, kw)
        g = gold["epochs"][ep]
        np.testing.assert_allclose(t[:5], g["train"], rtol=2e-4, atol=1e-5)
        assert t[5].tolist() == g["train_cm"]
        np.testing.assert_allclose(v[:5], g["val"], rtol=2e-4, atol=1e-5)
        assert v[5].tolist() == g["val_cm"] and v[6] == g["val_preds"]


def test_c_gather_norm(c_oracle):
    rng = np.random.Generator(np.random.PCG64(3))
    table = rng.standard_normal((200, 40), dtype=np.float32)
    mean = rng.standard_normal(40, dtype=np.float32)
    std = (rng.random(40, dtype=np.float32) + 0.5)
    starts = np.asarray([0, 7, 190, 55], dtype=np.int64)
    out = np.zeros((4, 10, 40), dtype=np.float32)
    c_oracle.med_oracle_gather_norm(_ptr(table), ctypes.c_int64(40), _ptr(mean), _ptr(std), _ptr(starts),
                                    ctypes.c_int64(4), ctypes.c_int64(10), _ptr(out))
    want = ((torch.from_numpy(table)[torch.from_numpy(starts)[:, None] + torch.arange(10)[None]] - torch.from_numpy(mean))
            / torch.from_numpy(std)).numpy()
    assert np.array_equal(out, want)
