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


# ------------------------------------------------------------------------------------------- fixed trained weights
def _oracle_loaders(path, kw, W, S):
    """The oracle's data path on a fold written to disk: window_data -> powerset -> Needle-Drop deletion -> dataset."""
    import pandas as pd
    from multimodal_error_detection_b200 import synthetic
    fold = synthetic.make_fold(**cases.FIXED_FOLD_ARGS)
    out = []
    for trials in (fold.train, fold.test):
        image, kin, g, e5, names, _ = synthetic.flat_tables(trials)
        iw, kw_, gw, ew, sw = window_index.window_data(image, kin, g, e5, names, W, S)
        e7, nd = window_index.powerset_error_labels(ew, kw["delete_ND"])
        keep = ~nd if kw["delete_ND"] else np.ones(len(nd), dtype=bool)
        stats = {"image": {"mean": torch.from_numpy(fold.mean_image), "std": torch.from_numpy(fold.std_image)},
                 "kinematics": {"mean": torch.from_numpy(fold.mean_kin), "std": torch.from_numpy(fold.std_kin)}}
        out.append(loops.OracleWindowDataset(torch.from_numpy(iw[keep]), torch.from_numpy(kw_[keep]), torch.from_numpy(gw[keep]),
                                             torch.from_numpy(e7[keep]), [s for s, k in zip(sw, keep) if k], stats))
    return loops.make_loaders(out[0], out[1], kw["batch_size"])


def _oracle_fixed(name, W):
    import fixed
    kw = cases.FIXED_CASES[name][0]
    fe, model, crit, _, _ = nets.build_objects(kw, cases.IN_FEATURES, (0.4, 0.6), W)
    fixed.load_trained(name, fe, model)
    return kw, fe, model, crit


@pytest.mark.parametrize("name", ["lstm_global", "cnn_global"])
def test_fixed_weights_validation_binary(name):
    """The reference's validate_single_epoch on reference-trained weights (tests/golden/fixed_weights.*): the oracle's
    predictions are identical, its pooled scores and roc_auc_score equal to 3 decimals (north_star bar) -- in fact to 1e-12."""
    import fixed
    from sklearn.metrics import roc_auc_score
    for (W, S) in cases.FIXED_CASES[name][1]:
        kw, fe, model, crit = _oracle_fixed(name, W)
        _, te = _oracle_loaders(None, kw, W, S)
        gold = fixed.meta()[name]["val"][f"w{W}_s{S}"]
        v = loops.validate_epoch(model, fe, te, crit, kw)
        assert len(te.dataset) == gold["n_test"]
        assert v[6] == gold["preds"] and v[8] == gold["labels"]
        assert np.abs(np.asarray(v[7]) - np.asarray(gold["probs"])).max() < 1e-6
        assert np.abs(np.asarray(v[1:5]) - np.asarray(gold["scores"][1:])).max() < 1e-12
        assert abs(roc_auc_score(v[8], v[7]) - gold["auc"]) < 1e-9


def test_fixed_weights_validation_es_and_cascade():
    import fixed
    kw, fe, model, crit = _oracle_fixed("lstm_es", 10)
    _, te = _oracle_loaders(None, kw, 10, 6)
    gold = fixed.meta()["lstm_es"]["val"]["w10_s6"]
    v = loops.validate_epoch_es(model, fe, te, torch.nn.CrossEntropyLoss(), kw)
    assert v[10] == gold["preds"] and v[11] == gold["labels"]
    assert abs(v[0] - gold["scores"][0]) < 1e-5 * abs(gold["scores"][0])
    assert np.abs(np.asarray(v[1:7]) - np.asarray(gold["scores"][1:])).max() < 1e-12
    assert np.abs(np.asarray(v[9]) - np.asarray(gold["probs"])).max() < 1e-6
    kws, sfe, smodel, _ = _oracle_fixed("lstm_seq", 10)
    kwb, bfe, bmodel, _ = _oracle_fixed("lstm_global", 10)
    _, te = _oracle_loaders(None, kws, 10, 6)
    gold = fixed.meta()["lstm_seq"]["val"]["w10_s6"]
    loss, pa, ps, la, ls = loops.validate_epoch_sequential(smodel, sfe, bmodel, bfe, te, kws)
    assert pa == gold["preds_all"] and la == gold["labels_all"] and ps == gold["preds_specific"] and ls == gold["labels_specific"]
    assert abs(loss - gold["scores"][0]) < 1e-5 * abs(gold["scores"][0])


# ------------------------------------------------------------------------------------------- host-side f1 / f4 functions
def test_load_data_both_schemas_and_binary_mask(golden_dir, tmp_path):
    """load_data (fold schema and the video_data_path schema, dataset_utils.py:73-96 / :118-136), create_binary_mask with
    mask_position_ND_<subject>.pth files (modeling_utils.py:2920-2976) and create_summary_df (:2979-3025): host-side
    functions of the drop-in package against the outputs of the executed reference."""
    from multimodal_error_detection_b200.dataset import dataset_utils as du
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    gold = json.load(open(os.path.join(golden_dir, "postproc.json")))
    fold = synthetic.make_fold(seed=5, n_train=3, n_test=2, t_lo=60, t_hi=90)
    path, vpath = synthetic.write_fold_video_schema(fold, str(tmp_path / "fold"), str(tmp_path / "video"))
    flat = du.load_data(path + "/", "train.csv", video_data_path=vpath + "/")
    g = gold["load_data_video"]
    assert (digest(flat[0]), digest(flat[1]), digest(flat[2]), digest(flat[3])) == (g["image"], g["kin"], g["g"], g["e"])
    assert flat[4]["subject"].tolist()[::40] == g["subjects"] and flat[0].shape[0] == g["n"]
    flat0 = du.load_data(path + "/", "train.csv")
    assert digest(flat0[0]) == gold["load_data_fold"]["image"] and digest(flat0[1]) == gold["load_data_fold"]["kin"]
    assert digest(flat0[0]) != digest(flat[0])
    # create_binary_mask
    bm = gold["binary_mask"]
    for subj, mk in bm["masks"].items():
        torch.save(torch.tensor(mk), os.path.join(path, f"mask_position_ND_{subj}.pth"))
    out_mask, out_subj = mu.create_binary_mask({"t": bm["preds"]}, {"t": bm["subjects"]}, "t", path, {"delete_ND": True})
    assert out_mask.tolist() == bm["mask_out"] and out_subj.tolist() == bm["subjects_out"]
    keep, _ = mu.create_binary_mask({"t": bm["preds"]}, {"t": bm["subjects"]}, "t", path, {"delete_ND": False})
    assert keep.tolist() == bm["mask_out_keep_nd"]
    # create_summary_df
    sd = gold["summary_df"]
    i = sd["inputs"]
    df = mu.create_summary_df(*[np.asarray(l) for l in i["lists"]], np.asarray(i["samples_train"]), np.asarray(i["samples_test"]),
                              np.asarray(i["rates"]), np.asarray(i["times"]))
    for key, want in sd["cells"].items():
        r, c = key.split("/")
        got = df.loc[r, c]
        assert (want is None and isinstance(got, float) and np.isnan(got)) or str(got) == want, (key, got, want)
