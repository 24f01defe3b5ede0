"""north_star's "identical to 3 decimals" bar, on FIXED weights: the reference trained these weights (dropout on) and its own
validate_single_epoch / _ES / _Sequential produced the golden outputs (tests/golden/fixed_weights.*, make_golden.py).  The
CUDA path loads the same weights and must give IDENTICAL predictions in fp32 -- hence identical confusion counts and F1 /
accuracy / Jaccard -- and a ROC AUC equal to 3 decimals; in the bf16 throughput mode every prediction that differs must sit on
the decision boundary (|p - 0.5| small), and the scores move by no more than those flips explain."""
import json
import os

import numpy as np
import pytest
import torch

import cases
import fixed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fold_path(tmp_path_factory):
    return fixed.write_fold(tmp_path_factory.mktemp("fixed"))


_COUNTS = {}


def _train_counts(name, fold_path):
    """Class balance of the TRAINING windows the reference built its criterion from (pos_weight = c0 / c1)."""
    if name not in _COUNTS:
        from multimodal_error_detection_b200.dataset import dataset_utils as du
        Wt, St = cases.FIXED_TRAIN_WS[name]
        tr, _ = du.retrieve_dataloaders_window(fold_path, cases.FIXED_CASES[name][0], window_size=Wt, stride=St)
        _COUNTS[name] = tr.dataset.binary_error_distribution
    return _COUNTS[name]


def _build(name, W, precision="fp32", fold_path=None):
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    kw = dict(cases.FIXED_CASES[name][0], precision=precision)
    counts = _train_counts(name, fold_path) if fold_path is not None else (0.4, 0.6)
    fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device(DEV), counts, W)
    fixed.load_trained(name, fe, model)
    return mu, kw, fe, model, crit


def _test_loader(fold_path, kw, W, S):
    from multimodal_error_detection_b200.dataset import dataset_utils as du
    _, te = du.retrieve_dataloaders_window(fold_path, kw, window_size=W, stride=S)
    return te


@pytest.mark.parametrize("name", ["lstm_global", "cnn_global"])
def test_validate_fixed_weights_fp32(name, fold_path):
    for (W, S) in cases.FIXED_CASES[name][1]:
        mu, kw, fe, model, crit = _build(name, W, fold_path=fold_path)
        te = _test_loader(fold_path, kw, W, S)
        gold = fixed.meta()[name]["val"][f"w{W}_s{S}"]
        v = mu.validate_single_epoch(model, fe, te, crit, DEV, kw)
        assert len(te.dataset) == gold["n_test"]
        assert v[10] == gold["labels"] and list(v[12]) == gold["subjects"]
        assert v[7] == gold["preds"], f"{int(np.sum(np.asarray(v[7]) != np.asarray(gold['preds'])))} predictions differ"
        assert np.abs(np.asarray(v[8]) - np.asarray(gold["probs"])).max() < 2e-5
        assert abs(v[0] - gold["scores"][0]) <= 1e-5 * abs(gold["scores"][0])
        # identical predictions -> identical counts -> the scores agree far beyond 3 decimals
        assert np.abs(np.asarray(v[1:5]) - np.asarray(gold["scores"][1:])).max() < 1e-9
        assert np.array_equal(np.asarray(v[5]), np.asarray(gold["cm"]))
        auc = mu.roc_auc_score(v[10], v[8])
        assert abs(auc - gold["auc"]) < 5e-4, (auc, gold["auc"])                        # 3 decimals
        assert abs(mu.roc_auc_score(gold["labels"], gold["probs"]) - gold["auc"]) < 1e-12   # the kernel alone, same inputs


@pytest.mark.parametrize("name", ["lstm_global", "cnn_global"])
def test_validate_fixed_weights_bf16(name, fold_path):
    """Throughput mode (bf16 tcgen05 FeatureExtractor; LSTM on the persistent tcgen05 recurrence): probabilities within 2e-2,
    every flipped prediction documented as a boundary case, scores within what the flips explain, AUC to 3 decimals."""
    from multimodal_error_detection_b200 import ops
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    for (W, S) in cases.FIXED_CASES[name][1]:
        mu, kw, fe, model, crit = _build(name, W, "bf16", fold_path=fold_path)
        te = _test_loader(fold_path, kw, W, S)
        gold = fixed.meta()[name]["val"][f"w{W}_s{S}"]
        v = mu.validate_single_epoch(model, fe, te, crit, DEV, kw)
        probs, gp = np.asarray(v[8]), np.asarray(gold["probs"])
        assert np.abs(probs - gp).max() < 2e-2, np.abs(probs - gp).max()
        flips = np.flatnonzero(np.asarray(v[7]) != np.asarray(gold["preds"]))
        n = gold["n_test"]
        print(f"{name} W={W}: {len(flips)} of {n} predictions flip under bf16; max |dp| = {np.abs(probs - gp).max():.2e}; "
              f"flipped windows {flips.tolist()} with reference probabilities {gp[flips].round(4).tolist()}")
        # only boundary cases may flip: a window whose reference probability is within the bf16 probability error of 0.5
        assert np.all(np.abs(gp[flips] - 0.5) < 2e-2)
        assert len(flips) <= int(np.sum(np.abs(gp - 0.5) < 2e-2))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bf16_flips.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "W": W, "n_test": n, "flipped_windows": flips.tolist(),
                                "reference_probs": gp[flips].tolist(), "bf16_probs": probs[flips].tolist(),
                                "max_abs_dprob": float(np.abs(probs - gp).max()), "scores_bf16": [float(x) for x in v[1:5]],
                                "scores_reference": gold["scores"][1:]}) + "\n")
        bound = 5e-4 if len(flips) == 0 else 2.5 * len(flips) / n
        assert np.abs(np.asarray(v[1:5]) - np.asarray(gold["scores"][1:])).max() <= bound
        auc = mu.roc_auc_score(v[10], v[8])
        assert abs(auc - gold["auc"]) < 1e-3, (auc, gold["auc"])


def test_validate_es_and_cascade_fixed_weights_fp32(fold_path, tmp_path, monkeypatch):
    """validate_single_epoch_ES and validate_single_epoch_Sequential (modeling_utils.py:793-1053) on fixed weights: EXACT
    prediction lists (row a15: the cascade gated by load_binary_model_local's frozen binary model)."""
    mu, kw, fe, model, crit = _build("lstm_es", 10)
    te = _test_loader(fold_path, kw, 10, 6)
    gold = fixed.meta()["lstm_es"]["val"]["w10_s6"]
    v = mu.validate_single_epoch_ES(model, fe, te, crit, DEV, kw)
    assert v[11] == gold["preds"] and v[12] == gold["labels"] and v[13] == gold["labels_binary"] and v[14] == gold["preds_binary"]
    assert abs(v[0] - gold["scores"][0]) <= 1e-5 * abs(gold["scores"][0])
    assert np.abs(np.asarray(v[1:7]) - np.asarray(gold["scores"][1:])).max() < 1e-9
    assert np.array_equal(np.asarray(v[7]), np.asarray(gold["cm_binary"])) and np.array_equal(np.asarray(v[8]), np.asarray(gold["cm_macro"]))
    assert np.abs(np.asarray(v[10]) - np.asarray(gold["probs"])).max() < 2e-5
    # cascade: the binary model comes back through save_model -> load_binary_model_local (reference file layout)
    _, kwb, bfe, bmodel, _ = _build("lstm_global", 10)
    kwb = dict(kwb, frequency=5)
    folder = tmp_path / "models" / kwb["data_type"] / "5Hz" / "binary_lstm"
    folder.mkdir(parents=True)
    mu.save_model({"feature_extractor": bfe.state_dict(), "model": bmodel.state_dict()}, str(folder / "best_model_LOSO_1out.pt"))
    monkeypatch.chdir(tmp_path)
    models, fes = mu.load_binary_model_local("ignored", "binary_lstm", ["1out"], kwb, DEV)
    _, kws, sfe, smodel, _ = _build("lstm_seq", 10)
    te = _test_loader(fold_path, kws, 10, 6)
    gold = fixed.meta()["lstm_seq"]["val"]["w10_s6"]
    v = mu.validate_single_epoch_Sequential(smodel, sfe, models["1out"], fes["1out"], te, DEV, kws)
    assert v[12] == gold["preds_all"] and v[15] == gold["labels_all"]
    assert v[13] == gold["preds_specific"] and v[16] == gold["labels_specific"]
    assert abs(v[0] - gold["scores"][0]) <= 2e-5 * abs(gold["scores"][0])
    assert np.abs(np.asarray(v[1:9]) - np.asarray(gold["scores"][1:])).max() < 1e-9
    assert np.array_equal(np.asarray(v[9]), np.asarray(gold["cm_all"])) and np.array_equal(np.asarray(v[10]), np.asarray(gold["cm_specific"]))
    # load_model_local (reference signature) reads the same file layout
    mu.save_model({"feature_extractor": sfe.state_dict(), "model": smodel.state_dict()}, str(tmp_path / "best_model_LOSO_1out.pt"))
    fe2, model2 = mu.load_model_local(str(tmp_path), "1out", "LOSO", kws, 58, 10, DEV)
    from hashing import state_digest
    assert state_digest(fixed.cpu_sd(model2)) == fixed.meta()["lstm_seq"]["model_trained_sd"]
    assert state_digest(fixed.cpu_sd(fe2)) == fixed.meta()["lstm_seq"]["fe_trained_sd"]


def test_roc_auc_kernel_vs_sklearn():
    """Device AUC against sklearn on random scores with heavy ties, single-class error, and sizes across the sort's
    shared / global stage boundary."""
    from sklearn.metrics import roc_auc_score
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    rng = np.random.Generator(np.random.PCG64(3))
    for n in (2, 17, 2048, 2049, 5000, 70_000, 300_001):
        y = (rng.random(n) > 0.37).astype(np.float32)
        y[0], y[1] = 0.0, 1.0
        for scores in (rng.random(n).astype(np.float32), np.round(rng.random(n) * 20).astype(np.float32) / 20,
                       rng.standard_normal(n).astype(np.float32), np.zeros(n, dtype=np.float32)):
            want = roc_auc_score(y, scores.astype(np.float64))
            got = mu.roc_auc_score(y, scores)
            assert abs(got - want) < 1e-12, (n, got, want)
    with pytest.raises(ValueError):
        mu.roc_auc_score(np.ones(10, dtype=np.float32), rng.random(10).astype(np.float32))
