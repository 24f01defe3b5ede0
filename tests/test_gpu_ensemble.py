"""GPU tests of the ensemble-inference path (multimodal_error_detection_b200/ensemble.py): ragged frame-model passes equal
the per-video loop, the frame -> window bridge equals ``window_predictions`` (pinned to the reference by
tests/golden/window_predictions.npz), the soft vote and the confusion counts equal their numpy statement."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _job(precision="fp32"):
    from multimodal_error_detection_b200 import synthetic
    from multimodal_error_detection_b200.dataset.dataset_utils import dataset_from_index
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    from multimodal_error_detection_b200.table import FrameTable
    fold = synthetic.make_fold(seed=5, n_train=7, n_test=0, t_lo=120, t_hi=400)
    image, kin, g, e5, names, offsets = synthetic.flat_tables(fold.train)
    table = FrameTable(torch.from_numpy(image), torch.from_numpy(kin), torch.from_numpy(g), torch.from_numpy(e5), names)
    stats = {"image": {"mean": torch.from_numpy(fold.mean_image), "std": torch.from_numpy(fold.std_image)},
             "kinematics": {"mean": torch.from_numpy(fold.mean_kin), "std": torch.from_numpy(fold.std_kin)}}
    ds = dataset_from_index(table.window_index(10, 6), True, stats)
    w_kw = cases.base_kwargs(model_name="SimpleLSTM", precision=precision)
    w_fe, w_model, _, _, _ = mu.define_model_objects(w_kw, cases.IN_FEATURES, torch.device(DEV), ds.binary_error_distribution, 10)
    f_kw = dict(cases.FRAME_EPOCH_CASES["tecno_multimodal"])
    f_fe, f_model, _, _, _ = mu.define_model_objects(f_kw, cases.IN_FEATURES, torch.device(DEV), (0.4, 0.6), 0)
    return table, ds, (f_fe, f_model, f_kw), (w_fe, w_model, w_kw), stats["kinematics"], (names, g)


def test_ensemble_inference_matches_per_video_and_numpy():
    from multimodal_error_detection_b200 import ensemble, ops
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    table, ds, fobj, wobj, kin_stats, (names, g) = _job()
    out = ensemble.ensemble_inference(table, ds, fobj, wobj, kin_stats, batch_size=97, frames_per_pass=700)
    f_fe, f_model, f_kw = fobj
    # (1) ragged passes (several videos per pass, 700-frame budget) == one forward per video
    off = table.offsets_host
    ref = []
    with torch.no_grad():
        for v in range(len(off) - 1):
            r0, r1 = int(off[v]), int(off[v + 1])
            kin = ops.standardise_rows(table.kin[r0:r1].contiguous(), ops.expand_stat(kin_stats["mean"], 26, 1, DEV),
                                       ops.expand_stat(kin_stats["std"], 26, 1, DEV))
            x = torch.cat([f_fe(table.image[r0:r1]).float(), kin], dim=1).unsqueeze(0).permute(0, 2, 1)
            ref.append(torch.argmax(f_model(x)[-1, 0], dim=0).float())
    ref = torch.cat(ref)
    assert (out["frame_preds"] != ref).float().mean().item() < 2e-3      # argmax ties only (1e-6-level logit differences)
    # (2) window model probabilities == the eager validate loop's
    w_fe, w_model, w_kw = wobj
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader
    crit = mu.FusedBCEWithLogitsLoss()
    v = mu.validate_single_epoch(w_model, w_fe, DeviceWindowLoader(ds, 64, shuffle=False), crit, DEV, w_kw)
    assert np.abs(np.asarray(v[8]).reshape(-1) - out["window_probs"].cpu().numpy()).max() < 1e-5
    # (3) the bridge: window vote of the frame predictions over the dataset's own index
    fp = out["frame_preds"].cpu().numpy()
    starts = ds.index.starts.cpu().numpy()
    fw = np.asarray([fp[s:s + 10].astype(np.float64).mean() >= 0.5 for s in starts], dtype=np.float32)
    assert np.array_equal(out["frame_windows"].cpu().numpy(), fw)
    # (4) soft vote (ensemble.ipynb cell 6) + confusion counts
    wp = out["window_probs"].cpu().numpy().astype(np.float64)
    fused = ((wp + fw) / 2 >= 0.5).astype(np.float32)
    assert np.array_equal(out["fused"].cpu().numpy(), fused)
    y = out["labels"].cpu().numpy()
    tn, fp_, fn, tp = [int(((y == a) & (fused == b)).sum()) for a, b in ((0, 0), (0, 1), (1, 0), (1, 1))]
    assert out["counts"].tolist() == [tn, fp_, fn, tp]


def test_postprocessing_vs_executed_reference(golden_dir):
    """frame2window / compute_window_metrics (binary and multi-class), soft_vote_ensemble, cascade_ensemble against the
    outputs of the executed reference (tests/golden/postproc.json, make_golden.py::gen_postproc)."""
    import json
    import os
    import cases
    from hashing import digest
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    gold = json.load(open(os.path.join(golden_dir, "postproc.json")))
    outs, preds_b, preds_m, labels, gests, subjects = cases.postproc_inputs()
    for tag, preds, binary in (("binary", preds_b, True), ("multi", preds_m, False)):
        g = gold[f"window_metrics_{tag}"]
        wp, wl, wg, ws = mu.frame2window(outs, preds, labels, gests, subjects, window_size=10, stride=6, binary=binary)
        assert {o: int(len(wp[o])) for o in wp} == g["n_windows"]
        for o in wp:
            assert digest(wp[o].numpy()) == g["preds_digest"][o] and digest(wl[o].numpy()) == g["labels_digest"][o]
            assert ws[o]["subject"].tolist()[:3] == g["first_subjects"][o]
        df, cm = mu.compute_window_metrics(outs, preds, labels, gests, subjects, window_size=10, stride=6, binary=binary)
        assert {c: df.loc["Windowed Metrics", c] for c in df.columns} == g["summary"]
        assert np.asarray(cm).tolist() == g["cm"]
    z = np.load(os.path.join(golden_dir, "postproc_inputs.npz"))
    g = gold["soft_vote"]
    preds, acc, f1, jac, cm = mu.soft_vote_ensemble(z["pa"], z["pb"], z["lab"])
    assert digest(preds.astype(np.int64)) == g["preds_digest"] and cm.tolist() == g["cm"]
    assert abs(acc - g["acc"]) < 1e-12 and abs(f1 - g["f1"]) < 1e-12 and abs(jac - g["jaccard"]) < 1e-12
    z = np.load(os.path.join(golden_dir, "postproc_cascade.npz"))
    assert np.array_equal(mu.cascade_ensemble(z["b"], z["m"]).astype(np.int64), z["ens"])


def test_window_model_probabilities_fused_route_bf16():
    """ensemble.window_model_probabilities in the configuration of BASELINE configs[4] (W = 16, bf16 FE + LSTM): the gather runs
    inside the first FeatureExtractor layer (no bf16 batch written) and inside the recurrence's first operand; the
    probabilities must agree with the public validate loop (gather_batch + define_inputs on the same weights) to bf16 noise."""
    from multimodal_error_detection_b200 import ensemble, ops, synthetic
    from multimodal_error_detection_b200.dataset.CustomWindowDataset import DeviceWindowLoader
    from multimodal_error_detection_b200.dataset.dataset_utils import dataset_from_index
    from multimodal_error_detection_b200.modeling import modeling_utils as mu
    from multimodal_error_detection_b200.table import FrameTable
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    fold = synthetic.make_fold(seed=6, n_train=6, n_test=0, t_lo=150, t_hi=400)
    image, kin, g, e5, names, offsets = synthetic.flat_tables(fold.train)
    table = FrameTable(torch.from_numpy(image), torch.from_numpy(kin), torch.from_numpy(g), torch.from_numpy(e5), names)
    stats = {"image": {"mean": torch.from_numpy(fold.mean_image), "std": torch.from_numpy(fold.std_image)},
             "kinematics": {"mean": torch.from_numpy(fold.mean_kin), "std": torch.from_numpy(fold.std_kin)}}
    ds = dataset_from_index(table.window_index(16, 4), True, stats)
    kw = cases.base_kwargs(model_name="SimpleLSTM", precision="bf16")
    fe, model, _, _, _ = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device(DEV), ds.binary_error_distribution, 16)
    n0 = ops._lib.launch_count()
    probs = ensemble.window_model_probabilities(ds, fe, model, kw, batch_size=150)
    assert ops._lib.launch_count() > n0 and probs.shape == (len(ds),)
    v = mu.validate_single_epoch(model, fe, DeviceWindowLoader(ds, 64, shuffle=False), mu.FusedBCEWithLogitsLoss(), DEV, kw)
    ref = np.asarray(v[8]).reshape(-1)
    assert np.abs(ref - probs.cpu().numpy()).max() < 5e-3, float(np.abs(ref - probs.cpu().numpy()).max())


@pytest.mark.parametrize("n", [128, 1000, 4097])
def test_frame_features_fused_rows_match_the_plain_forward(n):
    """ensemble.frame_features (bf16 inference: table rows -> fused gather + first-layer kernel in blocks of 128 consecutive
    frames, ragged tail as one overlapping block) against FeatureExtractor.forward on the same rows: same bf16 operands and
    fp32 accumulation, so the features agree to bf16 rounding of the hidden activations."""
    from multimodal_error_detection_b200 import ensemble, ops
    from multimodal_error_detection_b200.modeling.models import FeatureExtractor
    if not ops.has_tcgen05():
        pytest.skip("needs sm_100")
    torch.manual_seed(1)
    fe = FeatureExtractor(2048, 32, [512, 256], precision="bf16").to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(n)
    table = torch.randn(n + 300, 2048, device=DEV, generator=g).clamp_min_(0)
    r0 = 137
    with torch.no_grad():
        n0 = ops._lib.launch_count()
        got = ensemble.frame_features(fe, table, r0, r0 + n)
        assert ops._lib.launch_count() > n0
        want = fe(table[r0:r0 + n]).float()
    assert got.shape == want.shape == (n, 32)
    err = float((got - want).norm() / want.norm())
    assert err < 5e-3, err
