"""Generate the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Writes small ``.npz`` / ``.json`` files next to this script.  The reference has no tests or
fixtures of its own, so these are the pins for oracle/ and for the CUDA path.
Dropout is switched off on the returned modules (attribute writes only -- no reference source is
touched) wherever train-mode numbers are recorded, because CPU dropout streams cannot be
replayed on the GPU.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from multimodal_error_detection_b200 import synthetic  # noqa: E402
import cases  # noqa: E402
from hashing import digest, state_digest  # noqa: E402
from refharness import import_reference  # noqa: E402

import pandas as pd  # noqa: E402

ref = import_reference()
quiet = contextlib.redirect_stdout(io.StringIO())


def no_dropout(*mods):
    for mod in mods:
        if mod is None:
            continue
        for m in mod.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.LSTM):
                m.dropout = 0.0


# ---------------------------------------------------------------- window index (a1)
def gen_window_index():
    out = {}
    for name, seed, nv, lo, hi, W, S in cases.WINDOW_CASES:
        g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
        N = len(g)
        names = np.concatenate([[synthetic.trial_name(i)] * int(offsets[i + 1] - offsets[i]) for i in range(nv)])
        carrier = torch.arange(N, dtype=torch.float64).reshape(N, 1)  # index-carrier trick, SURVEY A-1
        kin = torch.zeros(N, 1)
        res = ref.dataset_utils.window_data(carrier, kin, torch.from_numpy(g).reshape(-1, 1),
                                            torch.from_numpy(e5), pd.DataFrame({"subject": names}),
                                            window_size=W, stride=S)
        starts = res[0][:, 0, 0].long().numpy()
        assert (res[0][:, :, 0].long().numpy() == starts[:, None] + np.arange(W)[None]).all()
        out[f"{name}/starts"] = starts
        out[f"{name}/g_win"] = res[2].numpy()
        out[f"{name}/e_win"] = res[3].numpy()
        out[f"{name}/subj_win"] = np.asarray(res[4]["subject"].tolist())
    np.savez_compressed(os.path.join(HERE, "window_index.npz"), **out)
    print("window_index:", {k: v.shape for k, v in out.items() if k.endswith("starts")})


# ---------------------------------------------------------------- powerset labels (a2)
def gen_powerset():
    rows = cases.all_label_rows()
    out = {"rows": rows}
    for flag in (True, False):
        with quiet:
            e7, mask = ref.dataset_utils.powerset_error_labels(torch.from_numpy(rows), delete_ND=flag)
        out[f"e7_{int(flag)}"] = e7.numpy()
        out[f"mask_{int(flag)}"] = mask.numpy()
        assert e7.dtype == torch.int32
    # the duplicate inside CustomFrameDataset (CustomFrameDataset.py:162-247)
    ds = ref.CustomFrameDataset.CustomFrameDataset.__new__(ref.CustomFrameDataset.CustomFrameDataset)
    with quiet:
        e7f, maskf = ds.powerset_error_labels(torch.from_numpy(rows), delete_ND=True)
    assert (e7f.numpy() == out["e7_1"]).all() and (maskf.numpy() == out["mask_1"]).all()
    np.savez_compressed(os.path.join(HERE, "powerset.npz"), **out)
    print("powerset:", out["e7_1"].shape)


# ---------------------------------------------------------------- models (a7-a11)
def gen_models():
    meta, arrays = {}, {}
    mu = ref.modeling_utils
    for name, (kw, W, counts) in cases.MODEL_CASES.items():
        with quiet:
            fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device("cpu"), counts, W)
        n_params = sum(p.numel() for p in model.parameters()) + (sum(p.numel() for p in fe.parameters()) if fe else 0)
        rec = {"n_params": n_params, "model_sd": state_digest(model.state_dict()),
               "fe_sd": state_digest(fe.state_dict()) if fe is not None else None,
               "model_keys": list(model.state_dict().keys()),
               "fe_keys": list(fe.state_dict().keys()) if fe is not None else None}
        images, kin, y = cases.model_inputs(name)
        # eval-mode logits
        model.eval(); fe.eval()
        with torch.no_grad():
            inputs = mu.define_inputs(images, kin, fe, kw, torch.device("cpu"))
            logits = model(inputs)
        arrays[f"{name}/inputs_digest"] = np.frombuffer(bytes.fromhex(digest(inputs)), dtype=np.uint8)
        arrays[f"{name}/fe_out"] = fe(images).detach().numpy() if kw["data_type"] != "video" or kw["video_dims"] != 2048 else np.zeros(1)
        arrays[f"{name}/logits_eval"] = logits.numpy()
        # train-mode loss + gradients, dropout off
        no_dropout(model, fe)
        model.train(); fe.train()
        inputs = mu.define_inputs(images, kin, fe, kw, torch.device("cpu"))
        out = model(inputs)
        if kw["dataset_type"] == "window" and kw["error_type"] == "all_errors":
            loss = crit(out, y.long())   # committed code passes float -> raises; long is the restatement
        else:
            loss, _ = mu.compute_loss(out, y, crit, kw["dataset_type"])
        opt.zero_grad()
        loss.backward()
        arrays[f"{name}/logits_train"] = out.detach().numpy()
        arrays[f"{name}/loss"] = np.asarray(loss.item(), dtype=np.float64)
        gnames, gnorms = [], []
        for prefix, mod in (("fe", fe), ("model", model)):
            for k, p in mod.named_parameters():
                gnames.append(f"{prefix}.{k}")
                gnorms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
                if p.grad is not None:
                    arrays[f"{name}/grad/{prefix}.{k}"] = p.grad.reshape(-1)[:16].numpy().copy()
        rec["grad_names"] = gnames
        arrays[f"{name}/grad_norms"] = np.asarray(gnorms)
        # one Adam step (coupled L2 decay) and the scheduler step
        opt.step()
        if sched is not None:
            sched.step()
        rec["lr_after_sched"] = opt.param_groups[0]["lr"]
        first = next(fe.parameters()) if n_params and kw["data_type"] != "kinematics" else next(model.parameters())
        arrays[f"{name}/fe_w0_after_step"] = first.detach().reshape(-1)[:64].numpy().copy()
        arrays[f"{name}/head_last_after_step"] = list(model.parameters())[-2].detach().reshape(-1)[:64].numpy().copy()
        meta[name] = rec
        print("model", name, "params", n_params, "loss", float(loss))
    np.savez_compressed(os.path.join(HERE, "models.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "models.json"), "w"), indent=1)


# ---------------------------------------------------------------- epochs on an on-disk fold (a3-a6, a12-a13, a16, a17)
def gen_epochs():
    mu, du = ref.modeling_utils, ref.dataset_utils
    fold = synthetic.make_fold(**cases.FOLD_ARGS)
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        path = synthetic.write_fold(fold, os.path.join(tmp, "fold")) + "/"
        for name, (kw, W, S) in cases.EPOCH_CASES.items():
            with quiet:
                tr, te = du.retrieve_dataloaders_window(path, kw, window_size=W, stride=S)
            ds = tr.dataset
            rec = {"n_train": len(ds), "n_test": len(te.dataset),
                   "binary_error_distribution": [float(v) for v in ds.binary_error_distribution],
                   "specific_error_distribution": [float(v) for v in ds.specific_error_distribution]}
            item = ds[3]
            rec["item3_image_digest"] = digest(item[0])
            rec["item3_kin_digest"] = digest(item[1])
            rec["item3_e7"] = item[3].tolist()
            rec["item3_subject"] = item[4]
            first = next(iter(du.DataLoader(ds, batch_size=kw["batch_size"], shuffle=True,
                                            generator=torch.Generator().manual_seed(42))))
            rec["first_batch_e7_digest"] = digest(first[3])
            rec["first_batch_image_digest"] = digest(first[0])
            counts = ds.binary_error_distribution
            with quiet:
                fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device("cpu"), counts, W)
            no_dropout(model, fe)
            epochs = []
            for ep in range(kw["n_epochs"]):
                with quiet, contextlib.redirect_stderr(io.StringIO()):
                    t = mu.train_single_epoch(model, fe, tr, crit, opt, sched, torch.device("cpu"), kw)
                    v = mu.validate_single_epoch(model, fe, te, crit, torch.device("cpu"), kw)
                epochs.append({"train": [float(x) for x in t[:5]], "train_cm": t[5].tolist(),
                               "val": [float(x) for x in v[:5]], "val_cm": v[5].tolist(),
                               "val_preds": [float(x) for x in v[7]], "val_probs": [float(x) for x in v[8]],
                               "val_labels": [float(x) for x in v[10]],
                               "lr": opt.param_groups[0]["lr"]})
                if kw["return_train_preds"]:
                    epochs[-1]["train_preds"] = [float(x) for x in t[7]]
                    epochs[-1]["train_labels"] = [float(x) for x in t[8]]
                    epochs[-1]["train_subjects"] = list(t[9])
            rec["epochs"] = epochs
            rec["final_fe_w0"] = next(fe.parameters()).detach().reshape(-1)[:32].tolist()
            res[name] = rec
            print("epoch case", name, rec["n_train"], rec["n_test"], epochs[-1]["train"][:2], epochs[-1]["val"][:2])
        # frame path
        for name, kw in cases.FRAME_EPOCH_CASES.items():
            FD = ref.CustomFrameDataset.CustomFrameDataset
            with quiet:
                dtr = FD(path, csv_filename="train.csv", delete_ND=kw["delete_ND"])
                dte = FD(path, csv_filename="test.csv", delete_ND=kw["delete_ND"])
            # as in train_frame.ipynb cell 2 (lines 58-66): one fresh seed-42 generator per loader
            tr = du.DataLoader(dtr, batch_size=1, shuffle=True, generator=torch.Generator().manual_seed(42))
            te = du.DataLoader(dte, batch_size=1, shuffle=False, generator=torch.Generator().manual_seed(42))
            item = dtr[1]
            rec = {"n_train": len(dtr), "n_frames": int(dtr.get_n_frames()),
                   "item1_shapes": [list(x.shape) for x in item if isinstance(x, torch.Tensor)],
                   "item1_kin_digest": digest(item[1]), "item1_e7_digest": digest(item[3]),
                   "item1_image_digest": digest(item[0]), "item1_skill": item[5][0].tolist(),
                   "item1_subject": item[4]}
            with quiet:
                fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, torch.device("cpu"), (0.4, 0.6), 0)
            no_dropout(model, fe)
            epochs = []
            for ep in range(kw["n_epochs"]):
                with quiet, contextlib.redirect_stderr(io.StringIO()):
                    t = mu.train_single_epoch(model, fe, tr, crit, opt, sched, torch.device("cpu"), kw)
                    v = mu.validate_single_epoch(model, fe, te, crit, torch.device("cpu"), kw)
                epochs.append({"train": [float(x) for x in t[:5]], "train_cm": t[5].tolist(),
                               "val": [float(x) for x in v[:5]], "val_cm": v[5].tolist(),
                               "val_preds": [float(x) for x in v[7]], "val_labels": [float(x) for x in v[10]],
                               "val_gestures": [float(x) for x in v[11]], "val_subjects": [str(s[0]) if isinstance(s, (list, tuple)) else str(s) for s in v[12]]})
            rec["epochs"] = epochs
            # frame -> window post-processing of the last validation pass (a18)
            last = epochs[-1]
            pw, ew, gw, sw = mu.window_predictions(np.asarray(last["val_preds"]), np.asarray(last["val_labels"]),
                                                    np.asarray(last["val_gestures"]), np.asarray(last["val_subjects"]),
                                                    window_size=10, stride=6, binary=True)
            rec["window_preds"] = pw.reshape(-1).tolist()
            rec["window_labels"] = ew.reshape(-1).tolist()
            rec["window_subjects"] = sw["subject"].tolist()
            res[name] = rec
            print("frame case", name, epochs[-1]["train"][:2], epochs[-1]["val"][:2], len(rec["window_preds"]))
    json.dump(res, open(os.path.join(HERE, "epochs.json"), "w"))


# ---------------------------------------------------------------- ES / Sequential loops (a14, a15)
class _LongTargetCE(torch.nn.CrossEntropyLoss):
    """The committed ES / Sequential loops hand CrossEntropyLoss float class indices (raises on CPU)
    and, in the cascade, the target -1 for masked rows (raises on CPU) -- SURVEY.md section 8c.  The
    reference functions are executed UNMODIFIED; only the criterion they are given / construct is
    this subclass, which applies the two restatements SURVEY prescribes: cast to long, clamp(min=0)
    (the clamped rows are multiplied by a zero mask afterwards, modeling_utils.py:623)."""

    def forward(self, input, target):
        if target.dtype.is_floating_point or target.dtype != torch.long:
            target = target.long()
        return super().forward(input, target.clamp(min=0))


def gen_epochs_es():
    mu, du = ref.modeling_utils, ref.dataset_utils
    fold = synthetic.make_fold(**cases.FOLD_ARGS)
    res = {}
    dev = torch.device("cpu")
    with tempfile.TemporaryDirectory() as tmp:
        path = synthetic.write_fold(fold, os.path.join(tmp, "fold")) + "/"
        # --- error-specific 6-class
        kw = cases.ES_CASE
        with quiet:
            tr, te = du.retrieve_dataloaders_window(path, kw, window_size=10, stride=6)
            fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, dev, tr.dataset.binary_error_distribution, 10)
        no_dropout(model, fe)
        crit = _LongTargetCE()
        epochs = []
        for ep in range(kw["n_epochs"]):
            with quiet, contextlib.redirect_stderr(io.StringIO()):
                t = mu.train_single_epoch_ES(model, fe, tr, crit, opt, sched, dev, kw)
                v = mu.validate_single_epoch_ES(model, fe, te, crit, dev, kw)
            epochs.append({"train": [float(x) for x in t[:7]], "train_cm_binary": t[7].tolist(), "train_cm_macro": t[8].tolist(),
                           "val": [float(x) for x in v[:7]], "val_cm_binary": v[7].tolist(), "val_cm_macro": v[8].tolist(),
                           "val_probs": [float(x) for x in v[10]], "val_preds": [int(x) for x in v[11]],
                           "val_labels": [int(x) for x in v[12]]})
        res["es"] = {"epochs": epochs, "n_train": len(tr.dataset)}
        print("ES", epochs[-1]["train"][:3], epochs[-1]["val"][:3])
        # --- cascade: frozen binary model + masked 5-class model
        kwb, kws = cases.SEQ_BINARY_CASE, cases.SEQ_CASE
        with quiet:
            bfe, bmodel, bcrit, bopt, bsched = mu.define_model_objects(kwb, cases.IN_FEATURES, dev, tr.dataset.binary_error_distribution, 10)
        no_dropout(bmodel, bfe)
        with quiet, contextlib.redirect_stderr(io.StringIO()):
            for _ in range(kwb["n_epochs"]):
                mu.train_single_epoch(bmodel, bfe, tr, bcrit, bopt, bsched, dev, kwb)
            tr2, te2 = du.retrieve_dataloaders_window(path, kws, window_size=10, stride=6)
            fe, model, crit, opt, sched = mu.define_model_objects(kws, cases.IN_FEATURES, dev, tr.dataset.binary_error_distribution, 10)
        no_dropout(model, fe)
        saved = torch.nn.CrossEntropyLoss
        torch.nn.CrossEntropyLoss = _LongTargetCE
        try:
            epochs = []
            for ep in range(kws["n_epochs"]):
                with quiet, contextlib.redirect_stderr(io.StringIO()):
                    t = mu.train_single_epoch_Sequential(model, fe, tr2, None, opt, dev, sched, kws)
                    v = mu.validate_single_epoch_Sequential(model, fe, bmodel, bfe, te2, dev, kws)
                epochs.append({"train": [float(x) for x in t[:9]], "train_cm_all": t[9].tolist(), "train_cm_specific": t[10].tolist(),
                               "val": [float(x) for x in v[:9]], "val_cm_all": v[9].tolist(), "val_cm_specific": v[10].tolist(),
                               "val_preds_all": [int(x) for x in v[12]], "val_preds_specific": [int(x) for x in v[13]],
                               "val_labels_all": [int(x) for x in v[15]], "val_labels_specific": [int(x) for x in v[16]]})
        finally:
            torch.nn.CrossEntropyLoss = saved
        res["sequential"] = {"epochs": epochs}
        print("SEQ", epochs[-1]["train"][:3], epochs[-1]["val"][:3])
    json.dump(res, open(os.path.join(HERE, "epochs_es.json"), "w"))


# ---------------------------------------------------------------- window_predictions, multi-class rounding (a18)
def gen_window_predictions():
    mu = ref.modeling_utils
    out = {}
    for name, seed, nv, lo, hi, W, S in cases.WINDOW_CASES[:3]:
        g, e5, offsets = synthetic.label_tracks(seed, nv, lo, hi)
        # subject names deliberately NOT in sorted order: np.unique re-sorts them (Appendix A-6)
        names = np.concatenate([[synthetic.trial_name(nv - 1 - i)] * int(offsets[i + 1] - offsets[i]) for i in range(nv)])
        rng = np.random.Generator(np.random.PCG64(seed + 100))
        pb = (rng.random(len(g)) > 0.5).astype(np.float64)
        pm = rng.integers(0, 6, len(g)).astype(np.float64)
        lab = e5[:, 4].astype(np.float64)
        for tag, p, binary in (("bin", pb, True), ("multi", pm, False)):
            pw, ew, gw, sw = mu.window_predictions(p, lab, g.astype(np.float64), names, window_size=W, stride=S, binary=binary)
            out[f"{name}/{tag}/preds"] = pw.numpy().reshape(-1)
            out[f"{name}/{tag}/labels"] = ew.numpy().reshape(-1)
            out[f"{name}/{tag}/gest"] = gw.numpy().reshape(-1)
            out[f"{name}/{tag}/subj"] = np.asarray(sw["subject"].tolist())
    np.savez_compressed(os.path.join(HERE, "window_predictions.npz"), **out)
    print("window_predictions:", len(out))


# ---------------------------------------------------------------- TeCNo: non-causal variant, several videos (a10, ragged batches)
TECNO_EXTRA = dict(stages=2, layers=4, maps=64, dim=58, classes=2, weight_seed=123, input_seed=7, lengths=[40, 9, 65])


def gen_tecno_extra():
    """Reference MultiStageModel (models_TCN.py:17-137), causal and NOT causal, on three videos one forward each (the
    reference's DataLoader(batch_size=1) schedule): eval logits per video, and train-mode (dropout off) input / parameter
    gradient norms of sum(out * w) on the longest video.  Weights are not stored: the test rebuilds them from the seed and
    checks the state digest."""
    c = TECNO_EXTRA
    out, meta = {}, dict(c)
    for causal in (True, False):
        tag = "causal" if causal else "noncausal"
        torch.manual_seed(c["weight_seed"])
        with quiet:
            model = ref.models_TCN.MultiStageModel(c["stages"], c["layers"], c["maps"], c["dim"], c["classes"], causal)
        meta[f"{tag}/state"] = state_digest(model.state_dict())
        gen = torch.Generator().manual_seed(c["input_seed"])
        videos = [torch.randn(1, n, c["dim"], generator=gen) for n in c["lengths"]]
        model.eval()
        with torch.no_grad():
            for i, x in enumerate(videos):
                out[f"{tag}/logits{i}"] = model(x.permute(0, 2, 1)).numpy()
        no_dropout(model)
        model.train()
        x = videos[-1].clone().requires_grad_(True)
        y = model(x.permute(0, 2, 1))
        w = torch.randn(y.shape, generator=gen)
        (y * w).sum().backward()
        out[f"{tag}/dx"] = x.grad.numpy()
        out[f"{tag}/w"] = w.numpy()
        meta[f"{tag}/grad_names"] = [k for k, _ in model.named_parameters()]
        out[f"{tag}/grad_norms"] = np.asarray([float(p.grad.double().norm()) for _, p in model.named_parameters()])
        out[f"{tag}/grad_sums"] = np.asarray([float(p.grad.double().sum()) for _, p in model.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "tecno_extra.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "tecno_extra.json"), "w"), indent=1)
    print("tecno_extra:", len(out))


# ---------------------------------------------------------------- fixed trained weights -> validation outputs ("3 decimals")
def _freeze_fe_trunk(fe):
    """FE layers 0 and 1 keep their seed-42 initialisation (bit-reproducible from the seed, checked by digest), so the
    fixture carries only the tensors that training changed."""
    for name, p in fe.named_parameters():
        if not name.startswith("linear.output"):
            p.requires_grad_(False)


def _trained_tensors(fe, model):
    out = {f"fe/{k}": v.detach().numpy().copy() for k, v in fe.state_dict().items() if k.startswith("linear.output")}
    out.update({f"model/{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()})
    return out


def gen_fixed_weights():
    """The reference trains (dropout ON, a few epochs), then its validate_single_epoch[_ES|_Sequential] runs on those
    weights: per-sample predictions / probabilities, pooled scores and sklearn's roc_auc_score are the golden."""
    from sklearn.metrics import roc_auc_score
    mu, du = ref.modeling_utils, ref.dataset_utils
    fold = synthetic.make_fold(**cases.FIXED_FOLD_ARGS)
    dev = torch.device("cpu")
    arrays, meta = {}, {}
    with tempfile.TemporaryDirectory() as tmp:
        path = synthetic.write_fold(fold, os.path.join(tmp, "fold")) + "/"
        trained = {}
        for name, (kw, val_cfgs, n_ep) in cases.FIXED_CASES.items():
            kw = dict(kw, n_epochs=n_ep)
            Wt, St = cases.FIXED_TRAIN_WS[name]
            with quiet:
                tr, te = du.retrieve_dataloaders_window(path, kw, window_size=Wt, stride=St)
                counts = tr.dataset.binary_error_distribution if kw["error_type"] == "global" else tr.dataset.binary_error_distribution
                fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, dev, counts, Wt)
            rec = {"fe_init_sd": state_digest(fe.state_dict()), "model_init_sd": state_digest(model.state_dict()),
                   "train_W": Wt, "train_S": St, "epochs": n_ep}
            _freeze_fe_trunk(fe)
            saved_ce = torch.nn.CrossEntropyLoss
            if name == "lstm_es":
                crit = _LongTargetCE()
            if name == "lstm_seq":
                torch.nn.CrossEntropyLoss = _LongTargetCE
            try:
                for ep in range(n_ep):
                    with quiet, contextlib.redirect_stderr(io.StringIO()):
                        if name == "lstm_es":
                            mu.train_single_epoch_ES(model, fe, tr, crit, opt, sched, dev, kw)
                        elif name == "lstm_seq":
                            mu.train_single_epoch_Sequential(model, fe, tr, None, opt, dev, sched, kw)
                        else:
                            mu.train_single_epoch(model, fe, tr, crit, opt, sched, dev, kw)
                for k, v in _trained_tensors(fe, model).items():
                    arrays[f"{name}/{k}"] = v
                rec["fe_trained_sd"] = state_digest(fe.state_dict())
                rec["model_trained_sd"] = state_digest(model.state_dict())
                trained[name] = (fe, model, crit)
                rec["val"] = {}
                for (W, S) in val_cfgs:
                    tag = f"w{W}_s{S}"
                    with quiet, contextlib.redirect_stderr(io.StringIO()):
                        _, te = du.retrieve_dataloaders_window(path, kw, window_size=W, stride=S)
                        if name == "lstm_es":
                            v = mu.validate_single_epoch_ES(model, fe, te, crit, dev, kw)
                        elif name == "lstm_seq":
                            bfe, bmodel, _ = trained["lstm_global"]
                            v = mu.validate_single_epoch_Sequential(model, fe, bmodel, bfe, te, dev, kw)
                        else:
                            v = mu.validate_single_epoch(model, fe, te, crit, dev, kw)
                    if name == "lstm_es":
                        r = {"scores": [float(x) for x in v[:7]], "cm_binary": v[7].tolist(), "cm_macro": v[8].tolist(),
                             "probs": [float(x) for x in v[10]], "preds": [int(x) for x in v[11]], "labels": [int(x) for x in v[12]],
                             "labels_binary": [int(x) for x in v[13]], "preds_binary": [int(x) for x in v[14]]}
                    elif name == "lstm_seq":
                        r = {"scores": [float(x) for x in v[:9]], "cm_all": v[9].tolist(), "cm_specific": v[10].tolist(),
                             "preds_all": [int(x) for x in v[12]], "preds_specific": [int(x) for x in v[13]],
                             "labels_all": [int(x) for x in v[15]], "labels_specific": [int(x) for x in v[16]]}
                    else:
                        preds, probs, labels = [float(x) for x in v[7]], [float(x) for x in v[8]], [float(x) for x in v[10]]
                        r = {"scores": [float(x) for x in v[:5]], "cm": v[5].tolist(), "preds": preds, "probs": probs,
                             "labels": labels, "auc": float(roc_auc_score(labels, probs)), "subjects": list(v[12])}
                    r["n_test"] = len(te.dataset)
                    rec["val"][tag] = r
                    print("fixed", name, tag, r["n_test"], [round(x, 4) for x in r["scores"]], r.get("auc"))
            finally:
                torch.nn.CrossEntropyLoss = saved_ce
            meta[name] = rec
    np.savez_compressed(os.path.join(HERE, "fixed_weights.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "fixed_weights.json"), "w"))
    print("fixed_weights:", sum(v.size for v in arrays.values()), "floats")


# ---------------------------------------------------------------- post-processing, fold aggregation, table ingest (f1, f4)
def gen_postproc():
    """frame2window / compute_window_metrics / create_summary_df / create_binary_mask / load_data(video_data_path) /
    the ensemble notebook's soft vote and cascade, all by executing the reference (the two notebook cells are restated
    from ensemble.ipynb cell 6 lines 10-34 and cell 15 lines 53-63 with the same numpy / sklearn calls)."""
    from sklearn.metrics import accuracy_score, confusion_matrix, f1_score, jaccard_score
    mu, du = ref.modeling_utils, ref.dataset_utils
    res = {}
    outs, preds_b, preds_m, labels, gests, subjects = cases.postproc_inputs()
    for tag, preds, binary in (("binary", preds_b, True), ("multi", preds_m, False)):
        wp, wl, wg, ws = mu.frame2window(outs, preds, labels, gests, subjects, window_size=10, stride=6, binary=binary)
        df, cm = mu.compute_window_metrics(outs, preds, labels, gests, subjects, window_size=10, stride=6, binary=binary)
        res[f"window_metrics_{tag}"] = {"summary": {c: df.loc["Windowed Metrics", c] for c in df.columns}, "cm": cm.tolist(),
                                        "n_windows": {o: int(len(wp[o])) for o in wp},
                                        "preds_digest": {o: digest(wp[o].numpy()) for o in wp},
                                        "labels_digest": {o: digest(wl[o].numpy()) for o in wl},
                                        "first_subjects": {o: ws[o]["subject"].tolist()[:3] for o in ws}}
    rng = np.random.Generator(np.random.PCG64(9))
    lists = [rng.random(5) for _ in range(6)]
    samples_train, samples_test = rng.integers(3000, 3700, 5), rng.integers(600, 1100, 5)
    rates, times = rng.random(5) * 2, rng.random(5) * 3
    df = mu.create_summary_df(*lists, samples_train, samples_test, rates, times)
    res["summary_df"] = {"inputs": {"lists": [l.tolist() for l in lists], "samples_train": samples_train.tolist(),
                                    "samples_test": samples_test.tolist(), "rates": rates.tolist(), "times": times.tolist()},
                         "cells": {f"{r}/{c}": (None if isinstance(df.loc[r, c], float) and np.isnan(df.loc[r, c]) else str(df.loc[r, c]))
                                   for r in df.index for c in df.columns}}
    # ensemble.ipynb cell 6: soft vote of two window models' probabilities
    n = 4252
    pa, pb = rng.random(n), rng.random(n)
    pa[:4], pb[:4] = [0.5, 0.25, 0.75, 0.0], [0.5, 0.75, 0.25, 1.0]
    pa, pb = pa.astype(np.float32), pb.astype(np.float32)
    lab = (rng.random(n) > 0.45).astype(np.int64)
    ens = ((pa.astype(np.float64) + pb.astype(np.float64)) / 2 >= 0.5).astype(int)
    res["soft_vote"] = {"seed": 9, "n": n, "preds_digest": digest(ens.astype(np.int64)), "acc": float(accuracy_score(lab, ens)),
                        "f1": float(f1_score(lab, ens)), "jaccard": float(jaccard_score(lab, ens)),
                        "cm": confusion_matrix(lab, ens).tolist()}
    np.savez_compressed(os.path.join(HERE, "postproc_inputs.npz"), pa=pa, pb=pb, lab=lab)
    # ensemble.ipynb cell 15: cascade
    b, m = rng.integers(0, 2, n), rng.integers(0, 6, n)
    ens = np.zeros_like(m)
    ens[b == 1] = m[b == 1]
    res["cascade"] = {"digest": digest(ens.astype(np.int64)), "binary_digest": digest(b.astype(np.int64))}
    np.savez_compressed(os.path.join(HERE, "postproc_cascade.npz"), b=b.astype(np.int64), m=m.astype(np.int64), ens=ens.astype(np.int64))
    # create_binary_mask with mask_position_ND files, load_data with the video_data_path schema
    fold = synthetic.make_fold(seed=5, n_train=3, n_test=2, t_lo=60, t_hi=90)
    with tempfile.TemporaryDirectory() as tmp:
        path, vpath = synthetic.write_fold_video_schema(fold, os.path.join(tmp, "fold"), os.path.join(tmp, "video"))
        flat = du.load_data(path + "/", "train.csv", video_data_path=vpath + "/")
        res["load_data_video"] = {"image": digest(flat[0]), "kin": digest(flat[1]), "g": digest(flat[2]), "e": digest(flat[3]),
                                  "subjects": flat[4]["subject"].tolist()[::40], "n": int(flat[0].shape[0])}
        flat0 = du.load_data(path + "/", "train.csv")
        res["load_data_fold"] = {"image": digest(flat0[0]), "kin": digest(flat0[1]), "n": int(flat0[0].shape[0])}
        subj = np.concatenate([[t.name] * len(t.g) for t in fold.test])
        pb_ = {"t": (rng.random(len(subj)) > 0.5).astype(int).tolist()}
        masks = {}
        for t in fold.test[:1]:
            mk = torch.from_numpy(rng.random(len(t.g)) < 0.2)
            torch.save(mk, os.path.join(path, f"mask_position_ND_{t.name}.pth"))
            masks[t.name] = mk.numpy().tolist()
        with quiet:
            bm, bs = mu.create_binary_mask(pb_, {"t": subj.tolist()}, "t", path, {"delete_ND": True})
            bm0, bs0 = mu.create_binary_mask(pb_, {"t": subj.tolist()}, "t", path, {"delete_ND": False})
        res["binary_mask"] = {"preds": pb_["t"], "subjects": subj.tolist(), "masks": masks, "mask_out": bm.tolist(),
                              "subjects_out": bs.tolist(), "mask_out_keep_nd": bm0.tolist(), "fold_seed": 5}
    json.dump(res, open(os.path.join(HERE, "postproc.json"), "w"))
    print("postproc:", list(res))


if __name__ == "__main__":
    which = sys.argv[1:] or ["window_index", "powerset", "models", "epochs", "epochs_es", "window_predictions", "tecno_extra",
                             "fixed_weights", "postproc"]
    for w in which:
        globals()[f"gen_{w}"]()
