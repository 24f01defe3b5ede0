"""Shared parity-case definitions: the same seeded inputs feed the reference (golden
generation), the oracle (CPU tests) and the CUDA path (GPU tests)."""
import numpy as np
import torch


def _synthetic():
    from multimodal_error_detection_b200 import synthetic
    return synthetic

# (name, seed, n_videos, t_lo, t_hi, W, S)
WINDOW_CASES = [
    ("p_10_6", 1, 7, 300, 900, 10, 6),       # parity config (5 Hz)
    ("t_16_4", 2, 6, 300, 900, 16, 4),       # throughput config
    ("hz15_30_20", 3, 5, 400, 700, 30, 20),  # 15 Hz config
    ("dense_5_1", 4, 4, 60, 120, 5, 1),
    ("sparse_10_20", 5, 4, 200, 300, 10, 20),
    ("short_videos", 6, 9, 12, 40, 10, 6),   # videos barely longer than one window
]


def base_kwargs(**over):
    kw = dict(dataset_type="window", error_type="global", pos_weight=False, n_epochs=2, batch_size=64,
              lr=3e-4, lr_scheduler=True, weight_decay=1e-4, num_layers=3, hidden_size=128, video_dims=32,
              data_type="multimodal", delete_ND=True, return_train_preds=False, siamese=False,
              model_name="SimpleCNN")
    kw.update(over)
    return kw


IN_FEATURES = {"multimodal": 58, "video": 32, "kinematics": 26}

# name -> (exp_kwargs, window_size, class_counts, batch, input dims)
MODEL_CASES = {
    "cnn_w10": (base_kwargs(model_name="SimpleCNN"), 10, (0.4, 0.6)),
    "cnn_w30": (base_kwargs(model_name="SimpleCNN"), 30, (0.4, 0.6)),
    "lstm_w10": (base_kwargs(model_name="SimpleLSTM"), 10, (0.4, 0.6)),
    "lstm_w16_pw": (base_kwargs(model_name="SimpleLSTM", pos_weight=True), 16, (0.4, 0.6)),
    "lstm_w10_c6": (base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=6), 10, (0.4, 0.6)),
    "lstm_w10_c6_weighted": (base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=6,
                                         pos_weight=True), 10, (1.0, 2.0, 3.0, 4.0, 5.0, 6.0)),
    "lstm_w10_c5": (base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=5), 10, (0.4, 0.6)),
    "tecno_f58": (base_kwargs(model_name="TeCNo", dataset_type="frame", mstcn_stages=2, mstcn_layers=8,
                              mstcn_f_maps=64, mstcn_f_dim=58, out_features=2, mstcn_causal_conv=True), 0, (0.4, 0.6)),
    "tecno_f2048": (base_kwargs(model_name="TeCNo", dataset_type="frame", data_type="video", video_dims=2048,
                                mstcn_stages=2, mstcn_layers=8, mstcn_f_maps=64, mstcn_f_dim=2048, out_features=2,
                                mstcn_causal_conv=True), 0, (0.4, 0.6)),
}

MODEL_BATCH = 12
FRAME_T = 150


def model_inputs(name):
    """Seeded synthetic batch for a MODEL_CASES entry: (images, kin, labels)."""
    kw, W, _ = MODEL_CASES[name]
    gen = torch.Generator().manual_seed(1234)
    if kw["dataset_type"] == "window":
        images = torch.randn(MODEL_BATCH, W, 2048, generator=gen).clamp_min(0)
        kin = torch.randn(MODEL_BATCH, W, 26, generator=gen)
        if kw["error_type"] == "global":
            y = (torch.rand(MODEL_BATCH, generator=gen) > 0.5).float()
        else:
            y = torch.randint(0, kw["out_features"], (MODEL_BATCH,), generator=gen)
    else:
        images = torch.randn(1, FRAME_T, 2048, generator=gen).clamp_min(0)
        kin = torch.randn(1, FRAME_T, 26, generator=gen)
        y = (torch.rand(1, FRAME_T, generator=gen) > 0.5).float()
    return images, kin, y


# End-to-end epochs on an on-disk synthetic fold: (name, exp_kwargs, W, S, fold args)
FOLD_ARGS = dict(seed=42, n_train=8, n_test=3, t_lo=250, t_hi=450)
EPOCH_CASES = {
    "cnn_global_pw": (base_kwargs(model_name="SimpleCNN", pos_weight=True), 10, 6),
    "lstm_global": (base_kwargs(model_name="SimpleLSTM"), 10, 6),
    "lstm_w16": (base_kwargs(model_name="SimpleLSTM", return_train_preds=True), 16, 4),
}
ES_CASE = base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=6, n_epochs=2)
SEQ_CASE = base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=5, n_epochs=2)
SEQ_BINARY_CASE = base_kwargs(model_name="SimpleLSTM", error_type="global", n_epochs=4)
FRAME_EPOCH_CASES = {
    "tecno_multimodal": base_kwargs(model_name="TeCNo", dataset_type="frame", mstcn_stages=2, mstcn_layers=8,
                                    mstcn_f_maps=64, mstcn_f_dim=58, out_features=2, mstcn_causal_conv=True,
                                    batch_size=1),
}


# Fixed-weights validation ("3 decimals" bar): the reference trains a model on FIXED_FOLD_ARGS (dropout on, FE layers 0/1
# frozen at their seed-42 initialisation so that the fixture only has to carry the trained tensors), the trained weights are
# committed (tests/golden/fixed_weights.npz) and the reference's validate_* outputs on them are the golden
# (tests/golden/fixed_weights.json).  name -> (exp_kwargs, [(W, S), ...] validation configs, epochs trained)
FIXED_FOLD_ARGS = dict(seed=77, n_train=6, n_test=6, t_lo=300, t_hi=600)
FIXED_CASES = {
    "lstm_global": (base_kwargs(model_name="SimpleLSTM", pos_weight=True, lr=1e-3, batch_size=128), [(16, 4), (10, 6)], 8),
    "cnn_global": (base_kwargs(model_name="SimpleCNN", lr=1e-3, batch_size=128), [(10, 6)], 8),
    "lstm_es": (base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=6, num_layers=1, lr=1e-3,
                            batch_size=128), [(10, 6)], 8),
    "lstm_seq": (base_kwargs(model_name="SimpleLSTM", error_type="all_errors", out_features=5, num_layers=1, lr=1e-3,
                             batch_size=128), [(10, 6)], 8),
}
FIXED_TRAIN_WS = {"lstm_global": (16, 4), "cnn_global": (10, 6), "lstm_es": (10, 6), "lstm_seq": (10, 6)}


def postproc_inputs():
    """Seeded frame-level prediction / label / gesture / subject lists for three 'folds' (shared with the tests)."""
    outs, preds_b, preds_m, labels, gests, subjects = ["1out", "2out", "3out"], {}, {}, {}, {}, {}
    for k, out in enumerate(outs):
        nv = 4 + k
        g, e5, offsets = _synthetic().label_tracks(200 + k, nv, 120, 260)
        names = np.concatenate([[_synthetic().trial_name(nv - 1 - i)] * int(offsets[i + 1] - offsets[i]) for i in range(nv)])
        rng = np.random.Generator(np.random.PCG64(300 + k))
        lab = e5[:, 4].astype(np.float64)
        flip = rng.random(len(g)) < 0.25
        preds_b[out] = np.where(flip, 1.0 - lab, lab).tolist()
        preds_m[out] = rng.integers(0, 6, len(g)).astype(np.float64).tolist()
        labels[out] = lab.tolist()
        gests[out] = g.astype(np.float64).tolist()
        subjects[out] = names.tolist()
    return outs, preds_b, preds_m, labels, gests, subjects


def all_label_rows():
    """Every binary combination of the five reference error columns, then seeded random rows."""
    rows = [[(i >> b) & 1 for b in range(5)] for i in range(32)]
    rng = np.random.Generator(np.random.PCG64(7))
    rows += rng.integers(0, 2, size=(200, 5)).tolist()
    return np.asarray(rows, dtype=np.float32)
