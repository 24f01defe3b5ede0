"""Fixed-weights cases (tests/cases.py::FIXED_CASES): the fold, and loading the committed trained tensors into freshly built
modules -- the FeatureExtractor trunk (layers 0 / 1) stays at its seed-42 initialisation, checked by digest."""
import json
import os

import numpy as np
import torch

import cases
from hashing import state_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def meta():
    return json.load(open(os.path.join(GOLDEN, "fixed_weights.json")))


def arrays():
    return np.load(os.path.join(GOLDEN, "fixed_weights.npz"))


def write_fold(tmp_path) -> str:
    from multimodal_error_detection_b200 import synthetic
    fold = synthetic.make_fold(**cases.FIXED_FOLD_ARGS)
    return synthetic.write_fold(fold, os.path.join(str(tmp_path), "fixed_fold")) + "/"


def cpu_sd(module):
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}


def load_trained(name, fe, model, check_init=True):
    """Load the reference-trained tensors of case `name`; asserts that the freshly built modules start from the reference's
    seed-42 weights and end up with exactly the reference's trained state (digests recorded by make_golden.py)."""
    m, arr = meta()[name], arrays()
    if check_init:
        assert state_digest(cpu_sd(fe)) == m["fe_init_sd"] and state_digest(cpu_sd(model)) == m["model_init_sd"]
    fe_sd = {k[len(name) + 4:]: torch.from_numpy(arr[k]) for k in arr.files if k.startswith(f"{name}/fe/")}
    model_sd = {k[len(name) + 7:]: torch.from_numpy(arr[k]) for k in arr.files if k.startswith(f"{name}/model/")}
    missing = fe.load_state_dict(fe_sd, strict=False)
    assert not missing.unexpected_keys and all(not k.startswith("linear.output") for k in missing.missing_keys)
    model.load_state_dict(model_sd)
    assert state_digest(cpu_sd(fe)) == m["fe_trained_sd"] and state_digest(cpu_sd(model)) == m["model_trained_sd"]
    return m
