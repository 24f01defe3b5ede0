"""How chaotic is the REFERENCE's own epoch?  Perturb the oracle's (= reference's) inputs by 1e-6 / 1e-7 relative and
watch the epoch scores move.  Used to set the epoch-level parity bar (tests/test_gpu_models.py::_scores_close).

    python tests/golden/sensitivity_probe.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
from multimodal_error_detection_b200 import synthetic  # noqa: E402
from oracle import loops, nets  # noqa: E402
from test_oracle_golden import _oracle_window_loaders  # noqa: E402

kw, W, S = cases.EPOCH_CASES["lstm_global"]
fold = synthetic.make_fold(**cases.FOLD_ARGS)
for pert in (0.0, 1e-6, 1e-7):
    tr, te = _oracle_window_loaders(fold, kw, W, S)
    if pert:
        g = torch.Generator().manual_seed(0)
        tr.dataset.image = tr.dataset.image * (1 + pert * torch.randn(tr.dataset.image.shape, generator=g))
    fe, model, crit, opt, sched = nets.build_objects(kw, cases.IN_FEATURES, tr.dataset.binary_error_distribution, W)
    nets.disable_dropout(model, fe)
    print(pert, [np.round(loops.train_epoch(model, fe, tr, crit, opt, sched, kw)[:5], 5).tolist() for _ in range(2)])
