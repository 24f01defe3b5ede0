"""Import the UNMODIFIED reference (``/root/reference/MED``) as an executable oracle.

Only usable in the build container (the GPU box has no /root/reference); used by
``tests/golden/make_golden.py`` to produce the committed fixtures and by the optional
``-m "not gpu"`` cross-checks that skip when the reference is absent.  Recipe: SURVEY.md
Appendix B -- the two third-party imports the reference needs but this image lacks
(``mlflow``, ``clip``) are replaced by empty stub modules; nothing on the hot path calls them.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200MED_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MED"))


def import_reference():
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("mlflow", "mlflow.pytorch", "mlflow.artifacts", "clip"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    sys.modules["mlflow"].pytorch = sys.modules["mlflow.pytorch"]
    sys.modules["mlflow"].artifacts = sys.modules["mlflow.artifacts"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from MED.dataset import dataset_utils, CustomWindowDataset, CustomFrameDataset
    from MED.modeling import modeling_utils, models, models_TCN
    return types.SimpleNamespace(dataset_utils=dataset_utils, CustomWindowDataset=CustomWindowDataset,
                                 CustomFrameDataset=CustomFrameDataset, modeling_utils=modeling_utils,
                                 models=models, models_TCN=models_TCN)
