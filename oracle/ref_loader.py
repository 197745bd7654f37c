"""ORACLE — TEST INFRASTRUCTURE ONLY. Loads the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
oracle/make_golden.py to mint tests/golden/*.npz and by optional cross-check tests that skip when the
reference is absent. No reference file is copied or modified; missing third-party imports are
satisfied with empty stub modules.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REF = Path("/root/reference")


def available() -> bool:
    return (REF / "ultralytics" / "nn" / "modules" / "transformer.py").exists()


def load_decoder_modules():
    """transformer.py + utils.py without running ultralytics/__init__.py (needs only torch/numpy)."""
    if "ultralytics.nn.modules.transformer" in sys.modules and getattr(
            sys.modules["ultralytics"], "_moyolo_light", False):
        return sys.modules["ultralytics.nn.modules.transformer"], sys.modules["ultralytics.nn.modules.utils"]
    for name, path in (("ultralytics", REF / "ultralytics"), ("ultralytics.nn", REF / "ultralytics/nn"),
                       ("ultralytics.nn.modules", REF / "ultralytics/nn/modules")):
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        m._moyolo_light = True
        sys.modules[name] = m
    T = importlib.import_module("ultralytics.nn.modules.transformer")
    U = importlib.import_module("ultralytics.nn.modules.utils")
    return T, U


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        stub = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, stub)
        return stub

    def __call__(self, *a, **k):
        return None


def load_full_reference(msda_module=None):
    """`import ultralytics` + MOTR for real (head.py, qim.py), stubbing absent third-party packages.
    `msda_module` is registered as `MultiScaleDeformableAttention` (moyolo_b200.msda_ext or a stub)."""
    if getattr(sys.modules.get("ultralytics"), "_moyolo_light", False):
        for k in [k for k in sys.modules if k == "ultralytics" or k.startswith("ultralytics.")]:
            del sys.modules[k]
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.figure", "seaborn", "thop",
                 "pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval", "motmetrics",
                 "MOTR.datasets", "MOTR.datasets.samplers", "MOTR.datasets.coco_eval", "MOTR.datasets.panoptic_eval",
                 "MOTR.datasets.data_prefetcher"):
        if name not in sys.modules:
            s = _Stub(name)
            s.__path__ = []
            sys.modules[name] = s
    sys.modules["MultiScaleDeformableAttention"] = msda_module or _Stub("MultiScaleDeformableAttention")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    sys.argv = ["x"]  # head.py:111 calls parse_args() on the process argv
    import ultralytics  # noqa: F401
    head = importlib.import_module("ultralytics.nn.modules.head")
    qim = importlib.import_module("MOTR.models.qim")
    structures = importlib.import_module("MOTR.models.structures")
    return head, qim, structures
