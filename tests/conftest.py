import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_golden(name):
    z = np.load(GOLD / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def rel_rms(a, b):
    """max|a-b| / rms(b): the per-tensor tolerance measure of SURVEY.md §8(c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rms = float(np.sqrt(np.mean(b * b))) or 1.0
    return float(np.max(np.abs(a - b))) / rms


@pytest.fixture(scope="session")
def golden():
    return load_golden
