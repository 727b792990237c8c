import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """npz with 'case/field' keys -> {case: {field: array}}"""

    def __init__(self, name: str):
        self._z = np.load(GOLDEN / name, allow_pickle=False)
        self.cases = {}
        for k in self._z.files:
            case, field = k.split("/", 1)
            self.cases.setdefault(case, {})[field] = self._z[k]


@pytest.fixture(scope="session")
def golden_chamfer():
    return Golden("chamfer_ref.npz").cases


@pytest.fixture(scope="session")
def golden_entropy():
    return Golden("entropy_ref.npz").cases


@pytest.fixture(scope="session")
def golden_fps():
    return Golden("fps_ref.npz").cases


@pytest.fixture(scope="session")
def golden_pointnet():
    return Golden("pointnet_ref.npz").cases


def max_rel(a, b):
    """max-norm relative error |a-b|_inf / |b|_inf"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
