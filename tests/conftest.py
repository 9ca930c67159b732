import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "scikit-gstat_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (plain `pytest tests` on a CPU box);
    on the GPU box they run and fail loudly if the library cannot be loaded."""
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy view on one golden npz (reference outputs, see oracle/make_golden.py)."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.manifest = json.loads(str(self._z["manifest"]))

    def inputs(self, dataset):
        return self._z["in/%s/coords" % dataset], self._z["in/%s/values" % dataset]

    def out(self, case, field):
        return self._z["%s/%s" % (case["key"], field)]


@pytest.fixture(scope="session")
def golden_variogram():
    return Golden("variogram")


@pytest.fixture(scope="session")
def golden_directional():
    return Golden("directional")


@pytest.fixture(scope="session")
def golden_kriging():
    return Golden("kriging")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden_cv():
    return Golden("cross_validation")


@pytest.fixture(scope="session")
def golden_mv():
    return Golden("multivariate")


@pytest.fixture(scope="session")
def golden_kriging_ext():
    return Golden("kriging_ext")


@pytest.fixture(scope="session")
def golden_binners():
    return Golden("binners")


@pytest.fixture(scope="session")
def golden_estimators():
    return Golden("estimators")


@pytest.fixture(scope="session")
def golden_sparse():
    return Golden("sparse")
