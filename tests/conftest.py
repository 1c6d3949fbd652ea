import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "markerless-robot-camera-calibration_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not os.path.isdir(REFERENCE):
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "reference_geometry.npz"))


@pytest.fixture(scope="session")
def cad_points():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "cad_hand_points.npz"))["xyz"]


@pytest.fixture(scope="session")
def built_lib():
    """make sure libb2me.so exists (cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2me_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
