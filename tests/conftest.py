import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_pf():
    return np.load(os.path.join(GOLDEN, "probability_fusion.npz"))


@pytest.fixture(scope="session")
def golden_scorer():
    arrays = np.load(os.path.join(GOLDEN, "scorer_cases.npz"))
    with open(os.path.join(GOLDEN, "scorer_cases.json")) as f:
        metas = json.load(f)
    return arrays, metas


@pytest.fixture(scope="session")
def golden_config1():
    arrays = np.load(os.path.join(GOLDEN, "config1.npz"))
    with open(os.path.join(GOLDEN, "config1.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def golden_mf():
    arrays = np.load(os.path.join(GOLDEN, "multifield_blockmax.npz"))
    with open(os.path.join(GOLDEN, "multifield_blockmax.json")) as f:
        meta = json.load(f)
    return arrays, meta
