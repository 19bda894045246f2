import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


# kernel-string name -> (family, amp, how the metric is given) for the oracle
def golden_kernel_specs(g):
    import re
    from numpy import array  # noqa: F401  (used by eval of the stored strings)

    specs = {}
    for name in [str(c) for c in g["cases"]]:
        s = str(g["kstr_" + name])
        specs[name] = s
    return specs


@pytest.fixture(scope="session")
def gpu_ready():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from treegp_b200 import _cabi

    _cabi.load()  # must not silently fall back: a missing library is an error on a GPU box
    return torch.device("cuda:0")
