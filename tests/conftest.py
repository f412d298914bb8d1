import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_c():
    import oracle

    return oracle.backend("C")


@pytest.fixture(scope="session")
def oracle_ob():
    import oracle

    return oracle.backend("OB")


@pytest.fixture(scope="session")
def bm():
    """The product package, on a GPU box only.  Fails loudly (never skips to a fallback) if CUDA is missing."""
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import bandedmatrices_b200 as bm

    bm.load()
    return bm


@pytest.fixture
def rng():
    return np.random.default_rng(20261017)
