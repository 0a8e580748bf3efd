"""Shared fixtures.  `-m "not gpu"` covers the oracle, the host logic and the C-ABI export check;
`-m gpu` holds the parity tests proper (CUDA path vs oracle / golden vectors, through the C ABI)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")  # ordered fp64 sums in the reference's OpenMP module

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle import RefLib, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference): `make -C oracle ref`")
    return RefLib()


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pixeltovoxelprojector_b200 as pvp
    return pvp.default_context()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
