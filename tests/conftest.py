import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gridapmhd_jl_b200  # noqa: E402,F401  (registers the in-tree package directory `gridapmhd.jl_b200/`)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def mhdlib():
    """Initialised libmhdb200 on cuda:0 (GPU tests only). Fails loudly when the extension is missing."""
    from gridapmhd_jl_b200 import lib as L

    L.init(0)
    yield L
    L.finalize()
