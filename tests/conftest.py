import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def b200():
    """The product package (ctypes plumbing over the C ABI)."""
    return importlib.import_module("computer-graphics_b200")


@pytest.fixture(scope="session")
def renderer(b200):
    r = b200.Renderer(0)
    yield r
    r.close()


@pytest.fixture(scope="session")
def cornell_rt():
    import helpers as h
    g = load_golden("rt_cornell.npz")
    return g["tris"].view(h.RT_TRI).copy(), g["spheres"].view(h.RT_SPHERE).copy()
