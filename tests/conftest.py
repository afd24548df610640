import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure); built on demand."""
    from oracle import pyoracle
    pyoracle.load()
    return pyoracle


@pytest.fixture(scope="session")
def trt():
    import tiny_raytracer_b200
    tiny_raytracer_b200._lib.load()
    return tiny_raytracer_b200


@pytest.fixture(scope="session")
def ctx(trt):
    """A GPU context. No skip-on-failure: a gpu-marked test without a usable B200 must fail loudly."""
    c = trt.Context(0)
    yield c
    c.close()
