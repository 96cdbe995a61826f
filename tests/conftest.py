import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """TEST INFRASTRUCTURE: the compiled, unmodified reference (oracle/_ref/libbox2d_ref.so)."""
    import harness as H
    if not H.have_reference():
        pytest.skip("oracle/_ref/libbox2d_ref.so not built and /root/reference absent")
    return H.load("reference")


@pytest.fixture(scope="session")
def emu():
    """TEST INFRASTRUCTURE: host emulation of the device step templates (tests/emu)."""
    import harness as H
    return H.load("emu")


@pytest.fixture(scope="session")
def product():
    """The product library; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as G
    G.build_product()
    import forge2d_b200
    return forge2d_b200.load_library()


@pytest.fixture(scope="session")
def gpu(product):
    if not product.f2dHasDevice():
        pytest.fail("forge2d_b200: no CUDA device visible - gpu tests need the B200 box (there is no CPU fallback)")
    return product
