import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def port():
    """The C restatement of the reference algorithm (oracle/restate.c), built on demand."""
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref/libssa_ref.so); skipped where it was not built."""
    from oracle import refapi
    r = refapi.reference()
    if r is None:
        pytest.skip("oracle/_ref/libssa_ref.so not built (needs /root/reference: make -C oracle ref)")
    return r


@pytest.fixture(scope="session")
def prod_host():
    """The product library for host-only calls (no GPU needed); fails if it is not built."""
    import libssa_b200
    return libssa_b200.load()


@pytest.fixture(scope="session")
def prod():
    """The product library for searches: requires a GPU, never falls back."""
    import torch
    assert torch.cuda.is_available(), "GPU test started without a CUDA device"
    import libssa_b200
    lib = libssa_b200.load()
    lib.set_output_mode(0)
    return lib
