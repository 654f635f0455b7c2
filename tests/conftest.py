import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu_mod():
    from tests.emu import emu
    emu.build()
    return emu


@pytest.fixture(scope="session")
def gpu_ctx():
    """One engine context on cuda:0.  Fails loudly (no skip, no fallback) when the CUDA library or
    the device is missing: `-m gpu` tests must run native code."""
    from mongeampere_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
