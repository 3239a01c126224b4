import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref/libqlref.so (test infrastructure)."""
    from oracle import refbridge
    if not refbridge.available():
        pytest.skip("oracle/_ref/libqlref.so not built (needs /root/reference at build time)")
    refbridge.lib()
    return refbridge


@pytest.fixture(scope="session")
def ctx():
    import tensortoolkit_b200 as tk
    return tk.Context(0)
