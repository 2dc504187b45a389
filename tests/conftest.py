"""pytest configuration: registers the `gpu` marker and puts the package on sys.path.

`-m "not gpu"` runs here (no GPU): oracle vs the reference's restated known-answer cases,
host logic, C-ABI symbol checks, gloo world_size-2 tests.  `-m gpu` runs on a B200 and
compares the CUDA path (through the C ABI) with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hoomd-tf_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle
