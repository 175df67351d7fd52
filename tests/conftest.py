"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors / reference, host logic,
ABI export checks. `-m gpu` runs on a B200: the CUDA path against the oracle.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """The product library is built by __graft_entry__.build(); the oracle side only
    needs gcc, so make sure it exists (seconds)."""
    from oracle import harness as H
    if not (os.path.exists(H.PORT) and os.path.exists(H.DRIVER)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "restatement", "driver"])
    prod = os.path.join(ROOT, "turtle_b200", "libturtle_b200.so")
    if not os.path.exists(prod):
        import __graft_entry__
        __graft_entry__.build()


_ensure_built()


@pytest.fixture(scope="session")
def scratch(tmp_path_factory):
    return str(tmp_path_factory.mktemp("turtle"))


@pytest.fixture(scope="session")
def small_stack(scratch):
    """2 x 2 grid of 1201 x 1201 `.hgt` tiles at N45-46 / E2-3, tile (46, 3) missing."""
    from turtle_b200 import synth
    d = os.path.join(scratch, "stack1201")
    synth.write_hgt_stack(d, 45, 2, 2, 2, n=1201, skip=((46, 3),))
    return d


@pytest.fixture(scope="session")
def has_gpu():
    import turtle_b200 as tb
    return tb.device_count() > 0
