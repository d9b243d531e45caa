"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, ABI surface.
``-m gpu`` runs on a B200: parity of the CUDA path (through the C ABI) against the oracle, the
golden fixtures and size-independent properties.  Nothing in the test-suite reads /root/reference.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load the in-tree CUDA library; works without a GPU."""
    from ship_track_estimators_b200 import _native, build

    build.build()
    return _native.load()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("a gpu-marked test ran without a CUDA device")
    return torch.device("cuda:0")


def pytest_terminal_summary(terminalreporter):
    """Worst parity error per fixture / test label (filtered mean, filtered cov, smoothed mean, smoothed cov)."""
    from _helpers import WORST

    if WORST:
        terminalreporter.write_line("worst parity errors (filtered mean, filtered cov, smoothed mean, smoothed cov):")
        for key in sorted(WORST):
            terminalreporter.write_line(f"  {key:40s} " + " ".join(f"{v:9.2e}" for v in WORST[key]))
