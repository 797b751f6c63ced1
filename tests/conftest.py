import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def emulated(monkeypatch):
    """CPU emulation of the C-ABI launch helpers (tests/emu.py) — host-logic tests only."""
    import emu
    emu.install(monkeypatch)
    emu.install_train(monkeypatch)
    return emu
