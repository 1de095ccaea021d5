import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def bplx_env(monkeypatch):
    """Set / unset the library's testing switches (BPLX_*): the library caches them (never a getenv on a launch path), so
    every change is followed by bplx_reload_env() -- also when the test ends."""
    from bpl_next_b200 import _abi

    def set_(**kw):
        for k, v in kw.items():
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, str(v))
        _abi.lib().bplx_reload_env()

    yield set_
    monkeypatch.undo()
    _abi.lib().bplx_reload_env()
