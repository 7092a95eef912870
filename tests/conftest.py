"""pytest configuration: markers, repo paths, build-on-demand of the test-side artefacts."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything the suite needs is compiled once per session (no-op when up to date)."""
    import __graft_entry__ as ge

    ge.build()
