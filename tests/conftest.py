import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference + oracle/_ref (only in the build container)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure libqsb.so and liboracle.so exist (built in-tree; a no-op when already built)."""
    import __graft_entry__ as g
    g.build(quiet=True)
