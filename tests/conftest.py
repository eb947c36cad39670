import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One amplisolve_b200 context on cuda:0.  GPU tests fail (not skip) without the native library."""
    from amplisolve_b200 import Context
    c = Context(0)
    yield c
    c.close()
