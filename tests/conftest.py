import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """The plain-C oracle port is compiled on demand (gcc, < 1 s)."""
    import oracle_lib as ol
    if not os.path.exists(ol.PORT_SO):
        ol.build_oracles(reference=False)
