import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One libchrom_cuda context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    import chromatography_b200 as cb
    c = cb.Context([0])
    yield c
    c.close()
