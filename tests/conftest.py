import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that stalls (a kernel waiting on a seam flag that never comes) must fail, not block the session:
    pytest-timeout's thread method ends the process, which tears the context and the kernel down."""
    if not config.pluginmanager.hasplugin("timeout") or config.getoption("timeout", None):
        return  # no plugin, or a --timeout given on the command line takes precedence
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def ctx():
    """One libchrom_cuda context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    import chromatography_b200 as cb
    c = cb.Context([0])
    yield c
    c.close()
