import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; built on demand (nvcc cross-compiles without a GPU)."""
    import agentgo_b200
    agentgo_b200.build_lib()
    return agentgo_b200.lib()


@pytest.fixture(scope="session")
def port_lib():
    from oracle import pyoracle
    pyoracle.build_port()
    return pyoracle
