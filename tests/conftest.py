import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def blp_lib():
    """Build (if stale) and load libblp.so; the GPU tests call the product only through it."""
    from simple_mip_solver_b200 import _build, engine
    _build.build_extension()
    return engine.load_library()
