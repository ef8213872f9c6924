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


def _cuda_available() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` need a CUDA device and libblp.so: without a device they are skipped, not
    failed (the product itself still fails loudly there: test_instances_and_cabi)."""
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this environment')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
