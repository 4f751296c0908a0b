import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The C-ABI library must exist for both suites (CPU: load + symbol check; GPU: everything)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('_gsatb_build', os.path.join(ROOT, 'dp_gsat_b200', 'build.py'))
    b = importlib.util.module_from_spec(spec)      # by path: importing the package itself requires the built .so
    spec.loader.exec_module(b)
    if b.needs_build():
        b.build()
    yield


GOLDEN = os.path.join(ROOT, 'tests', 'golden')
