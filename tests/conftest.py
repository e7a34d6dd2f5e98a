import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    # the library is built in-tree (cross-compiled here, shipped to the GPU box); never JIT at test time on the box
    # (loaded by path: importing the package itself requires the built library)
    import importlib.util
    spec = importlib.util.spec_from_file_location('_amtfeat_build', os.path.join(ROOT, 'amt_tools_b200', 'build.py'))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    if build.needs_build() and os.path.exists('/usr/local/cuda/bin/nvcc'):
        build.build()
    yield
