import importlib.util
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def load_package():
    """The package directory is named aom-av1-psy_b200 (not an identifier): register it as aom_av1_psy_b200."""
    name = "aom_av1_psy_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "aom-av1-psy_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def tfgpu(pkg):
    ctx = pkg.TemporalFilterGpu()
    yield ctx
    ctx.close()
