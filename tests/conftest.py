import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_product():
    """Import 2d-ekf-slam_b200/ekf_b200.py (the directory name is not a Python identifier)."""
    name = "ekf_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "2d-ekf-slam_b200", "ekf_b200.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def built():
    """Everything compiled (products + checkers). Cheap when already built."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def ekf(built):
    return load_product()


@pytest.fixture(scope="session")
def oracle(built):
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref(built):
    from oracle_lib import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference or a prebuilt binary)")
    return Ref()
