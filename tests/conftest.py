"""pytest configuration: the ``gpu`` marker and loaders for the oracle (test infrastructure)."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_ref_cppsim():
    """The compiled, unmodified reference pybind module (oracle/_ref, built by oracle/Makefile) or None."""
    import glob

    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_cppsim*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("_cppsim", hits[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError:
        return None
    return mod


@pytest.fixture(scope="session")
def ref_cppsim():
    mod = load_ref_cppsim()
    if mod is None:
        pytest.skip("oracle/_ref/_cppsim not built (run `make -C oracle`)")
    return mod
