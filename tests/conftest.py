import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libdic_ref.so (the compiled reference)")


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Compile the checkers (always) and the CUDA library (if nvcc is around and it is missing)."""
    import oracle
    oracle.build()
    lib = os.path.join(ROOT, "correlation_b200", "libdic_b200.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        import __graft_entry__
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def have_gpu():
    try:
        from correlation_b200 import engine
        return engine.load_library().dic_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    import oracle
    skip_ref = pytest.mark.skip(reason="oracle/_ref/libdic_ref.so not built here")
    for item in items:
        if "ref" in item.keywords and not oracle.have_ref():
            item.add_marker(skip_ref)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(bits(a), bits(b)), (what, a, b)
