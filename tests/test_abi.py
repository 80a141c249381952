"""CPU suite: the C-ABI library loads and exports every symbol include/dic_b200.h declares."""
import ctypes
import os
import re

from correlation_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dic_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(engine.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(engine.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_result_struct_layout():
    # dic_result: 12 floats, chi, 3 ints, 2 floats, 3 x 8 ints
    assert ctypes.sizeof(engine.DicResult) == 4 * (12 + 1 + 3 + 2 + 24)


def test_no_cpu_fallback_without_gpu():
    lib = engine.load_library()
    if lib.dic_device_count() == 0:
        try:
            engine.CudaEngine(0)
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
        else:
            raise AssertionError("engine construction must fail loudly without a GPU")
