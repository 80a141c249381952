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


def test_host_deform_points_restates_the_reference_distort_functions():
    """managerClass::deformPoints (manager_class.cpp:2527-2600) over interpolation_class.cpp:3-43, fp32,
    left-to-right like the reference; no GPU involved."""
    import numpy as np
    from correlation_b200 import host
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 500, (50, 2)).astype(np.float32)
    p = np.array([1.5, -0.75, 0.01, -0.02, 0.03, 0.005], np.float32)
    cx, cy = np.float32(250.0), np.float32(240.0)
    x, y = xy[:, 0], xy[:, 1]
    want = {
        0: (x + p[0], y),
        1: (x + p[0], y + p[1]),
        2: (x + p[0] - (y - cy) * p[2], y + p[1] + (x - cx) * p[2]),
        3: (x + p[0] + (x - cx) * p[2] + (y - cy) * p[3], y + p[1] + (x - cx) * p[4] + (y - cy) * p[5]),
    }
    for model, (wx, wy) in want.items():
        got = host.deform_points(model, p, (cx, cy), xy)
        assert np.array_equal(got[:, 0], wx.astype(np.float32)) and np.array_equal(got[:, 1], wy.astype(np.float32)), model


def test_header_is_plain_c_and_the_cpp_facade_compiles_against_it(tmp_path):
    """include/dic_b200.h is the drop-in boundary: it must be includable from C (no C++ or torch types in the
    signatures), and the reference-side binding INTEGRATION.md documents (host/dic_cuda_class.hpp, incl. the pair
    staging and async batch methods) must compile against it with nothing but the standard library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "dic_b200.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / "use_facade.cpp"
    src.write_text('#include "dic_cuda_class.hpp"\n'
                   "int probe(CudaClass &c, const unsigned char *u, const unsigned char *d, float *g, CorrelationResult *r) {\n"
                   "  c.stageNextPair(u, d, 64, 64); c.advancePair();\n"
                   "  c.correlateBatchAsync(0, 4, g); return c.correlateBatchWait(0, 4, g, r);\n}\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "correlation_b200", "host"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
