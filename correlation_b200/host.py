"""ctypes binding of libdic_host.so: the headless C++ host (correlation_b200/host/dic_manager.hpp),
i.e. what the reference's managerClass does around CudaClass -- frame loop, image rotation,
sector arithmetic, initial-guess extrapolation, CSV report (manager_class.cpp:1297-1541,
2430-2525, 2602-2707)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdic_host.so")
DOMAIN_RECT, DOMAIN_ANNULUS, DOMAIN_BLOB = 0, 1, 2


class HostConfig(C.Structure):
    _fields_ = [("domain_type", C.c_int), ("rect", C.c_float * 4), ("subdivisions", C.c_int * 2),
                ("annulus", C.c_float * 4), ("contour_xy", C.c_void_p), ("n_contour", C.c_int),
                ("model", C.c_int), ("interpolation", C.c_int), ("pyramid", C.c_int * 3),
                ("precision", C.c_float), ("max_iters", C.c_int), ("deformation_description", C.c_int),
                ("reference_image", C.c_int), ("global_initial_guess", C.c_float * 12),
                ("arith_mode", C.c_int), ("batch_sectors", C.c_int), ("device", C.c_int),
                ("error_handling_mode", C.c_int)]


ROW_DTYPE = np.dtype([("frame", np.int32), ("sector", np.int32), ("und_center_x", np.float32),
                      ("und_center_y", np.float32), ("def_center_x", np.float32), ("def_center_y", np.float32),
                      ("def_angle", np.float32), ("params", np.float32, (12,)), ("initial_guess", np.float32, (12,)),
                      ("chi", np.float32), ("number_of_points", np.int32), ("iterations", np.int32),
                      ("error_code", np.int32), ("und_global_center_x", np.float32), ("und_global_center_y", np.float32),
                      ("def_global_center_x", np.float32), ("def_global_center_y", np.float32),
                      ("def_global_angle", np.float32), ("und_angle", np.float32)])
ERROR_STOP_ALL, ERROR_STOP_FRAME, ERROR_CONTINUE = 0, 1, 2  # errorHandlingModeEnum, enums.hpp:80-85


def run_sequence(frames, *, rect=None, subdivisions=(1, 1), annulus=None, contour=None, model=3,
                 interpolation=2, pyramid=(0, 1, 2), precision=1e-3, max_iters=50, deformation=2,
                 reference=0, guess=None, arith_mode=0, batch=False, device=0, on_error=ERROR_STOP_ALL):
    """frames: list of equally sized uint8 2-D arrays (host). Returns dict(csv, seconds, rows, error)."""
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB_PATH)
    lib.dic_host_run.restype = C.c_int
    # explicit prototype: csv_cap is a long long -- an undeclared Python int travels as a 32-bit int and leaves
    # the upper half of the register undefined (seen as an intermittently empty report)
    lib.dic_host_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
    rows, cols = frames[0].shape
    c = HostConfig()
    keep = None
    if rect is not None:
        c.domain_type = DOMAIN_RECT
        c.rect[:] = [float(v) for v in rect]
    elif annulus is not None:
        c.domain_type = DOMAIN_ANNULUS
        c.annulus[:] = [float(v) for v in annulus]
    else:
        c.domain_type = DOMAIN_BLOB
        keep = np.ascontiguousarray(contour, np.float32).reshape(-1, 2)
        c.contour_xy = keep.ctypes.data
        c.n_contour = keep.shape[0]
    c.subdivisions[:] = [int(v) for v in subdivisions]
    c.model, c.interpolation = int(model), int(interpolation)
    c.pyramid[:] = [int(v) for v in pyramid]
    c.precision, c.max_iters = float(precision), int(max_iters)
    c.deformation_description, c.reference_image = int(deformation), int(reference)
    g = np.zeros(12, np.float32)
    if guess is not None:
        g[:len(guess)] = guess
    c.global_initial_guess[:] = g.tolist()
    c.arith_mode, c.batch_sectors, c.device = int(arith_mode), int(bool(batch)), int(device)
    c.error_handling_mode = int(on_error)
    ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    n_sectors = int(subdivisions[0]) * int(subdivisions[1]) if contour is None else 1
    csv = C.create_string_buffer(max(1 << 16, 600 * n_sectors * len(frames)))
    need, secs, written = C.c_longlong(), C.c_double(), C.c_int()
    out = np.zeros(n_sectors, ROW_DTYPE)
    rc = lib.dic_host_run(C.addressof(c), C.addressof(ptrs), len(frames), rows, cols, C.addressof(csv), len(csv),
                          C.addressof(need), C.addressof(secs), out.ctypes.data, n_sectors, C.addressof(written))
    if rc < 0:
        raise RuntimeError("dic_host_run failed (no GPU?)")
    return dict(csv=csv.value.decode(), seconds=secs.value, rows=out[:written.value], error=rc)


def deform_points(model, params, center, xy):
    """managerClass::deformPoints (manager_class.cpp:2527-2600): contour points carried by a result."""
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB_PATH)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    p = np.zeros(12, np.float32)
    p[: len(params)] = params
    out = np.zeros_like(xy)
    lib.dic_host_deform_points.restype = C.c_int
    lib.dic_host_deform_points.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
    rc = lib.dic_host_deform_points(int(model), p.ctypes.data, float(center[0]), float(center[1]), xy.ctypes.data, len(xy), out.ctypes.data)
    if rc != 0:
        raise ValueError("dic_host_deform_points: bad argument")
    return out


def parse_report(csv_text):
    """CSV report -> list of dict rows (floats where possible)."""
    lines = [l for l in csv_text.strip().split("\n") if l]
    hdr = lines[0].split(",")
    rows = []
    for l in lines[1:]:
        vals = l.split(",")
        row = {}
        for k, v in zip(hdr, vals):
            try:
                row[k] = float(v)
            except ValueError:
                row[k] = v
        rows.append(row)
    return hdr, rows
