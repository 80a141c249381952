"""oracle -- TEST INFRASTRUCTURE. ctypes bindings to the two CPU checkers.

  RefEngine     oracle/_ref/libdic_ref.so : the UNMODIFIED reference CPU engine (built only where
                /root/reference exists; the .so travels to the GPU box prebuilt)
  OracleEngine  oracle/libdic_oracle.so   : the C restatement (oracle/dic_oracle.c), builds anywhere

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (correlation_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libdic_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libdic_ref.so")

IM_NEAREST, IM_BILINEAR, IM_BICUBIC = 0, 1, 2
FM_U, FM_UV, FM_UVQ, FM_AFFINE, FM_QUAD = 0, 1, 2, 3, 4
MAXLEV = 12

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """make the restatement (always) and _ref (when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "dic_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class RefResult(C.Structure):
    _fields_ = [("params", C.c_float * 12), ("chi", C.c_float), ("number_of_points", C.c_int),
                ("iterations", C.c_int), ("error_code", C.c_int), ("error_status", C.c_int),
                ("und_center_x", C.c_float), ("und_center_y", C.c_float), ("seconds", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = RefResult._fields_ + [
        ("evaluations", C.c_int * MAXLEV), ("iterations_per_level", C.c_int * MAXLEV),
        ("points_per_level", C.c_long * MAXLEV), ("pixel_evaluations", C.c_double)]


def _result_dict(r, n_params):
    d = dict(params=np.array(r.params[:n_params], np.float32), chi=np.float32(r.chi),
             number_of_points=r.number_of_points, iterations=r.iterations,
             error_code=r.error_code, error_status=r.error_status,
             und_center=(np.float32(r.und_center_x), np.float32(r.und_center_y)),
             seconds=r.seconds)
    if isinstance(r, OrcResult):
        d["evaluations"] = list(r.evaluations)
        d["iterations_per_level"] = list(r.iterations_per_level)
        d["points_per_level"] = list(r.points_per_level)
        d["pixel_evaluations"] = r.pixel_evaluations
    return d


N_PARAMS = {FM_U: 1, FM_UV: 2, FM_UVQ: 3, FM_AFFINE: 6, FM_QUAD: 12}


class _Base:
    prefix = ""
    Result = RefResult

    def _fn(self, name, restype, *argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    def _common(self):
        P = C.c_void_p
        self._set_img = {w: self._fn(f"set_{w}_image", None, P, _u8p, C.c_int, C.c_int)
                         for w in ("und", "def", "nxt")}
        self._und_from_def = self._fn("und_from_def", None, P)
        self._def_from_nxt = self._fn("def_from_nxt", None, P)
        self._again = self._fn("correlate_again", C.c_int, P, _f32p, C.POINTER(self.Result))
        self._level_center = self._fn("level_center", None, P, C.c_int,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float))
        self._pyr = self._fn("pyramid_level", None, P, C.c_int, C.c_int, C.c_void_p,
                             C.POINTER(C.c_int), C.POINTER(C.c_int))
        self._eval = self._fn("evaluate", C.c_int, P, C.c_int, _f32p, _f32p, _f32p,
                              C.POINTER(C.c_float))
        self._solve = self._fn("solve_step", None, P, _f32p, _f32p, C.c_float, C.c_float, _f32p)
        self._destroy = self._fn("destroy", None, P)

    # -- images
    def set_image(self, which, img):
        img = np.ascontiguousarray(img, np.uint8)
        assert (img.ndim == 3 and img.shape[2] == 3) == (self.colors == 3), "image channels must match the engine's colours"
        self._set_img[which](self.h, img.reshape(img.shape[0], -1), img.shape[0], img.shape[1])

    def und_from_def(self):
        self._und_from_def(self.h)

    def def_from_nxt(self):
        self._def_from_nxt(self.h)

    # -- solve
    def correlate(self, guess, xy, center=None):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        g = np.zeros(12, np.float32)
        g[:self.n_params] = np.asarray(guess, np.float32)[:self.n_params]
        r = self.Result()
        cx, cy = center if center is not None else (0.0, 0.0)
        self._correlate(self.h, g, xy, xy.shape[0], int(center is not None), cx, cy, C.byref(r))
        return _result_dict(r, self.n_params)

    def correlate_again(self, guess):
        g = np.zeros(12, np.float32)
        g[:self.n_params] = np.asarray(guess, np.float32)[:self.n_params]
        r = self.Result()
        self._again(self.h, g, C.byref(r))
        return _result_dict(r, self.n_params)

    # -- introspection
    def set_points(self, xy, center=None):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        cx, cy = center if center is not None else (0.0, 0.0)
        self._set_points(self.h, xy, xy.shape[0], int(center is not None), cx, cy)

    def level_points(self, level):
        n = self._level_n(self.h, level)
        out = np.empty((n, 2), np.float32)
        if n:
            self._level_pts(self.h, level, out)
        return out

    def level_center(self, level):
        cx, cy = C.c_float(), C.c_float()
        self._level_center(self.h, level, C.byref(cx), C.byref(cy))
        return np.float32(cx.value), np.float32(cy.value)

    def pyramid_level(self, which, level):
        r, c = C.c_int(), C.c_int()
        self._pyr(self.h, which, level, None, C.byref(r), C.byref(c))
        out = np.empty((r.value, c.value) if self.colors == 1 else (r.value, c.value, 3), np.uint8)
        self._pyr(self.h, which, level, out.ctypes.data_as(C.c_void_p), C.byref(r), C.byref(c))
        return out

    def evaluate(self, level, params):
        n = self.n_params
        p = np.ascontiguousarray(params, np.float32)[:n].copy()
        A = np.zeros((n, n), np.float32)
        b = np.zeros(n, np.float32)
        chi = C.c_float()
        err = self._eval(self.h, level, p, A.reshape(-1), b, C.byref(chi))
        return A, b, np.float32(chi.value), err

    def solve_step(self, A_upper, b, lam, scaling):
        n = self.n_params
        dp = np.zeros(n, np.float32)
        self._solve(self.h, np.ascontiguousarray(A_upper, np.float32).reshape(-1),
                    np.ascontiguousarray(b, np.float32), lam, scaling, dp)
        return dp

    def close(self):
        if getattr(self, "h", None):
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefEngine(_Base):
    """The unmodified reference CorrelationClass behind oracle/ref_driver.cpp."""
    prefix = "ref_"
    Result = RefResult

    def __init__(self, model=FM_AFFINE, interp=IM_BICUBIC, n_threads=20, precision=1e-3,
                 max_iters=50, pyramid=(0, 1, 2), colors=1):
        if model == FM_QUAD:
            raise ValueError("the reference has no 12-parameter model (SURVEY fact 2)")
        self.lib = C.CDLL(REF_SO)
        self.n_params = N_PARAMS[model]
        self.colors = 3 if colors == 3 else 1
        P = C.c_void_p
        create = self._fn("create_color", P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                          C.c_int, C.c_int, C.c_int, C.c_int)
        self._common()
        self._correlate = self._fn("correlate", C.c_int, P, _f32p, _f32p, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.POINTER(RefResult))
        self._set_points = self._fn("set_points", None, P, _f32p, C.c_int, C.c_int,
                                    C.c_float, C.c_float)
        self._level_n = self._fn("level_num_points", C.c_int, P, C.c_int)
        self._level_pts = self._fn("level_points", None, P, C.c_int, _f32p)
        self.h = create(n_threads, interp, model, precision, max_iters, *pyramid, self.colors)

    def blob_points(self, contour):
        contour = np.ascontiguousarray(contour, np.float32).reshape(-1, 2)
        f = self._fn("blob_points", C.c_long, _f32p, C.c_int, _f32p, C.c_long)
        n = f(contour, contour.shape[0], np.zeros(2, np.float32), 0)
        if n < 0:
            return None
        out = np.zeros((n, 2), np.float32)
        f(contour, contour.shape[0], out, n)
        return out


class OracleEngine(_Base):
    """The C restatement (oracle/dic_oracle.c)."""
    prefix = "orc_"
    Result = OrcResult

    def __init__(self, model=FM_AFFINE, interp=IM_BICUBIC, n_threads=20, precision=1e-3,
                 max_iters=50, pyramid=(0, 1, 2), accum_double=False, real_threads=False, solve_double=False, colors=1):
        """accum_double / solve_double: arbitration variants -- fp64 accumulators for A, b, chi and an fp64 solve of
        the damped system, each the reference's algorithm with ONE source of fp32 rounding removed."""
        self.lib = C.CDLL(ORACLE_SO)
        self.n_params = N_PARAMS[model]
        self.colors = 3 if colors == 3 else 1
        P = C.c_void_p
        create = self._fn("create", P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
        self._common()
        self._correlate = self._fn("correlate", C.c_int, P, _f32p, _f32p, C.c_long, C.c_int,
                                   C.c_float, C.c_float, C.POINTER(OrcResult))
        self._set_points = self._fn("set_points", None, P, _f32p, C.c_long, C.c_int,
                                    C.c_float, C.c_float)
        self._level_n = self._fn("level_num_points", C.c_long, P, C.c_int)
        self._level_pts = self._fn("level_points", None, P, C.c_int, _f32p)
        self.h = create(n_threads, interp, model, precision, max_iters, *pyramid,
                        int(accum_double), int(real_threads))
        if solve_double:
            self._fn("set_solve_double", None, P, C.c_int)(self.h, 1)
        if self.colors == 3:
            self._fn("set_colors", None, P, C.c_int)(self.h, 3)


# ---- free functions of the restatement (no engine needed) ---------------------------------

def _olib():
    return C.CDLL(ORACLE_SO)


def rect_points(x0, y0, x1, y1):
    lib = _olib()
    lib.orc_rect_points.restype = C.c_long
    lib.orc_rect_points.argtypes = [C.c_int] * 4 + [_f32p, C.c_long]
    n = (x1 - x0 + 1) * (y1 - y0 + 1)
    out = np.zeros((max(n, 1), 2), np.float32)
    m = lib.orc_rect_points(x0, y0, x1, y1, out, n)
    return out[:m]


def annulus_points(r, dr, a, da, cx, cy, n_as):
    lib = _olib()
    lib.orc_annulus_points.restype = C.c_long
    lib.orc_annulus_points.argtypes = [C.c_float] * 6 + [C.c_int, _f32p, C.c_long]
    m = lib.orc_annulus_points(r, dr, a, da, cx, cy, n_as, np.zeros(2, np.float32), 0)
    out = np.zeros((max(m, 1), 2), np.float32)
    lib.orc_annulus_points(r, dr, a, da, cx, cy, n_as, out, m)
    return out[:m]


def blob_points(contour, with_triangles=False):
    lib = _olib()
    lib.orc_blob_points.restype = C.c_long
    lib.orc_blob_points.argtypes = [_f32p, C.c_int, _f32p, C.c_long, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_int)]
    contour = np.ascontiguousarray(contour, np.float32).reshape(-1, 2)
    nv = contour.shape[0]
    nt = C.c_int()
    m = lib.orc_blob_points(contour, nv, np.zeros(2, np.float32), 0, None, 0, C.byref(nt))
    if m < 0:
        return (None, None) if with_triangles else None
    out = np.zeros((max(m, 1), 2), np.float32)
    tri = np.zeros((max(nt.value, 1), 3, 2), np.float32)
    lib.orc_blob_points(contour, nv, out, m, tri.ctypes.data_as(C.c_void_p), nt.value, C.byref(nt))
    return (out[:m], tri[:nt.value]) if with_triangles else out[:m]


def seq_mean_center(xy):
    lib = _olib()
    lib.orc_seq_mean_center.restype = None
    lib.orc_seq_mean_center.argtypes = [_f32p, C.c_long, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    cx, cy = C.c_float(), C.c_float()
    lib.orc_seq_mean_center(xy, xy.shape[0], C.byref(cx), C.byref(cy))
    return np.float32(cx.value), np.float32(cy.value)


def bicubic_matrix():
    lib = _olib()
    lib.orc_bicubic_matrix.restype = C.POINTER(C.c_float * 256)
    return np.array(lib.orc_bicubic_matrix().contents, np.float32).reshape(16, 16)
