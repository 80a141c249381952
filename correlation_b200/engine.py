"""ctypes binding of libdic_b200.so (include/dic_b200.h) for tests and bench.py.

`CudaEngine` keeps the method names of the reference facade `CudaClass`
(cuda_class.cuh:46-79): resetImagePyramids, resetNextPyramid, makeUndPyramidFromDef,
makeDefPyramidFromNxt, resetPolygon (rect / annular / blob), updatePolygon, correlate,
getUndXY0ToCPU, getDefXY0ToCPU -- taking numpy arrays where the reference takes file
paths / cv::Mat / v_points. There is NO CPU fallback: construction raises if the CUDA
library is missing or no device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DIC_B200_LIB: an alternative build of the same library (A/B runs of kernel variants, tools/ab.sh)
LIB_PATH = os.environ.get("DIC_B200_LIB") or os.path.join(HERE, "libdic_b200.so")

MAX_PARAMS, MAX_LEVELS = 12, 8
FM_U, FM_UV, FM_UVQ, FM_UVUxUyVxVy, FM_QUADRATIC = range(5)
IM_NEAREST, IM_BILINEAR, IM_BICUBIC = range(3)
DEF_STRICT_LAGRANGIAN, DEF_LAGRANGIAN, DEF_EULERIAN = range(3)
MODE_PARITY, MODE_FAST = 0, 1
CENTER_REFERENCE, CENTER_EXACT = 0, 1
N_PARAMS = {FM_U: 1, FM_UV: 2, FM_UVQ: 3, FM_UVUxUyVxVy: 6, FM_QUADRATIC: 12}

ERROR_NAMES = ["error_none", "error_model_out_of_image", "error_interpolation_out_of_image",
               "error_correlation_max_iters_reached", "error_bad_domain", "error_cuSolver",
               "error_cuda", "error_multiThread", "error_bad_argument"]


class DicResult(C.Structure):
    """dic_result (domains.hpp:110-118 CorrelationResult, widened)."""
    _fields_ = [("resultingParameters", C.c_float * MAX_PARAMS), ("chi", C.c_float),
                ("numberOfPoints", C.c_int), ("iterations", C.c_int), ("errorCode", C.c_int),
                ("undCenterX", C.c_float), ("undCenterY", C.c_float),
                ("iterationsPerLevel", C.c_int * MAX_LEVELS),
                ("evaluationsPerLevel", C.c_int * MAX_LEVELS),
                ("pointsPerLevel", C.c_int * MAX_LEVELS)]

    def as_dict(self, n_params):
        ev, pts = list(self.evaluationsPerLevel), list(self.pointsPerLevel)
        return dict(params=np.array(self.resultingParameters[:n_params], np.float32),
                    chi=np.float32(self.chi), number_of_points=self.numberOfPoints,
                    iterations=self.iterations, error_code=self.errorCode,
                    und_center=(np.float32(self.undCenterX), np.float32(self.undCenterY)),
                    iterations_per_level=list(self.iterationsPerLevel), evaluations=ev,
                    points_per_level=pts,
                    pixel_evaluations=float(sum(e * p for e, p in zip(ev, pts))))


RESULT_DTYPE = np.dtype([("resultingParameters", np.float32, (MAX_PARAMS,)), ("chi", np.float32),
                         ("numberOfPoints", np.int32), ("iterations", np.int32), ("errorCode", np.int32),
                         ("undCenterX", np.float32), ("undCenterY", np.float32),
                         ("iterationsPerLevel", np.int32, (MAX_LEVELS,)),
                         ("evaluationsPerLevel", np.int32, (MAX_LEVELS,)),
                         ("pointsPerLevel", np.int32, (MAX_LEVELS,))])
assert RESULT_DTYPE.itemsize == C.sizeof(DicResult)

_lib = None

EXPORTS = [
    "dic_device_count", "dic_create", "dic_destroy", "dic_last_error", "dic_set_max_iters",
    "dic_set_precision", "dic_set_fitting_model", "dic_set_interpolation_model",
    "dic_set_arith_mode", "dic_set_center_mode", "dic_set_kernel_variant", "dic_reset_image_pyramids",
    "dic_reset_image_pyramids_device", "dic_reset_next_pyramid", "dic_reset_next_pyramid_async", "dic_reset_next_pyramid_device",
    "dic_reset_def_pyramid", "dic_reset_def_pyramid_device", "dic_make_und_pyramid_from_def",
    "dic_make_def_pyramid_from_nxt", "dic_reset_polygon_rect", "dic_reset_polygon_annular",
    "dic_reset_polygon_blob", "dic_reset_polygon_rect_band", "dic_reset_polygon_rect_grid", "dic_set_cluster_mode", "dic_set_batch_queue", "dic_correlate_batch_async", "dic_correlate_batch_wait", "dic_last_cluster_size", "dic_reset_polygon_points", "dic_set_polygon_center",
    "dic_stage_next_pair", "dic_stage_next_pair_rows", "dic_advance_pair",
    "dic_update_polygon", "dic_rowsplit_mailbox_handle", "dic_rowsplit_connect", "dic_rowsplit_disconnect", "dic_correlate", "dic_correlate_batch", "dic_correlate_async",
    "dic_correlate_wait", "dic_get_und_xy0", "dic_get_def_xy0", "dic_get_pyramid_level",
    "dic_get_level_points", "dic_get_level_center", "dic_evaluate", "dic_solve_step",
    "dic_last_correlate_ms", "dic_last_step_ms", "dic_get_timeline", "dic_pipe_trace", "dic_get_cta_times", "dic_get_cta_smids", "dic_kernel_launches", "dic_correlation_stream", "dic_synchronize",
]


def load_library():
    """dlopen libdic_b200.so and declare the prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P, I, F, I64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    fp = C.POINTER(C.c_float)
    sig = {
        "dic_device_count": (I, []),
        "dic_create": (P, [I]),
        "dic_destroy": (None, [P]),
        "dic_last_error": (C.c_char_p, [P]),
        "dic_set_max_iters": (I, [P, I]),
        "dic_set_precision": (I, [P, F]),
        "dic_set_fitting_model": (I, [P, I]),
        "dic_set_interpolation_model": (I, [P, I]),
        "dic_set_arith_mode": (I, [P, I]),
        "dic_set_center_mode": (I, [P, I]),
        "dic_set_kernel_variant": (I, [P, I]),
        "dic_reset_image_pyramids": (I, [P, P, P, P, I, I, I, I, I, I]),
        "dic_reset_image_pyramids_device": (I, [P, P, P, P, I, I, I, I, I, I]),
        "dic_reset_next_pyramid": (I, [P, P, I, I]),
        "dic_reset_next_pyramid_async": (I, [P, P, I, I]),
        "dic_reset_next_pyramid_device": (I, [P, P, I, I, I]),
        "dic_reset_def_pyramid": (I, [P, P, I, I]),
        "dic_reset_def_pyramid_device": (I, [P, P, I, I, I]),
        "dic_make_und_pyramid_from_def": (I, [P]),
        "dic_stage_next_pair": (I, [P, P, P, I, I]),
        "dic_stage_next_pair_rows": (I, [P, P, P, I, I, I, I]),
        "dic_advance_pair": (I, [P]),
        "dic_make_def_pyramid_from_nxt": (I, [P]),
        "dic_reset_polygon_rect": (I, [P, I, I, I, I, I]),
        "dic_reset_polygon_annular": (I, [P, I, F, F, F, F, F, F, I]),
        "dic_reset_polygon_blob": (I, [P, I, P, I]),
        "dic_reset_polygon_points": (I, [P, I, P, I64, I, F, F]),
        "dic_reset_polygon_rect_band": (I, [P, I, I, I, I, I, I, I]),
        "dic_reset_polygon_rect_grid": (I, [P, I, I, P]),
        "dic_set_cluster_mode": (I, [P, I]),
        "dic_set_batch_queue": (I, [P, I]),
        "dic_pipe_trace": (I, [P, I, P, P, I, P]),
        "dic_correlate_batch_async": (I, [P, I, I, P]),
        "dic_correlate_batch_wait": (I, [P, I, I, P, P]),
        "dic_last_cluster_size": (I, [P]),
        "dic_rowsplit_mailbox_handle": (I, [P, P, I]),
        "dic_rowsplit_connect": (I, [P, I, I, P, I]),
        "dic_rowsplit_disconnect": (I, [P]),
        "dic_set_polygon_center": (I, [P, I, F, F]),
        "dic_update_polygon": (I, [P, I, I]),
        "dic_correlate": (I, [P, I, P, C.POINTER(DicResult)]),
        "dic_correlate_batch": (I, [P, I, I, P, P]),
        "dic_correlate_async": (I, [P, I, P]),
        "dic_correlate_wait": (I, [P, I, P, C.POINTER(DicResult)]),
        "dic_get_und_xy0": (I, [P, I, P, I64, C.POINTER(I64)]),
        "dic_get_def_xy0": (I, [P, I, P, I64, C.POINTER(I64)]),
        "dic_get_pyramid_level": (I, [P, I, I, P, C.POINTER(I), C.POINTER(I)]),
        "dic_get_level_points": (I, [P, I, I, P, I64, C.POINTER(I64)]),
        "dic_get_level_center": (I, [P, I, I, fp, fp]),
        "dic_evaluate": (I, [P, I, I, P, P, P, fp, C.POINTER(I)]),
        "dic_solve_step": (I, [P, P, P, F, F, P]),
        "dic_last_correlate_ms": (F, [P]),
        "dic_last_step_ms": (F, [P]),
        "dic_get_timeline": (I, [P, P, I]),
        "dic_get_cta_times": (I, [P, P, I]),
        "dic_get_cta_smids": (I, [P, P, I]),
        "dic_kernel_launches": (I64, [P]),
        "dic_correlation_stream": (P, [P]),
        "dic_synchronize": (I, [P]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib


class DicError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        name = ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else str(code)
        super().__init__(f"{name} ({code}) {msg}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CudaEngine:
    """Python face of the C-ABI; method names follow CudaClass (cuda_class.cuh:46-79)."""

    def __init__(self, device=0, fitting_model=FM_UVUxUyVxVy, interpolation_model=IM_BICUBIC,
                 max_iters=50, precision=1e-3, arith_mode=MODE_PARITY, center_mode=CENTER_REFERENCE):
        self.lib = load_library()
        if self.lib.dic_device_count() <= 0:
            raise RuntimeError("no CUDA device visible (there is no CPU fallback)")
        self.h = self.lib.dic_create(device)
        if not self.h:
            raise RuntimeError(f"dic_create({device}) failed")
        self.set_fitting_model(fitting_model)
        self.set_interpolation_model(interpolation_model)
        self.set_max_iters(max_iters)
        self.set_precision(precision)
        self.set_arith_mode(arith_mode)
        self.set_center_mode(center_mode)

    # -- plumbing
    def _ck(self, rc, soft=()):
        if rc != 0 and rc not in soft:
            raise DicError(rc, self.lib.dic_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.dic_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setters (cuda_class.cu:95-101, 475-496)
    def set_max_iters(self, v):
        self._ck(self.lib.dic_set_max_iters(self.h, int(v)))

    def set_precision(self, v):
        self._ck(self.lib.dic_set_precision(self.h, float(v)))

    def set_fitting_model(self, m):
        self._ck(self.lib.dic_set_fitting_model(self.h, int(m)))
        self.model, self.n_params = int(m), N_PARAMS[int(m)]

    def set_interpolation_model(self, m):
        self._ck(self.lib.dic_set_interpolation_model(self.h, int(m)))

    def set_arith_mode(self, m):
        self._ck(self.lib.dic_set_arith_mode(self.h, int(m)))

    def set_center_mode(self, m):
        self._ck(self.lib.dic_set_center_mode(self.h, int(m)))

    def set_kernel_variant(self, v):
        self._ck(self.lib.dic_set_kernel_variant(self.h, int(v)))

    # -- images
    @staticmethod
    def _img(a):
        a = np.ascontiguousarray(a, np.uint8)
        assert a.ndim == 2 or (a.ndim == 3 and a.shape[2] == 3), "rows x cols (monochrome) or rows x cols x 3 (colour)"
        return a

    def resetImagePyramids(self, und, dfm, nxt=None, pyramid=(0, 1, 2)):
        und, dfm = self._img(und), self._img(dfm)
        nx = self._img(nxt) if nxt is not None else None
        self.channels = 3 if und.ndim == 3 else 1
        self._ck(self.lib.dic_reset_image_pyramids(
            self.h, _ptr(und), _ptr(dfm), _ptr(nx) if nx is not None else None,
            und.shape[0], und.shape[1], self.channels, *pyramid))

    def resetImagePyramidsDevice(self, und_ptr, def_ptr, nxt_ptr, rows, cols, pitch, pyramid=(0, 1, 2)):
        self.channels = 1
        self._ck(self.lib.dic_reset_image_pyramids_device(self.h, und_ptr, def_ptr, nxt_ptr, rows, cols,
                                                          pitch, *pyramid))

    def resetNextPyramid(self, nxt):
        nxt = self._img(nxt)
        self._ck(self.lib.dic_reset_next_pyramid(self.h, _ptr(nxt), nxt.shape[0], nxt.shape[1]))

    def resetDefPyramid(self, dfm):
        dfm = self._img(dfm)
        self._ck(self.lib.dic_reset_def_pyramid(self.h, _ptr(dfm), dfm.shape[0], dfm.shape[1]))

    def stageNextPair(self, und_ptr, def_ptr, rows, cols, row_range=None):
        """Enqueue upload + pyramids of the next pair from (pinned) host pointers; returns at once.
        row_range = (begin, end): transfer only that band of rows (see dic_stage_next_pair_rows)."""
        if row_range is None:
            self._ck(self.lib.dic_stage_next_pair(self.h, und_ptr, def_ptr, rows, cols))
        else:
            self._ck(self.lib.dic_stage_next_pair_rows(self.h, und_ptr, def_ptr, rows, cols, int(row_range[0]), int(row_range[1])))

    def advancePair(self):
        self._ck(self.lib.dic_advance_pair(self.h))

    def makeUndPyramidFromDef(self):
        self._ck(self.lib.dic_make_und_pyramid_from_def(self.h))

    def makeDefPyramidFromNxt(self):
        self._ck(self.lib.dic_make_def_pyramid_from_nxt(self.h))

    # -- domains: the three resetPolygon overloads (cuda_class.cu:574-605)
    def resetPolygon(self, iSector, *args):
        if len(args) == 4:
            return self.resetPolygonRect(iSector, *args)
        if len(args) == 7:
            return self.resetPolygonAnnular(iSector, *args)
        if len(args) == 1:
            return self.resetPolygonBlob(iSector, args[0])
        raise TypeError("resetPolygon(iSector, x0,y0,x1,y1 | r,dr,a,da,cx,cy,as | contour)")

    def resetPolygonRect(self, iSector, x0, y0, x1, y1):
        return self._ck(self.lib.dic_reset_polygon_rect(self.h, iSector, int(x0), int(y0), int(x1), int(y1)),
                        soft=(4,))

    def resetPolygonRectBand(self, iSector, x0, y0, x1, y1, band_y0, band_y1):
        return self._ck(self.lib.dic_reset_polygon_rect_band(self.h, iSector, int(x0), int(y0), int(x1), int(y1),
                                                             int(band_y0), int(band_y1)), soft=(4,))

    def resetPolygonRectGrid(self, first_sector, boxes):
        """boxes: (n, 4) int array of x0, y0, x1, y1 -- n rectangles in one go (dic_reset_polygon_rect_grid)."""
        b = np.ascontiguousarray(boxes, np.int32).reshape(-1, 4)
        return self._ck(self.lib.dic_reset_polygon_rect_grid(self.h, int(first_sector), b.shape[0], _ptr(b)), soft=(4,))

    def pipe_trace(self, on, read=True):
        """Arm (on=True) / disarm the staged-pair pipeline marks; returns (stage_ms [n, 4], solve_ms [m, 2]) of the
        marks armed before this call (dic_pipe_trace)."""
        st = np.zeros((48, 4), np.float32); so = np.zeros((48, 2), np.float32); m = C.c_int(0)
        n = self.lib.dic_pipe_trace(self.h, int(bool(on)), st.ctypes.data if read else None, so.ctypes.data if read else None, 48, C.byref(m))
        return st[:n], so[:m.value]

    def set_batch_queue(self, mode):
        """0 automatic, 1 resident CTAs + ticket queue, 2 one CTA per sector (dic_set_batch_queue)"""
        self._ck(self.lib.dic_set_batch_queue(self.h, int(mode)))

    def set_cluster_mode(self, mode):
        self._ck(self.lib.dic_set_cluster_mode(self.h, int(mode)))

    def last_cluster_size(self):
        return int(self.lib.dic_last_cluster_size(self.h))

    def rowsplit_handle(self):
        buf = np.zeros(64, np.uint8)
        self._ck(self.lib.dic_rowsplit_mailbox_handle(self.h, _ptr(buf), 64))
        return buf

    def rowsplit_connect(self, rank, world, handles):
        h = np.ascontiguousarray(handles, np.uint8).reshape(world, 64)
        self._ck(self.lib.dic_rowsplit_connect(self.h, rank, world, _ptr(h), 64))

    def rowsplit_disconnect(self):
        self._ck(self.lib.dic_rowsplit_disconnect(self.h))

    def resetPolygonAnnular(self, iSector, r, dr, a, da, cx, cy, n_as):
        return self._ck(self.lib.dic_reset_polygon_annular(self.h, iSector, r, dr, a, da, cx, cy, int(n_as)),
                        soft=(4,))

    def resetPolygonBlob(self, iSector, contour):
        c = np.ascontiguousarray(contour, np.float32).reshape(-1, 2)
        return self._ck(self.lib.dic_reset_polygon_blob(self.h, iSector, _ptr(c), c.shape[0]), soft=(4,))

    def resetPolygonPoints(self, iSector, xy, center=None):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        cx, cy = center if center is not None else (0.0, 0.0)
        return self._ck(self.lib.dic_reset_polygon_points(self.h, iSector, _ptr(xy), xy.shape[0],
                                                          int(center is not None), cx, cy), soft=(4,))

    def setPolygonCenter(self, iSector, cx, cy):
        self._ck(self.lib.dic_set_polygon_center(self.h, iSector, cx, cy))

    def updatePolygon(self, iSector, deformation_description):
        self._ck(self.lib.dic_update_polygon(self.h, iSector, int(deformation_description)))

    # -- solve
    def correlate(self, iSector, guess):
        g = np.zeros(self.n_params, np.float32)
        g[:] = np.asarray(guess, np.float32)[:self.n_params]
        r = DicResult()
        self._ck(self.lib.dic_correlate(self.h, iSector, _ptr(g), C.byref(r)), soft=(1, 2, 3, 5))
        return r.as_dict(self.n_params)

    def correlate_raw(self, iSector, guess_inout, result):
        """Lowest-overhead form: guess_inout (float32[n_params]) and result (DicResult) are caller-owned
        and reused; returns the error code."""
        return self.lib.dic_correlate(self.h, iSector, guess_inout.ctypes.data, C.byref(result))

    def correlate_async(self, iSector, guess):
        g = np.zeros(self.n_params, np.float32)
        g[:] = np.asarray(guess, np.float32)[:self.n_params]
        self._ck(self.lib.dic_correlate_async(self.h, iSector, _ptr(g)))

    def correlate_wait(self, iSector):
        r = DicResult()
        g = np.zeros(self.n_params, np.float32)
        self._ck(self.lib.dic_correlate_wait(self.h, iSector, _ptr(g), C.byref(r)), soft=(1, 2, 3, 5))
        return r.as_dict(self.n_params)

    def correlate_batch_raw(self, first_sector, guesses, out=None):
        """One launch for len(guesses) consecutive sectors. Returns (guesses_out, results) as numpy
        arrays (results: structured RESULT_DTYPE); no per-sector Python objects."""
        g = np.ascontiguousarray(guesses, np.float32).reshape(-1, self.n_params).copy()
        n = g.shape[0]
        res = out if out is not None else np.zeros(n, RESULT_DTYPE)
        self._ck(self.lib.dic_correlate_batch(self.h, first_sector, n, _ptr(g), _ptr(res)), soft=(1, 2, 3, 5))
        return g, res

    def correlate_batch(self, first_sector, guesses):
        _, res = self.correlate_batch_raw(first_sector, guesses)
        out = []
        for r in res:
            ev, pts = r["evaluationsPerLevel"].tolist(), r["pointsPerLevel"].tolist()
            out.append(dict(params=r["resultingParameters"][:self.n_params].copy(), chi=np.float32(r["chi"]),
                            number_of_points=int(r["numberOfPoints"]), iterations=int(r["iterations"]),
                            error_code=int(r["errorCode"]),
                            und_center=(np.float32(r["undCenterX"]), np.float32(r["undCenterY"])),
                            iterations_per_level=r["iterationsPerLevel"].tolist(), evaluations=ev,
                            points_per_level=pts,
                            pixel_evaluations=float(sum(e * p for e, p in zip(ev, pts)))))
        return out

    @staticmethod
    def pixel_evaluations(res):
        """sum over sectors and levels of points x evaluations for a RESULT_DTYPE array"""
        return float((res["evaluationsPerLevel"].astype(np.float64) * res["pointsPerLevel"]).sum())

    # -- read-back
    def _points(self, fn, iSector, *lead):
        need = C.c_int64()
        self._ck(fn(self.h, iSector, *lead, None, 0, C.byref(need)))
        out = np.zeros((max(need.value, 1), 2), np.float32)
        self._ck(fn(self.h, iSector, *lead, _ptr(out), need.value, C.byref(need)))
        return out[:need.value]

    def getUndXY0ToCPU(self, iSector):
        return self._points(self.lib.dic_get_und_xy0, iSector)

    def getDefXY0ToCPU(self, iSector):
        return self._points(self.lib.dic_get_def_xy0, iSector)

    def level_points(self, iSector, level):
        return self._points(self.lib.dic_get_level_points, iSector, level)

    def level_center(self, iSector, level):
        cx, cy = C.c_float(), C.c_float()
        self._ck(self.lib.dic_get_level_center(self.h, iSector, level, C.byref(cx), C.byref(cy)))
        return np.float32(cx.value), np.float32(cy.value)

    def pyramid_level(self, which, level):
        r, c = C.c_int(), C.c_int()
        self._ck(self.lib.dic_get_pyramid_level(self.h, which, level, None, C.byref(r), C.byref(c)))
        ch = getattr(self, "channels", 1)
        out = np.zeros((r.value, c.value) if ch == 1 else (r.value, c.value, 3), np.uint8)
        self._ck(self.lib.dic_get_pyramid_level(self.h, which, level, _ptr(out), C.byref(r), C.byref(c)))
        return out

    def evaluate(self, iSector, level, params):
        n = self.n_params
        p = np.ascontiguousarray(params, np.float32)[:n].copy()
        A = np.zeros((n, n), np.float32)
        b = np.zeros(n, np.float32)
        chi, oob = C.c_float(), C.c_int()
        self._ck(self.lib.dic_evaluate(self.h, iSector, level, _ptr(p), _ptr(A), _ptr(b), C.byref(chi),
                                       C.byref(oob)))
        return A, b, np.float32(chi.value), oob.value

    def solve_step(self, A_upper, b, lam, scaling):
        n = self.n_params
        A = np.ascontiguousarray(A_upper, np.float32).reshape(n, n)
        bb = np.ascontiguousarray(b, np.float32)
        dp = np.zeros(n, np.float32)
        self._ck(self.lib.dic_solve_step(self.h, _ptr(A), _ptr(bb), lam, scaling, _ptr(dp)))
        return dp

    def last_correlate_ms(self):
        return float(self.lib.dic_last_correlate_ms(self.h))

    def last_step_ms(self):
        return float(self.lib.dic_last_step_ms(self.h))

    def timeline(self):
        m = np.zeros((129, 4), np.uint64)
        n = self.lib.dic_get_timeline(self.h, _ptr(m), 129)
        self.slow_units = int(m[n, 0]) if n < 129 else -1
        return m[:n].astype(np.int64)

    def cta_times(self, n=296):
        m = np.zeros(n, np.uint64)
        k = self.lib.dic_get_cta_times(self.h, _ptr(m), n)
        return m[:k].astype(np.int64)

    def cta_smids(self, n=296):
        m = np.zeros(n, np.uint32)
        k = self.lib.dic_get_cta_smids(self.h, _ptr(m), n)
        return m[:k].astype(np.int64)

    def kernel_launches(self):
        return int(self.lib.dic_kernel_launches(self.h))

    def synchronize(self):
        self._ck(self.lib.dic_synchronize(self.h))
