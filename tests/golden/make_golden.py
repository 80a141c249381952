"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference CPU engine
(oracle/_ref/libdic_ref.so, built from /root/reference by oracle/Makefile).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The input images are stored with the outputs so that no libm / numpy difference on another
host can move a pixel. All floats are stored as raw fp32 (bit-exact comparisons).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from correlation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")


def run_case(d, tag, und, dfm, model, interp, pyramid, threads, xy, center, guess):
    R = oracle.RefEngine(model=model, interp=interp, n_threads=threads, pyramid=pyramid)
    R.set_image("und", und)
    R.set_image("def", dfm)
    r = R.correlate(guess, xy, center=center)
    d[f"{tag}/params"] = r["params"]
    d[f"{tag}/chi"] = np.float32(r["chi"])
    d[f"{tag}/iterations"] = np.int32(r["iterations"])
    d[f"{tag}/error_code"] = np.int32(r["error_code"])
    d[f"{tag}/number_of_points"] = np.int32(r["number_of_points"])
    d[f"{tag}/center"] = np.array(r["und_center"], np.float32)
    # one evaluation at the initial guess, coarsest and finest level
    for lv in (pyramid[2], pyramid[0]):
        g = np.array(guess, np.float32).copy()
        g[:min(2, g.size)] *= np.float32(1.0 / (1 << lv))
        A, b, chi, err = R.evaluate(lv, g)
        d[f"{tag}/eval{lv}/A"] = A
        d[f"{tag}/eval{lv}/b"] = b
        d[f"{tag}/eval{lv}/chi"] = np.float32(chi)
        d[f"{tag}/eval{lv}/n"] = np.int32(R.level_points(lv).shape[0])
    R.close()


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/libdic_ref.so missing: run `make -C oracle ref` first")
    d = {}
    # ---- case A: affine / bicubic / 3 levels / rectangle (C1 in miniature)
    truth = (1.75, -0.6, .004, -.003, .002, .005)
    und, dfm = synth.make_pair(192, 192, 11, truth, center=(96, 96))
    d["A/und"], d["A/def"] = und, dfm
    xy = oracle.rect_points(40, 44, 150, 146)
    d["A/rect"] = np.array([40, 44, 150, 146], np.int32)
    center = (95.0, 95.0)
    for T in (1, 20):
        run_case(d, f"A/T{T}", und, dfm, oracle.FM_AFFINE, oracle.IM_BICUBIC, (0, 1, 2), T, xy, center,
                 np.zeros(6, np.float32))
    # pyramid levels of both images (bit-exact target of the pyramid kernel)
    R = oracle.RefEngine(pyramid=(0, 1, 2))
    R.set_image("und", und)
    R.set_image("def", dfm)
    for lv in (1, 2):
        d[f"A/pyr_und{lv}"] = R.pyramid_level(0, lv)
        d[f"A/pyr_def{lv}"] = R.pyramid_level(1, lv)
    R.set_points(xy, center=center)
    for lv in (0, 1, 2):
        d[f"A/points{lv}"] = R.level_points(lv)
    R.close()
    # ---- case B: the other fitting / interpolation models (same images, centre from the list)
    for model, mname, n in ((oracle.FM_U, "U", 1), (oracle.FM_UV, "UV", 2), (oracle.FM_UVQ, "UVQ", 3)):
        for interp, iname in ((oracle.IM_NEAREST, "nearest"), (oracle.IM_BILINEAR, "bilinear"),
                              (oracle.IM_BICUBIC, "bicubic")):
            run_case(d, f"B/{mname}_{iname}", und, dfm, model, interp, (0, 1, 1), 4, xy, None,
                     np.zeros(n, np.float32))
    run_case(d, "B/AFF_bilinear", und, dfm, oracle.FM_AFFINE, oracle.IM_BILINEAR, (0, 2, 2), 4, xy, None,
             np.zeros(6, np.float32))
    # ---- case C: blob rasterisation (polygon_class.cpp) and its sequential fp32 centre
    contour = synth.star_polygon(96.0, 96.0, 60.0, n_vertices=24, seed=5)
    pts = oracle.RefEngine().blob_points(contour)
    d["C/contour"] = contour
    d["C/points"] = pts
    R = oracle.RefEngine(pyramid=(0, 1, 2))
    R.set_image("und", und)
    R.set_image("def", dfm)
    R.set_points(pts)
    d["C/center"] = np.array(R.level_center(0), np.float32)
    R.close()
    run_case(d, "C/blob", und, dfm, oracle.FM_AFFINE, oracle.IM_BICUBIC, (0, 1, 2), 20, pts, None,
             np.zeros(6, np.float32))
    # a self-intersecting contour must be refused (polygon_class.cpp:225-229)
    bow = np.array([[10, 10], [100, 100], [100, 10], [10, 100]], np.float32)
    d["C/bowtie_is_bad"] = np.int32(oracle.RefEngine().blob_points(bow) is None)
    # ---- case D: out-of-image -> error 2 on the first evaluation (correlation_class.cpp:413-419)
    xy_edge = oracle.rect_points(2, 2, 60, 60)
    run_case(d, "D/oob", und, dfm, oracle.FM_AFFINE, oracle.IM_BICUBIC, (0, 1, 1), 4, xy_edge, (31.0, 31.0),
             np.array([-8, -8, 0, 0, 0, 0], np.float32))
    # ---- case E: max iterations (error 3)
    R = oracle.RefEngine(max_iters=1, precision=1e-9, pyramid=(0, 1, 0))
    R.set_image("und", und)
    R.set_image("def", dfm)
    r = R.correlate(np.zeros(6, np.float32), xy, center=center)
    d["E/maxit/params"] = r["params"]
    d["E/maxit/chi"] = np.float32(r["chi"])
    d["E/maxit/iterations"] = np.int32(r["iterations"])
    d["E/maxit/error_code"] = np.int32(r["error_code"])
    R.close()
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(d), "arrays")


if __name__ == "__main__":
    main()
