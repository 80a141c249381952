"""CPU suite: the oracle restatement against the golden vectors of the unmodified reference
(tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from oracle/_ref) and, where the
compiled reference is present, against the reference itself -- bit for bit."""
import os

import numpy as np
import pytest

import oracle
from conftest import assert_bit_equal
from correlation_b200 import synth

MODELS = {"U": oracle.FM_U, "UV": oracle.FM_UV, "UVQ": oracle.FM_UVQ, "AFF": oracle.FM_AFFINE}
INTERPS = {"nearest": oracle.IM_NEAREST, "bilinear": oracle.IM_BILINEAR, "bicubic": oracle.IM_BICUBIC}


def _engine(g, **kw):
    e = oracle.OracleEngine(**kw)
    e.set_image("und", g["A/und"])
    e.set_image("def", g["A/def"])
    return e


def _check_case(g, tag, r):
    assert_bit_equal(r["params"], g[f"{tag}/params"], tag + " params")
    assert_bit_equal(r["chi"], g[f"{tag}/chi"], tag + " chi")
    assert r["iterations"] == int(g[f"{tag}/iterations"])
    assert r["error_code"] == int(g[f"{tag}/error_code"])
    assert r["number_of_points"] == int(g[f"{tag}/number_of_points"])
    assert_bit_equal(np.array(r["und_center"]), g[f"{tag}/center"], tag + " centre")


@pytest.mark.parametrize("threads", [1, 20])
def test_affine_bicubic_matches_reference_golden(golden, threads):
    g = golden
    x0, y0, x1, y1 = g["A/rect"]
    xy = oracle.rect_points(x0, y0, x1, y1)
    e = _engine(g, n_threads=threads, pyramid=(0, 1, 2))
    r = e.correlate(np.zeros(6), xy, center=(95.0, 95.0))
    _check_case(g, f"A/T{threads}", r)
    for lv in (2, 0):
        A, b, chi, err = e.evaluate(lv, np.zeros(6))
        assert_bit_equal(A, g[f"A/T{threads}/eval{lv}/A"])
        assert_bit_equal(b, g[f"A/T{threads}/eval{lv}/b"])
        assert_bit_equal(chi, g[f"A/T{threads}/eval{lv}/chi"])


def test_pyramid_and_lists_match_reference_golden(golden):
    g = golden
    e = _engine(g, pyramid=(0, 1, 2))
    for lv in (1, 2):
        assert np.array_equal(e.pyramid_level(0, lv), g[f"A/pyr_und{lv}"])
        assert np.array_equal(e.pyramid_level(1, lv), g[f"A/pyr_def{lv}"])
    x0, y0, x1, y1 = g["A/rect"]
    e.set_points(oracle.rect_points(x0, y0, x1, y1), center=(95.0, 95.0))
    for lv in (0, 1, 2):
        assert_bit_equal(e.level_points(lv), g[f"A/points{lv}"])


@pytest.mark.parametrize("mname", ["U", "UV", "UVQ"])
@pytest.mark.parametrize("iname", ["nearest", "bilinear", "bicubic"])
def test_other_models_match_reference_golden(golden, mname, iname):
    g = golden
    x0, y0, x1, y1 = g["A/rect"]
    xy = oracle.rect_points(x0, y0, x1, y1)
    e = _engine(g, model=MODELS[mname], interp=INTERPS[iname], n_threads=4, pyramid=(0, 1, 1))
    r = e.correlate(np.zeros(e.n_params), xy)
    _check_case(g, f"B/{mname}_{iname}", r)


def test_affine_bilinear_step2_matches_reference_golden(golden):
    g = golden
    x0, y0, x1, y1 = g["A/rect"]
    e = _engine(g, interp=oracle.IM_BILINEAR, n_threads=4, pyramid=(0, 2, 2))
    _check_case(g, "B/AFF_bilinear", e.correlate(np.zeros(6), oracle.rect_points(x0, y0, x1, y1)))


def test_blob_rasterisation_matches_reference_golden(golden):
    g = golden
    pts, tri = oracle.blob_points(g["C/contour"], with_triangles=True)
    assert_bit_equal(pts, g["C/points"])
    assert tri.shape[0] == g["C/contour"].shape[0] - 2  # ear clipping: n - 2 triangles
    assert_bit_equal(np.array(oracle.seq_mean_center(pts)), g["C/center"])
    bow = np.array([[10, 10], [100, 100], [100, 10], [10, 100]], np.float32)
    assert oracle.blob_points(bow) is None and int(g["C/bowtie_is_bad"]) == 1
    e = _engine(g, n_threads=20, pyramid=(0, 1, 2))
    _check_case(g, "C/blob", e.correlate(np.zeros(6), pts))


def test_error_paths_match_reference_golden(golden):
    g = golden
    e = _engine(g, n_threads=4, pyramid=(0, 1, 1))
    r = e.correlate(np.array([-8, -8, 0, 0, 0, 0], np.float32), oracle.rect_points(2, 2, 60, 60),
                    center=(31.0, 31.0))
    assert r["error_code"] == int(g["D/oob/error_code"]) == 2
    assert_bit_equal(r["params"], g["D/oob/params"])
    assert_bit_equal(r["chi"], g["D/oob/chi"])
    x0, y0, x1, y1 = g["A/rect"]
    e = _engine(g, n_threads=20, pyramid=(0, 1, 0), max_iters=1, precision=1e-9)
    r = e.correlate(np.zeros(6), oracle.rect_points(x0, y0, x1, y1), center=(95.0, 95.0))
    assert r["error_code"] == int(g["E/maxit/error_code"]) == 3
    assert_bit_equal(r["params"], g["E/maxit/params"])
    assert_bit_equal(r["chi"], g["E/maxit/chi"])
    assert r["iterations"] == int(g["E/maxit/iterations"])


def test_bicubic_matrix_is_exact_inverse_of_hermite_constraints():
    """Closed-form pin: interpolation_class.cpp:539-558 times the constraint matrix
    of the commented-out construction (:415-536) is the identity."""
    M = oracle.bicubic_matrix().astype(np.float64)
    Cm = np.zeros((16, 16))
    nodes = [(1, 1), (2, 1), (1, 2), (2, 2)]
    for n, (x, y) in enumerate(nodes):
        for j in range(4):
            for i in range(4):
                Cm[n, j * 4 + i] = y**j * x**i
                Cm[4 + n, j * 4 + i] = i * y**j * x**(i - 1) if i else 0
                Cm[8 + n, j * 4 + i] = j * y**(j - 1) * x**i if j else 0
                Cm[12 + n, j * 4 + i] = i * j * y**(j - 1) * x**(i - 1) if i and j else 0
    assert np.array_equal(M @ Cm, np.eye(16))


def test_lm_constants_and_solver_step():
    """The damped solve: (A/N with diag*(1+lambda)) dp = b/N, checked against fp64 numpy."""
    rng = np.random.default_rng(0)
    J = rng.normal(size=(400, 6)) * np.array([1, 1, 30, 30, 30, 30])
    r = rng.normal(size=400)
    A = (J.T @ J).astype(np.float32)
    b = (J.T @ r).astype(np.float32)
    e = oracle.OracleEngine(n_threads=1)
    lam, scaling = 1e-4, 1.0 / 400
    dp = e.solve_step(np.triu(A), b, lam, scaling)
    Ad = A.astype(np.float64) * scaling
    Ad[np.diag_indices(6)] *= 1 + lam
    want = np.linalg.solve(Ad, b.astype(np.float64) * scaling)
    assert np.allclose(dp, want, rtol=2e-4, atol=1e-7)


def test_double_accumulators_only_move_chi_slightly(golden):
    g = golden
    x0, y0, x1, y1 = g["A/rect"]
    xy = oracle.rect_points(x0, y0, x1, y1)
    e = _engine(g, n_threads=20, pyramid=(0, 1, 2), accum_double=True)
    r = e.correlate(np.zeros(6), xy, center=(95.0, 95.0))
    assert abs(r["chi"] - g["A/T20/chi"]) < 1e-4 * g["A/T20/chi"]
    assert np.abs(r["params"] - g["A/T20/params"]).max() < 2e-5


def test_threaded_run_is_bit_identical_to_sequential_chunks(golden):
    g = golden
    x0, y0, x1, y1 = g["A/rect"]
    xy = oracle.rect_points(x0, y0, x1, y1)
    e = _engine(g, n_threads=20, pyramid=(0, 1, 2), real_threads=True)
    _check_case(g, "A/T20", e.correlate(np.zeros(6), xy, center=(95.0, 95.0)))


def test_quadratic_extension_recovers_truth():
    """12-parameter model: our extension, parity unpinned -- checked against ground truth."""
    truth = np.array([1.2, -0.8, .003, -.002, .001, .004, 2e-5, -1e-5, 1.5e-5, -2e-5, 1e-5, 5e-6])
    und, dfm = synth.make_pair(256, 256, 21, truth, center=(128, 128))
    e = oracle.OracleEngine(model=oracle.FM_QUAD, n_threads=4, pyramid=(0, 1, 1), accum_double=True)
    e.set_image("und", und)
    e.set_image("def", dfm)
    r = e.correlate(np.zeros(12), oracle.rect_points(40, 40, 216, 216), center=(128.0, 128.0))
    assert r["error_code"] == 0
    assert np.abs(r["params"][:2] - truth[:2]).max() < 0.02
    assert np.abs(r["params"][2:6] - truth[2:6]).max() < 5e-4
    assert np.abs(r["params"][6:] - truth[6:]).max() < 1.5e-5


# ---- against the compiled reference itself (this container; the .so travels to the GPU box)

@pytest.mark.ref
@pytest.mark.parametrize("threads", [1, 7, 20])
def test_oracle_equals_compiled_reference_bitwise(threads):
    truth = (0.9, 1.3, -.002, .004, .003, -.001)
    und, dfm = synth.make_pair(300, 280, 5, truth, center=(140, 150))
    xy = oracle.rect_points(50, 60, 229, 240)
    R = oracle.RefEngine(n_threads=threads, pyramid=(0, 1, 2))
    O = oracle.OracleEngine(n_threads=threads, pyramid=(0, 1, 2))
    for E in (R, O):
        E.set_image("und", und)
        E.set_image("def", dfm)
    rr = R.correlate(np.zeros(6), xy, center=(139.0, 150.0))
    ro = O.correlate(np.zeros(6), xy, center=(139.0, 150.0))
    assert_bit_equal(rr["params"], ro["params"])
    assert_bit_equal(rr["chi"], ro["chi"])
    assert rr["iterations"] == ro["iterations"] and rr["error_code"] == ro["error_code"]
    for lv in (0, 1, 2):
        assert np.array_equal(R.pyramid_level(1, lv), O.pyramid_level(1, lv))
        assert_bit_equal(R.level_points(lv), O.level_points(lv))
        A1, b1, c1, _ = R.evaluate(lv, rr["params"])
        A2, b2, c2, _ = O.evaluate(lv, ro["params"])
        assert_bit_equal(A1, A2)
        assert_bit_equal(b1, b2)
        assert_bit_equal(c1, c2)


@pytest.mark.ref
def test_oracle_frame_rotation_equals_compiled_reference():
    """und <- def <- nxt pointer rotation (pyramid_class.cpp:211-258) over three frames."""
    frames = [synth.make_image(160, 160, 9, p, (80, 80)) for p in
              (None, (0.5, 0.2, 0, 0, 0, 0), (1.0, 0.4, 0.001, 0, 0, 0.001))]
    xy = oracle.rect_points(40, 40, 120, 120)
    R = oracle.RefEngine(n_threads=3, pyramid=(0, 1, 1))
    O = oracle.OracleEngine(n_threads=3, pyramid=(0, 1, 1))
    for E in (R, O):
        E.set_image("und", frames[0])
        E.set_image("def", frames[1])
        E.set_image("nxt", frames[2])
    a, b = R.correlate(np.zeros(6), xy, center=(80., 80.)), O.correlate(np.zeros(6), xy, center=(80., 80.))
    assert_bit_equal(a["params"], b["params"])
    for E in (R, O):
        E.und_from_def()
        E.def_from_nxt()
    a, b = R.correlate(np.zeros(6), xy, center=(80., 80.)), O.correlate(np.zeros(6), xy, center=(80., 80.))
    assert_bit_equal(a["params"], b["params"])
    assert_bit_equal(a["chi"], b["chi"])


@pytest.mark.ref
def test_blob_points_equal_compiled_reference():
    for seed in (1, 2, 3):
        contour = synth.star_polygon(200.0, 180.0, 120.0, n_vertices=40, seed=seed)
        ref = oracle.RefEngine().blob_points(contour)
        assert_bit_equal(oracle.blob_points(contour), ref)


# ---------------------------------------------------------------- three-channel colour images

@pytest.fixture(scope="module")
def golden_color():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_color_v1.npz"))


@pytest.mark.parametrize("iname,interp", [("nearest", oracle.IM_NEAREST), ("bilinear", oracle.IM_BILINEAR), ("bicubic", oracle.IM_BICUBIC)])
def test_color_images_match_reference_golden(golden_color, iname, interp):
    """number_of_colors = 3: per-channel pyramid and the per-colour loop of interpolation_class.cpp:712-750 with the
    column indexing the reference's coefficient builders execute (tests/golden/make_golden_color.py)."""
    g = golden_color
    o = oracle.OracleEngine(n_threads=3, interp=interp, pyramid=tuple(int(v) for v in g["pyramid"]), colors=3)
    o.set_image("und", g["und"])
    o.set_image("def", g["def"])
    r = o.correlate(np.zeros(6, np.float32), oracle.rect_points(*(int(v) for v in g["rect"])), center=tuple(float(v) for v in g["center"]))
    assert_bit_equal(r["params"], g[f"{iname}/params"], iname)
    assert np.float32(r["chi"]) == g[f"{iname}/chi"] and r["iterations"] == int(g[f"{iname}/iterations"])
    assert r["error_code"] == int(g[f"{iname}/error_code"]) and r["number_of_points"] == int(g[f"{iname}/number_of_points"])
    if iname == "bicubic":
        assert np.array_equal(o.pyramid_level(0, 1), g["pyr_und1"]) and np.array_equal(o.pyramid_level(1, 1), g["pyr_def1"])
        for lv in (1, 0):
            p = np.array([0.6, -0.35, 0.002, 0, 0, 0.003], np.float32)
            p[:2] *= np.float32(1.0 / (1 << lv))
            A, b, chi, err = o.evaluate(lv, p)
            assert_bit_equal(np.triu(A), np.triu(g[f"bicubic/eval{lv}/A"]))
            assert_bit_equal(b, g[f"bicubic/eval{lv}/b"])
            assert np.float32(chi) == g[f"bicubic/eval{lv}/chi"]
