"""GPU suite: the headless C++ host (frame loop, constant-velocity guess, domain updates, CSV report)
against the same sequence driven on the CPU oracle by a restatement of the manager's bookkeeping
(manager_class.cpp:1297-1541, 2602-2707) kept in this test file."""
import numpy as np
import pytest

import oracle
from correlation_b200 import host, synth

pytestmark = pytest.mark.gpu


def make_frames(n, rows, cols, seed, rate, center):
    """frame k = field warped by k * rate (constant velocity)."""
    return [synth.make_image(rows, cols, seed, None if k == 0 else tuple(k * np.array(rate)), center)
            for k in range(n)]


def oracle_sequence(frames, xy0, center, model=oracle.FM_AFFINE, pyramid=(0, 1, 2), reference_first=True,
                    lagrangian=False, rect_center=None):
    """The manager's CPU path for ONE sector whose centre equals the global centre (no sector offset)."""
    n = oracle.N_PARAMS[model]
    O = oracle.OracleEngine(model=model, n_threads=20, pyramid=pyramid, accum_double=True)
    O.set_image("und", frames[0])
    O.set_image("def", frames[1])
    xy = np.array(xy0, np.float32)
    p = np.zeros(n, np.float32)
    p_prev = np.zeros(n, np.float32)
    out = []
    c = rect_center
    for k in range(len(frames) - 1):
        if k > 0:
            if not reference_first:
                O.und_from_def()
            O.set_image("nxt", frames[k + 1])
            O.def_from_nxt()
        if k == 0:
            guess = np.zeros(n, np.float32)
            p_prev = guess.copy()
        else:
            if (not lagrangian) and reference_first:
                guess = (p + (p - p_prev)).astype(np.float32)  # manager_class.cpp:2677-2686
            else:
                guess = p.copy()
            p_prev = p.copy()
            if lagrangian:  # add_pair, manager_class.cpp:37-47, offset = und_center - past_und_center = (u, v)
                xy = np.stack([np.floor(np.float32(p[0]) + xy[:, 0] + np.float32(0.5)),
                               np.floor(np.float32(p[1]) + xy[:, 1] + np.float32(0.5))], 1).astype(np.float32)
                if c is not None:
                    c = (float(int(c[0] + p[0] + 0.5)), float(int(c[1] + p[1] + 0.5)))
        r = O.correlate(guess, xy, center=c)
        p = r["params"].copy()
        out.append(dict(r, guess=guess))
    return out


def test_blob_sequence_constant_velocity_vs_oracle():
    rate = (0.8, -0.5, 0.0004, -0.0005, 0.0005, 0.0004)
    frames = make_frames(6, 300, 320, 3, rate, (160, 150))
    contour = synth.star_polygon(160.0, 150.0, 95.0, n_vertices=24, seed=3)
    got = host.run_sequence(frames, contour=contour, pyramid=(0, 1, 2))
    assert got["error"] == 0
    hdr, rows = host.parse_report(got["csv"])
    assert hdr[:3] == ["Frame#", "und_file_string", "def_file_string"] and hdr[-5:] == ["chi", "number_of_points", "iterations", "error_status", "error_code"]
    assert len(rows) == 5
    want = oracle_sequence(frames, oracle.blob_points(contour), None)
    for k, (row, w) in enumerate(zip(rows, want)):
        for p in range(6):
            tol = 2e-4 if p < 2 else 2e-6  # CSV carries 6 significant digits
            assert abs(row[f"parameter_{p}"] - w["params"][p]) < tol + 1e-5 * abs(w["params"][p]), (k, p)
            assert abs(row[f"Initial_guess_{p}"] - w["guess"][p]) < tol + 1e-5 * abs(w["guess"][p]), (k, p)
        assert int(row["number_of_points"]) == w["number_of_points"]
        assert abs(int(row["iterations"]) - w["iterations"]) <= 1
        assert int(row["error_code"]) == 0
    # exact floats of the last frame
    last = got["rows"][0]
    d = np.abs(last["params"][:6] - want[-1]["params"])
    assert d[:2].max() < 1e-4 and d[2:].max() < 1e-6
    assert abs(last["chi"] - want[-1]["chi"]) < 1e-4 * want[-1]["chi"]
    # constant velocity: from frame 2 on the extrapolated guess is already within a few 1e-3 px
    assert abs(rows[3]["Initial_guess_0"] - rows[3]["parameter_0"]) < 2e-2


def test_rect_lagrangian_previous_image_vs_oracle():
    rate = (1.3, 0.7, 0.0, 0.0, 0.0, 0.0)
    frames = make_frames(4, 256, 256, 9, rate, (128, 128))
    rect = (64, 64, 193, 193)  # 129 wide: xdim = 64, centre 128
    got = host.run_sequence(frames, rect=rect, pyramid=(0, 1, 1), deformation=1, reference=1)
    assert got["error"] == 0
    xy = oracle.rect_points(64, 64, 192, 192)
    want = oracle_sequence(frames, xy, None, pyramid=(0, 1, 1), reference_first=False, lagrangian=True,
                           rect_center=(128.0, 128.0))
    hdr, rows = host.parse_report(got["csv"])
    assert len(rows) == 3
    for k, (row, w) in enumerate(zip(rows, want)):
        assert abs(row["parameter_0"] - w["params"][0]) < 3e-4 and abs(row["parameter_1"] - w["params"][1]) < 3e-4, k
        assert int(row["number_of_points"]) == w["number_of_points"]
    # the domain followed the material: centre moved by the rounded displacement each frame
    assert rows[2]["und_center_x"] == pytest.approx(128 + 2 * round(1.3), abs=1.01)


def test_rect_subdivisions_batch_equals_serial():
    truth = (1.1, 0.6, 0.002, -0.001, 0.001, 0.002)
    frames = [synth.make_image(384, 384, 41, None, (192, 192)), synth.make_image(384, 384, 41, truth, (192, 192))]
    rect = (32, 32, 352, 352)
    a = host.run_sequence(frames, rect=rect, subdivisions=(4, 4), pyramid=(0, 1, 2))
    b = host.run_sequence(frames, rect=rect, subdivisions=(4, 4), pyramid=(0, 1, 2), batch=True)
    assert a["error"] == 0 and b["error"] == 0
    # grid-wide launch per sector vs one CTA per sector: same arithmetic per pixel, different
    # summation tree (fp64 atomics across CTAs vs one CTA), hence a slightly different LM path
    d = np.abs(a["rows"]["params"] - b["rows"]["params"])
    assert d[:, :2].max() < 5e-5 and d[:, 2:].max() < 1e-6
    # sector k sees the global displacement plus the gradient times its centre offset
    for r in a["rows"]:
        dx, dy = r["und_center_x"] - 192.0, r["und_center_y"] - 192.0
        assert abs(r["params"][0] - (truth[0] + truth[2] * dx + truth[3] * dy)) < 0.02
        assert abs(r["params"][1] - (truth[1] + truth[4] * dx + truth[5] * dy)) < 0.02


def warp_f32(xy, c, p):
    """model_class.cpp:171-172 in fp32, left to right (what getDefXY0 returns for the affine model)."""
    x, y = xy[:, 0].astype(np.float32), xy[:, 1].astype(np.float32)
    p = np.asarray(p, np.float32)
    dx, dy = x - np.float32(c[0]), y - np.float32(c[1])
    return np.stack([x + p[0] + p[2] * dx + p[3] * dy, y + p[1] + p[4] * dx + p[5] * dy], 1).astype(np.float32)


def test_rect_strict_lagrangian_previous_image_vs_oracle():
    """def_strict_Lagrangian (manager_class.cpp:359-373): the undeformed points of frame k + 1 ARE the deformed
    points of frame k (non-integer: the generic pixel-list kernel runs), reference image = previous frame."""
    rate = (1.3, 0.7, 0.001, 0.0, 0.0, -0.001)
    frames = make_frames(4, 256, 256, 9, rate, (128, 128))
    rect = (64, 64, 193, 193)
    got = host.run_sequence(frames, rect=rect, pyramid=(0, 1, 1), deformation=0, reference=1)
    assert got["error"] == 0
    O = oracle.OracleEngine(n_threads=20, pyramid=(0, 1, 1), accum_double=True)
    O.set_image("und", frames[0])
    O.set_image("def", frames[1])
    xy = oracle.rect_points(64, 64, 192, 192)
    c = (128.0, 128.0)
    p = np.zeros(6, np.float32)
    want = []
    for k in range(3):
        if k > 0:
            O.und_from_def()
            O.set_image("nxt", frames[k + 1])
            O.def_from_nxt()
            xy = warp_f32(xy, c, p)                                    # und points := last deformed points
            c = (float(int(c[0] + p[0] + 0.5)), float(int(c[1] + p[1] + 0.5)))  # rounded deformed centre, :2085-2086
        r = O.correlate(p.copy() if k else np.zeros(6, np.float32), xy, center=c)
        p = r["params"].copy()
        want.append(r)
    hdr, rows = host.parse_report(got["csv"])
    assert len(rows) == 3
    for k, (row, w) in enumerate(zip(rows, want)):
        for q in range(6):
            tol = 3e-4 if q < 2 else 3e-6
            assert abs(row[f"parameter_{q}"] - w["params"][q]) < tol + 1e-5 * abs(w["params"][q]), (k, q, row, w["params"])
        assert int(row["number_of_points"]) == w["number_of_points"] == 129 * 129
        assert int(row["error_code"]) == 0
    last = got["rows"][0]
    d = np.abs(last["params"][:6] - want[-1]["params"])
    assert d[:2].max() < 1e-4 and d[2:].max() < 1e-6


def test_annular_2x3_sectors_sequence_and_global_results():
    """perform_single_frame_correlation_annular (manager_class.cpp:557-814) over 2 radial x 3 angular sectors and
    update_global_results (:2709-2753): pixel-weighted mean of the sectors' deformed centres and angles."""
    rate = (0.9, -0.6, 0.0, -0.004, 0.004, 0.0)  # translation + 0.004 rad rotation per frame
    frames = make_frames(3, 360, 360, 17, rate, (180, 180))
    ann = (180.0, 180.0, 40.0, 150.0)  # x_center y_center r_inside r_outside
    got = host.run_sequence(frames, annulus=ann, subdivisions=(2, 3), pyramid=(0, 1, 1), on_error=host.ERROR_CONTINUE)
    assert got["error"] == 0
    rows = got["rows"]
    assert len(rows) == 6
    O = oracle.OracleEngine(n_threads=20, pyramid=(0, 1, 1), accum_double=True)
    PI = np.float32(3.14159265359)
    dr, da = np.float32((150.0 - 40.0) / 2), np.float32(2) * PI / np.float32(3)
    p = np.zeros((6, 6), np.float32)
    p_prev = np.zeros((6, 6), np.float32)
    for k in range(2):
        O.set_image("und", frames[0])
        O.set_image("def", frames[k + 1])
        res = []
        for i in range(2):
            for j in range(3):
                s = i * 3 + j
                xy = oracle.annulus_points(np.float32(40.0) + i * dr, dr, np.float32(j) * da, da, 180.0, 180.0, 3)
                guess = np.zeros(6, np.float32) if k == 0 else (p[s] + (p[s] - p_prev[s])).astype(np.float32)
                res.append(O.correlate(guess, xy))
        p_prev = p.copy() if k else np.zeros((6, 6), np.float32)
        p = np.array([r["params"] for r in res], np.float32)
    n = np.array([r["number_of_points"] for r in res], np.float64)
    ang = np.array([np.arctan2(float(q[4] - q[3]), float(q[2] + q[5] + 2.0)) for q in p])
    cx = np.array([float(r["und_center"][0]) + float(q[0]) for r, q in zip(res, p)])
    cy = np.array([float(r["und_center"][1]) + float(q[1]) for r, q in zip(res, p)])
    for s in range(6):
        assert int(rows[s]["number_of_points"]) == int(n[s])
        d = np.abs(rows[s]["params"][:6] - p[s])
        assert d[:2].max() < 1e-4 and d[2:].max() < 1e-6, (s, rows[s]["params"][:6], p[s])
        assert abs(rows[s]["def_angle"] - ang[s]) < 2e-6
        assert abs(rows[s]["def_global_angle"] - (ang * n).sum() / n.sum()) < 2e-6
        assert abs(rows[s]["def_global_center_x"] - (cx * n).sum() / n.sum()) < 2e-3
        assert abs(rows[s]["def_global_center_y"] - (cy * n).sum() / n.sum()) < 2e-3
    assert abs(rows[0]["def_global_angle"] - 2 * 0.004) < 2e-4  # and it is the imposed rotation after two frames


@pytest.mark.parametrize("mode,frames_done,sector3_done", [(host.ERROR_STOP_ALL, 1, False), (host.ERROR_STOP_FRAME, 2, False),
                                                           (host.ERROR_CONTINUE, 2, True)])
def test_error_handling_modes(mode, frames_done, sector3_done):
    """errorHandlingModeEnum (manager_class.cpp:535-546, :1493): sectors 2 and 3 (right column) leave the image under
    a 25 px displacement -> error_interpolation_out_of_image. stopAll ends the frame at the first failing sector and
    the run after that frame; stopFrame ends only the frame; continue processes everything."""
    shift = (25.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    frames = make_frames(2, 256, 256, 5, shift, (128, 128))
    frames.append(frames[1])  # the same displacement again: the extrapolated guess (p + (p - p_prev) = 25) stays valid
    got = host.run_sequence(frames, rect=(40, 40, 236, 236), subdivisions=(2, 2), pyramid=(0, 1, 1), guess=(25.0, 0.0),
                            on_error=mode)
    assert got["error"] == 1  # managerClass::error = status of the last sector processed
    hdr, rows = host.parse_report(got["csv"])
    assert len(rows) == 4 * frames_done
    last = got["rows"]
    assert int(last[0]["error_code"]) == 0 and int(last[1]["error_code"]) == 0
    assert abs(last[0]["params"][0] - 25.0) < 0.05
    assert int(last[2]["error_code"]) == 2
    assert (int(last[3]["number_of_points"]) > 0) == sector3_done
    assert int(last[0]["frame"]) == frames_done - 1


def test_full_size_c3_sequence_vs_oracle_and_truth():
    """BASELINE config 3 at full size through the C++ host loop (bench.run_c3): 100 frames of 2048^2, 64-vertex star
    blob (1.54 M pixels), Eulerian + first-image reference, constant-velocity guesses. The first frame pairs against
    the oracle run of the same sequence (fp64 accumulators; displacement / gradients / iterations at the BASELINE
    tolerances, chi up to the reference's own spread over its two builds), the last pair against the truth of the
    synthetic sequence, and no frame may report an error."""
    import argparse
    import bench
    w = bench.workload("c3")
    args = argparse.Namespace(steps=1, warmup=0, mode="parity")
    line = bench.run_c3(args, w, with_cpu=True)
    assert line["config"]["errors"] == 0 and line["config"]["frame_pairs"] == 99
    blk = line["parity"]["vs_oracle_first_frames"]
    assert blk["units_on_oracle_lm_path"] == blk["units_compared"] == 4
    assert blk["max_abs_duv"] < 1e-4 and blk["max_abs_dgrad"] < 1e-6 and blk["max_abs_diterations"] <= 1, blk
    assert line["parity"]["chi_within_reference_self_spread"], line["parity"]
    got, truth = np.array(line["config"]["last_frame_params"]), np.array(line["config"]["last_frame_truth"])
    # the parameters are about the reference's sequential-fp32 centre of the blob list, about a pixel off the geometric
    # centre the synthetic field is defined about: u, v of a rotating field differ by (gradient x offset)
    assert np.abs(got[:2] - truth[:2]).max() < 1.5 and np.abs(got[2:] - truth[2:]).max() < 2e-5, (got, truth)
    assert line["value"] > 500.0  # frames / s end to end; the CPU oracle does ~10
