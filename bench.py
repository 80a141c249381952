#!/usr/bin/env python
"""bench.py -- domain pixel*GN-evaluations / s of the DIC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c1|c5|c3] [--impl ours|reference]

Default workload = BASELINE config 4 for EVERY N: the 8192^2 pair whose domain is subdivided into 64 x 64 = 4096
subsets of 125^2 (affine model, pyramid 0/1/2), the subsets sharded over the N GPUs in whole rows of subsets
(strong scaling, no data-path collective; manager_class.cpp:304-547 is the serial loop this replaces). It is the
workload the north star's 1/2/4/8-GPU target is quoted on, it fits one GPU, and its CPU arm is the UNMODIFIED
reference engine (oracle/_ref). A *step* is one pass of the hot path over the whole domain from the zero initial
guess: every pyramid level, every LM iteration, reduction and solve of every subset
(correlation_class.cpp:349-640 semantics). Work per step = sum over subsets and levels of N_level x
evaluations_level, reported by the engine itself.

  value     pixel*evaluations / s, device-resident inputs, all ranks (CUDA events, max over ranks)
  e2e       same metric through the C-ABI with HOST (pinned) images: per step H2D of this rank's band of the
            image pair, both pyramid builds, correlate, D2H of the result records
  parity    results against the fp64-accumulator oracle on a stratified sample of subsets and, at N > 1, the
            gathered records of all ranks bit-compared with rank 0's own single-GPU run of all 4096 subsets
  roofline  the fused GN kernel against the measured HBM copy bandwidth, with SURVEY 8d's 10 algorithmic bytes
            per pixel*evaluation
  other_workloads   short runs of config 2 (one 9 M-pixel annulus, 12 parameters; N = 1) and config 5 (one 236 M
            pixel domain row-split over the ranks with the in-kernel NVLink all-reduce; N > 1), each with its own
            parity block
  cpu_baseline / --impl reference
            the reference CPU algorithm on this box's host cores: oracle/_ref (unmodified reference, 6-parameter
            workloads) or the oracle port (12-parameter workload, which the reference does not have).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_PIXEL_EVAL = 10.0  # 8 B coordinates + 1 B und + 1 B def
TOLERANCES = {"duv": 1e-4, "dgrad": 1e-6, "rel_dchi": 1e-5, "iterations": 1}  # BASELINE.json north_star


# ------------------------------------------------------------------------------ workloads

def workload(name):
    two_pi = 2.0 * np.pi
    if name == "c1":  # BASELINE config 1 (the reference's CPU-runnable case)
        return dict(key="c1", name="c1: 1024^2 pair, 511x511 rect, affine 6-param, bicubic, pyramid 0/1/2",
                    rows=1024, cols=1024, seed=1, model="affine", pyramid=(0, 1, 2),
                    truth=(1.75, -0.6, .004, -.003, .002, .005), center=(512.0, 512.0),
                    domain=("rect", 257, 257, 767, 767))
    if name == "c2":  # BASELINE config 2
        return dict(key="c2", name="c2: 4096^2 pair, annulus ri=600 ro=1800, quadratic 12-param, bicubic, pyramid 0..3",
                    rows=4096, cols=4096, seed=2, model="quad", pyramid=(0, 1, 3),
                    truth=(2.5, -1.75, .002, -.0015, .001, .0025, 1e-6, -5e-7, 8e-7, -1e-6, 6e-7, 4e-7),
                    center=(2048.0, 2048.0), domain=("annulus", 600.0, 1200.0, 0.0, two_pi, 2048.0, 2048.0, 1))
    if name == "c4":  # BASELINE config 4: 4096 subsets of 125^2 -- the default for every N
        return dict(key="c4", name="c4: 8192^2 pair, 64x64 subsets of 125^2, affine, pyramid 0/1/2",
                    rows=8192, cols=8192, seed=4, model="affine", pyramid=(0, 1, 2),
                    truth=(1.25, -0.75, .0004, -.0003, .0002, .0005), center=(4096.0, 4096.0),
                    domain=("subsets", 64, 8128, 64))
    if name == "c3":  # BASELINE config 3: frames/s of a 100-frame sequence
        return dict(key="c3", name="c3: 100-frame 2048^2 sequence, 64-vertex star blob (mean radius 700), affine, pyramid 0/1/2, "
                                   "Eulerian + first-image reference, constant-velocity initial guess",
                    rows=2048, cols=2048, seed=3, model="affine", pyramid=(0, 1, 2), frames=100,
                    rate=(0.8, -0.5, 0.0, -0.0005, 0.0005, 0.0), center=(1024.0, 1024.0),
                    truth=(0.8, -0.5, 0.0, -0.0005, 0.0005, 0.0), domain=("blob", 1024.0, 1024.0, 700.0, 64))
    if name == "c5":  # BASELINE config 5: one huge domain, row-split across GPUs
        return dict(key="c5", name="c5: 16384^2 pair, one 15361^2 rect domain, affine, pyramid 0..4, row-split + in-kernel all-reduce",
                    rows=16384, cols=16384, seed=5, model="affine", pyramid=(0, 1, 4),
                    truth=(25.0, -15.0, .004, -.003, .002, .005), center=(8192.0, 8192.0),
                    domain=("rowsplit", 512, 512, 15872, 15872), spectrum=(5.0, 600.0), n_waves=64)
    raise SystemExit(f"unknown workload {name}")


def make_images(w, device):
    """(und, def) as torch uint8 CUDA tensors; synthetic analytic speckle, SURVEY 8d."""
    from correlation_b200 import synth
    kw = dict(spectrum=w.get("spectrum"), n_waves=w.get("n_waves", synth.N_WAVES))
    und = synth.make_image(w["rows"], w["cols"], w["seed"], None, w["center"], device=device, **kw)
    dfm = synth.make_image(w["rows"], w["cols"], w["seed"], w["truth"], w["center"], device=device, **kw)
    return und, dfm


def subset_boxes(lo, hi, n):
    """n x n sectors of a rectangle, manager_class.cpp:274-310 arithmetic."""
    fdim = (abs(hi - lo) / n - 1.0) / 2.0
    dim = (abs(hi - lo) // n - 1) // 2
    boxes = []
    for i in range(n):
        cx = int(0.5 + lo + fdim + (2.0 * fdim + 1.0) * i)
        for j in range(n):
            cy = int(0.5 + lo + fdim + (2.0 * fdim + 1.0) * j)
            boxes.append((cx - dim, cy - dim, cx + dim, cy + dim))
    return boxes


def upload_band(y_first, y_last, rows, pyramid_stop):
    """Image rows [begin, end) a rank must hold to work on domain rows y_first..y_last: the domain plus a halo for
    the workloads' displacement (<= 64 px), the 2-pixel bicubic support and 2^(level + 2) rows of pyramid support
    at every cut (dic_stage_next_pair_rows)."""
    halo = 64 + (8 << pyramid_stop)
    return max(0, y_first - halo), min(rows, y_last + 1 + halo)


def stratified_sample(n_units, n_sample):
    """n_sample unit ids spread evenly over 0..n_units-1 (first and last included)."""
    if n_sample >= n_units:
        return list(range(n_units))
    return sorted({int(round(k * (n_units - 1) / (n_sample - 1))) for k in range(n_sample)})


SELF_SPREAD_NOTE = ("within_tolerance applies the north-star numbers literally against the oracle with fp64 accumulators; "
                    "chi_within_reference_self_spread asks instead whether the GPU's chi is at most 1.5 x as far from that oracle as the "
                    "reference's own fp32 arithmetic is on the same input in either of two builds (sums cut into 20 chunks, its default "
                    "NUMBER_OF_THREADS, or into one): chi of a large domain is not defined to 1e-5 by the reference (DESIGN.md section 5)")


def reference_spread(run_ref, want, gpu_block, what):
    """The reference's OWN arithmetic against the fp64-accumulator oracle results `want` (list of result dicts):
    run_ref(n_chunks) -> list of result dicts of the restatement with fp32 accumulators cut into n_chunks pieces, i.e. the
    reference built with NUMBER_OF_THREADS = n_chunks (defines.hpp:10; 20 is its default, 1 a single-threaded build).
    Returns the keys to merge into a parity record."""
    cols = lambda rs: ([r["params"] for r in rs], [r["chi"] for r in rs], [r["iterations"] for r in rs])
    evs = lambda rs: [r["evaluations"][:8] for r in rs]
    out, worst = {}, 0.0
    for chunks, key in ((20, "reference_self_spread"), (1, "reference_self_spread_one_chunk")):
        try:
            rs = run_ref(chunks)
            sp = parity_block(*cols(rs), *cols(want), gpu_evals=evs(rs), want_evals=evs(want))
            sp.pop("within_tolerance", None)
            sp["what"] = (f"oracle with the reference's fp32 accumulators in {chunks} chunk(s) (NUMBER_OF_THREADS = {chunks}) vs the "
                          f"oracle with fp64 accumulators, {what}")
            out[key] = sp
            worst = max(worst, sp["max_rel_dchi"])
        except Exception as ex:
            out[key] = {"error": repr(ex)}
    out["chi_within_reference_self_spread"] = bool(gpu_block["max_rel_dchi"] <= max(TOLERANCES["rel_dchi"], 1.5 * worst))
    out["note"] = SELF_SPREAD_NOTE
    return out


def parity_block(gpu_params, gpu_chi, gpu_iters, want_params, want_chi, want_iters, gpu_evals=None, want_evals=None, n_grad_to=6):
    """Deviations of GPU results from oracle results (arrays over the compared units) and the verdict.

    With the evaluations-per-level of both sides, units are split into those ON the oracle's LM path (same number of
    evaluations at every level) and those off it: a convergence test |d chi| < precision that falls on the other side
    costs or saves one evaluation, and the two stopping points are then a convergence threshold (~1e-3 px) apart by
    construction -- the BASELINE tolerances (which allow iterations +-1) are applied to the on-path units, the off-path
    units are counted and bounded separately. `within_tolerance_literal` applies them to every unit regardless."""
    gp, wp = np.atleast_2d(np.asarray(gpu_params, np.float64)), np.atleast_2d(np.asarray(want_params, np.float64))
    gc, wc = np.atleast_1d(np.asarray(gpu_chi, np.float64)), np.atleast_1d(np.asarray(want_chi, np.float64))
    gi, wi = np.atleast_1d(np.asarray(gpu_iters)), np.atleast_1d(np.asarray(want_iters))
    n = gp.shape[0]
    on = np.ones(n, bool)
    if gpu_evals is not None:
        ge, we = np.atleast_2d(np.asarray(gpu_evals)), np.atleast_2d(np.asarray(want_evals))
        on = (ge == we).all(1)
    duv_u = np.abs(gp[:, :2] - wp[:, :2]).max(1)
    dgrad_u = np.abs(gp[:, 2:n_grad_to] - wp[:, 2:n_grad_to]).max(1) if gp.shape[1] > 2 else np.zeros(n)
    rel = np.abs(gc - wc) / np.maximum(np.abs(wc), 1e-30)
    dit = np.abs(gi.astype(np.int64) - wi.astype(np.int64))
    mx = lambda a, m: float(a[m].max()) if m.any() else 0.0
    ok_on = mx(duv_u, on) < TOLERANCES["duv"] and mx(dgrad_u, on) < TOLERANCES["dgrad"] and mx(rel, on) <= TOLERANCES["rel_dchi"]
    ok_lit = mx(duv_u, on | ~on) < TOLERANCES["duv"] and float(dgrad_u.max()) < TOLERANCES["dgrad"] and float(rel.max()) <= TOLERANCES["rel_dchi"]
    off = ~on
    blk = {"units_compared": int(n), "units_on_oracle_lm_path": int(on.sum()),
           "max_abs_duv": mx(duv_u, on), "max_abs_dgrad": mx(dgrad_u, on), "max_rel_dchi": mx(rel, on),
           "median_rel_dchi": float(np.median(rel[on])) if on.any() else 0.0,
           "units_with_rel_dchi_above_1e-5": int((rel[on] > 1e-5).sum()), "max_abs_diterations": int(dit.max()),
           "tolerances": TOLERANCES,
           "within_tolerance": bool(ok_on and int(dit.max()) <= 1 and off.sum() <= max(1, 0.05 * n) and mx(duv_u, off) < 2e-3),
           "within_tolerance_literal": bool(ok_lit and int(dit.max()) <= 1)}
    if off.any():
        blk["off_path"] = {"units": int(off.sum()), "max_abs_duv": mx(duv_u, off), "max_abs_dgrad": mx(dgrad_u, off),
                           "max_rel_dchi": mx(rel, off)}
    return blk


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region on a side thread.
    NVML in-process (what nvidia-smi itself reads; ~10 us per sample, no fork inside the timed region);
    falls back to the nvidia-smi command line when the binding is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, world=1):
        # every rank samples its own GPU; the period grows with the number of ranks so that the box-wide
        # rate of driver queries stays the same (NVML calls serialise on a driver lock shared by all GPUs)
        self.index, self.samples, self.stop = index, [], False
        self.period = 0.005 * max(1, world)
        self.t = threading.Thread(target=self._run, daemon=True)
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(get(self.handle))
            return [str(sm), str(mx)] + ["Active" if mask & bit else "Not Active" for _, bit in self.BITS]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        return [x.strip() for x in out.strip().split(",")]

    def _run(self):
        while not self.stop:
            try:
                f = self._sample()
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(self.period if self.nvml is not None else max(0.05, self.period))

    def poke(self):
        """One sample taken by the caller, between two timed steps (outside the CUDA events that time a step): the side
        thread alone gets one or two samples into a 25 ms region."""
        try:
            f = self._sample()
            if len(f) >= 6:
                self.samples.append(f)
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for k, (n, _) in enumerate(self.BITS) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------ CPU reference arm

def cpu_run(w, und, dfm, threads):
    """One 'step' of the reference CPU algorithm: both pyramid builds + Newton_Raphson (cold cache).
    Returns (pixel_evaluations, seconds, kind, result, sample, pyramid_seconds)."""
    import oracle
    use_ref = oracle.have_ref() and w["model"] == "affine" and w["rows"] <= 8192  # reference cache overflows int above ~11k^2
    pyr = w["pyramid"]
    if use_ref:
        eng = oracle.RefEngine(model=oracle.FM_AFFINE, n_threads=threads, pyramid=pyr)
        counter = oracle.OracleEngine(model=oracle.FM_AFFINE, n_threads=threads, pyramid=pyr, real_threads=True)
    else:
        model = oracle.FM_QUAD if w["model"] == "quad" else oracle.FM_AFFINE
        eng = oracle.OracleEngine(model=model, n_threads=threads, pyramid=pyr, real_threads=True)
        counter = None
    d = w["domain"]
    sample_note = None
    if d[0] == "rect":
        xy, center = oracle.rect_points(*d[1:]), ((d[1] + d[3]) / 2.0, (d[2] + d[4]) / 2.0)
    elif d[0] == "rowsplit":
        # bounded sample: the central 1/16 of the domain (same images, same pyramid, same centre)
        cx, cy = (d[1] + d[3]) // 2, (d[2] + d[4]) // 2
        hw, hh = (d[3] - d[1]) // 8, (d[4] - d[2]) // 8
        xy, center = oracle.rect_points(cx - hw, cy - hh, cx + hw, cy + hh), (float(cx), float(cy))
        sample_note = f"central {2 * hw + 1}x{2 * hh + 1} px of the {d[3] - d[1] + 1}^2 domain, one cold step"
    elif d[0] == "annulus":
        xy, center = oracle.annulus_points(*d[1:]), None
    else:  # subsets: a bounded sample of the 4096 subsets, run one after the other like the manager
        all_boxes = subset_boxes(d[1], d[2], d[3])
        boxes = [all_boxes[i] for i in stratified_sample(len(all_boxes), 64)]
        xy, center = None, None
    n_par = 12 if w["model"] == "quad" else 6
    if counter is not None:
        counter.set_image("und", und)
        counter.set_image("def", dfm)
    t0 = time.perf_counter()
    eng.set_image("und", und)
    eng.set_image("def", dfm)
    t_pyr = time.perf_counter() - t0
    work, secs, res = 0.0, t_pyr, None
    if xy is not None:
        t1 = time.perf_counter()
        res = eng.correlate(np.zeros(n_par, np.float32), xy, center=center)
        secs += time.perf_counter() - t1
        work = res.get("pixel_evaluations") or counter.correlate(np.zeros(n_par, np.float32), xy, center=center)["pixel_evaluations"]
        sample = sample_note or "whole workload, one cold step (pyramids + Newton_Raphson)"
    else:
        for bx in boxes:
            pts = oracle.rect_points(*bx)
            c = ((bx[0] + bx[2]) / 2.0, (bx[1] + bx[3]) / 2.0)
            t1 = time.perf_counter()
            res = eng.correlate(np.zeros(n_par, np.float32), pts, center=c)
            secs += time.perf_counter() - t1
            work += res.get("pixel_evaluations") or counter.correlate(np.zeros(n_par, np.float32), pts, center=c)["pixel_evaluations"]
        # the two whole-image pyramid builds serve all 4096 subsets: the sample is charged its share of them, so that
        # work / seconds is the rate the whole workload would run at (pyramids + Newton_Raphson of every subset)
        share = len(boxes) / float(len(all_boxes))
        secs -= t_pyr * (1.0 - share)
        sample = (f"{len(boxes)} of {d[3] * d[3]} subsets (stratified), one after the other as the manager does, one cold "
                  f"step: Newton_Raphson per subset ({threads} threads per evaluation) + {share:.4f} of the two 8192^2 "
                  f"pyramid builds ({t_pyr:.2f} s for all subsets)")
    return work, secs, ("reference" if use_ref else "port"), res, sample, t_pyr


# ------------------------------------------------------------------------------ config 3: frames / s

def run_c3(args, w, with_cpu=True):
    """The frame loop of the headless C++ host (correlation_b200/host/dic_manager.hpp) on pinned host
    frames: per frame H2D of the next image (second stream, overlapped), pyramid, rotation, GN."""
    import torch
    from correlation_b200 import host, synth
    dev = torch.device("cuda", 0)
    n = w["frames"]
    frames = []
    for k in range(n):
        t = synth.make_image(w["rows"], w["cols"], w["seed"], None if k == 0 else tuple(k * np.array(w["rate"])),
                             w["center"], device=dev)
        pin = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
        pin.copy_(t)
        frames.append(pin.numpy())
    torch.cuda.synchronize()
    d = w["domain"]
    contour = synth.star_polygon(d[1], d[2], d[3], n_vertices=d[4], seed=w["seed"])
    mode = 0 if args.mode == "parity" else 1
    runs = []
    with ClockSampler(0) as clk:
        for i in range(args.warmup + args.steps):
            r = host.run_sequence(frames, contour=contour, pyramid=w["pyramid"], arith_mode=mode)
            if i >= args.warmup:
                runs.append(r)
    secs_all = [r["seconds"] for r in runs]
    secs = sum(secs_all) / len(runs)
    fps = (n - 1) / secs
    hdr, rows = host.parse_report(runs[-1]["csv"])
    last = runs[-1]["rows"][0]
    truth_last = (n - 1) * np.array(w["rate"])
    line = {"metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic speckle sequence, constant velocity",
            "config": {"workload": w["name"], "arith_mode": args.mode, "frame_pairs": n - 1,
                       "points": int(last["number_of_points"]), "errors": int(sum(int(r["error_code"]) != 0 for r in rows)),
                       "ms_per_frame": 1e3 * secs / (n - 1),
                       "ms_per_frame_spread": [1e3 * min(secs_all) / (n - 1), 1e3 * max(secs_all) / (n - 1)],
                       "ms_per_frame_median_run": 1e3 * float(np.median(secs_all)) / (n - 1),
                       "last_frame_params": [float(v) for v in last["params"][:6]],
                       "last_frame_truth": [float(v) for v in truth_last],
                       "step": "one step = the whole 99-pair sequence through dic_host_run (C++ host loop)"},
            "clocks": clk.summary(),
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": n * w["rows"] * w["cols"],
                    "d2h_bytes_per_step": 176 * (n - 1)},
            "gpu_launches": (n - 1) * 4,
            "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "frames/s is a pipeline metric (upload + pyramid + GN per frame); see the c4 / c2 lines for the kernel roofline"}}
    if with_cpu:
        import oracle
        # the same bookkeeping on the CPU oracle (fp64 accumulators): first frames timed as the baseline, and the
        # GPU's report rows of those frames checked against it
        t0 = time.perf_counter()
        O = oracle.OracleEngine(n_threads=os.cpu_count() or 1, pyramid=w["pyramid"], real_threads=True, accum_double=True)
        O.set_image("und", frames[0])
        xy = oracle.blob_points(contour)
        p = np.zeros(6, np.float32)
        p_prev = p.copy()
        nf = 4
        want = []
        for k in range(nf):
            O.set_image("def", frames[k + 1])
            guess = p + (p - p_prev) if k else p
            p_prev = p
            r = O.correlate(guess, xy)
            p = r["params"]
            want.append(r)
        cs = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": nf / cs, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"first {nf} frame pairs of the sequence (pyramid + Newton_Raphson each)"}
        gp = np.array([[rows[k][f"parameter_{q}"] for q in range(6)] for k in range(nf)])
        blk = parity_block(
            gp, [rows[k]["chi"] for k in range(nf)], [rows[k]["iterations"] for k in range(nf)],
            [r["params"] for r in want], [r["chi"] for r in want], [r["iterations"] for r in want])
        line["parity"] = {"vs_oracle_first_frames": blk,
            "note": "GPU values are the 6-significant-digit CSV report rows of the first frames (constant-velocity guesses "
                    "included); the oracle runs the same sequence with fp64 accumulators"}
        def ref_frames(chunks):  # the same frames and the fp64 run's guesses, fp32 accumulators
            R = oracle.OracleEngine(n_threads=chunks, pyramid=w["pyramid"], accum_double=False)
            R.set_image("und", frames[0])
            refs = []
            q, q_prev = np.zeros(6, np.float32), np.zeros(6, np.float32)
            for k in range(nf):
                R.set_image("def", frames[k + 1])
                guess = q + (q - q_prev) if k else q
                q_prev = q
                refs.append(R.correlate(guess, xy))
                q = want[k]["params"]
            return refs
        line["parity"].update(reference_spread(ref_frames, want, blk, "same frames and guesses"))
    return line


# ------------------------------------------------------------------------------ one workload on the GPU(s)

class Run:
    """One workload set up on this rank's GPU: images resident and pinned, engine, domain(s) built."""

    def __init__(self, args, w, dist, rank, world, local_rank):
        import torch
        from correlation_b200 import engine
        self.args, self.w, self.dist, self.rank, self.world = args, w, dist, rank, world
        self.torch, self.engine = torch, engine
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        self.n_par = 12 if w["model"] == "quad" else 6
        self.und_t, self.dfm_t = make_images(w, self.dev)
        self.und_pin = torch.empty(self.und_t.shape, dtype=torch.uint8, pin_memory=True)
        self.dfm_pin = torch.empty(self.dfm_t.shape, dtype=torch.uint8, pin_memory=True)
        self.und_pin.copy_(self.und_t)
        self.dfm_pin.copy_(self.dfm_t)
        torch.cuda.synchronize()
        self.mode = engine.MODE_PARITY if args.mode == "parity" else engine.MODE_FAST
        self.eng = engine.CudaEngine(local_rank, fitting_model=engine.FM_QUADRATIC if w["model"] == "quad" else engine.FM_UVUxUyVxVy,
                                     arith_mode=self.mode)
        rows, cols = w["rows"], w["cols"]
        self.eng.resetImagePyramidsDevice(self.und_t.data_ptr(), self.dfm_t.data_ptr(), None, rows, cols, cols, pyramid=w["pyramid"])
        d = w["domain"]
        self.scaling = "weak"
        self.all_boxes = self.my_ids = None
        self.band_rows = None
        t_dom = time.perf_counter()
        if d[0] in ("rect", "annulus"):
            self.eng.resetPolygon(0, *d[1:])
            self.n_sectors = 1
        elif d[0] == "rowsplit":
            # one domain, pixel rows in equal bands per rank, per-evaluation sum inside the kernel
            from correlation_b200 import rowsplit
            rowsplit.connect(self.eng, dist)
            self.band_rows = rowsplit.equal_row_bands(d[2], d[4], world)[rank]
            self.eng.resetPolygonRectBand(0, d[1], d[2], d[3], d[4], *self.band_rows)
            self.n_sectors = 1
            self.scaling = "strong"
        else:
            # independent subsets shard across ranks in whole rows of subsets: each rank then needs
            # only a band of image rows for the end-to-end path; no collective on the data path
            from correlation_b200 import sharding
            self.all_boxes = subset_boxes(d[1], d[2], d[3])
            self.my_ids = sharding.shard_grid_rows(d[3], d[3], world, rank)
            self.eng.resetPolygonRectGrid(0, np.array([self.all_boxes[i] for i in self.my_ids], np.int32))
            self.n_sectors = len(self.my_ids)
            self.scaling = "strong"
        self.eng.synchronize()
        self.domain_build_s = time.perf_counter() - t_dom
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)
        self.res_buf = np.zeros(self.n_sectors, engine.RESULT_DTYPE)
        self.guess_buf = np.zeros((self.n_sectors, self.n_par), np.float32)
        self.one_guess = np.zeros(self.n_par, np.float32)
        self.one_result = engine.DicResult()
        # a rank that owns a band of the image (sharded subsets, row-split domain) transfers only its rows plus
        # a halo for displacement, bicubic support and pyramid support (dic_stage_next_pair_rows)
        self.band = None
        if world > 1 and d[0] == "rowsplit":
            self.band = upload_band(self.band_rows[0], self.band_rows[1], rows, w["pyramid"][2])
        elif world > 1 and d[0] == "subsets":
            mine = [self.all_boxes[i] for i in self.my_ids]
            self.band = upload_band(min(bx[1] for bx in mine), max(bx[3] for bx in mine), rows, w["pyramid"][2])
        self.h2d_bytes = 2 * cols * ((self.band[1] - self.band[0]) if self.band else rows)

    # -- one step, inputs resident
    def step_resident(self):
        eng, engine = self.eng, self.engine
        if self.n_sectors == 1:
            self.one_guess[:] = 0.0
            eng.correlate_raw(0, self.one_guess, self.one_result)
            work = 0.0
            for lv in range(engine.MAX_LEVELS):
                work += float(self.one_result.evaluationsPerLevel[lv]) * float(self.one_result.pointsPerLevel[lv])
            if self.w["domain"][0] == "rowsplit":
                work /= self.world  # the record counts the whole domain; this rank evaluated its band of it
            return work, eng.last_correlate_ms(), self.one_result
        self.guess_buf[:] = 0.0
        eng.lib.dic_correlate_batch(eng.h, 0, self.n_sectors, self.guess_buf.ctypes.data, self.res_buf.ctypes.data)
        return eng.pixel_evaluations(self.res_buf), eng.last_correlate_ms(), self.res_buf

    def stage_pair(self):
        # this step's inputs: both images from pinned host memory, upload + pyramids on the copy / image streams
        self.eng.stageNextPair(self.und_pin.data_ptr(), self.dfm_pin.data_ptr(), self.w["rows"], self.w["cols"], row_range=self.band)

    def enqueue_step(self):
        eng = self.eng
        if self.n_sectors == 1:
            self.one_guess[:] = 0.0
            eng.lib.dic_correlate_async(eng.h, 0, self.one_guess.ctypes.data)
        else:
            self.guess_buf[:] = 0.0
            eng.lib.dic_correlate_batch_async(eng.h, 0, self.n_sectors, self.guess_buf.ctypes.data)

    def wait_step(self):
        # blocks until the step's result record(s) are in this process's own buffers (the D2H read of the step)
        eng = self.eng
        if self.n_sectors == 1:
            eng.lib.dic_correlate_wait(eng.h, 0, self.one_guess.ctypes.data, ctypes.byref(self.one_result))
        else:
            eng.lib.dic_correlate_batch_wait(eng.h, 0, self.n_sectors, None, self.res_buf.ctypes.data)

    def step_work(self):
        eng, engine = self.eng, self.engine
        if self.n_sectors == 1:
            work = 0.0
            for lv in range(engine.MAX_LEVELS):
                work += float(self.one_result.evaluationsPerLevel[lv]) * float(self.one_result.pointsPerLevel[lv])
            return work / self.world if self.w["domain"][0] == "rowsplit" else work
        return eng.pixel_evaluations(self.res_buf)

    def e2e_loop(self, n):
        # double-buffered ingest (dic_stage_next_pair / dic_advance_pair): the PCIe transfer of pair k + 1
        # overlaps the solve of pair k; every step copies its own pair and reads its own result record(s).
        # The solve is enqueued BEFORE the next pair is staged, so that the host work of the staging call is
        # off the solve's critical path (dic_correlate[_batch]_async / _wait, include/dic_b200.h).
        # Two pairs are staged ahead: the transfer of pair k + 2 is on the copy stream's queue before the host
        # starts to wait for solve k, so the bus does not idle while the host is blocked.
        # As soon as the records of step k are in the caller's buffer the next solve is enqueued; the host's own work
        # on those records (here: summing points x evaluations) runs while that solve does.
        tot = 0.0
        for _ in range(min(2, n)):
            self.stage_pair()
        self.eng.advancePair()
        self.enqueue_step()
        for k in range(n):
            if k + 2 < n:
                self.stage_pair()
            self.wait_step()
            if k + 1 < n:
                self.eng.advancePair()
                self.enqueue_step()
            tot += self.step_work()
        return tot

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def last_record(self):
        if self.n_sectors == 1:
            return self.one_result.as_dict(self.n_par)
        rs = self.res_buf
        return dict(params=rs[0]["resultingParameters"][:self.n_par].copy(), chi=rs[0]["chi"],
                    iterations=int(rs[0]["iterations"]), evaluations=rs[0]["evaluationsPerLevel"].tolist(),
                    points_per_level=rs[0]["pointsPerLevel"].tolist(), errors=int((rs["errorCode"] != 0).sum()))

    def measure(self, steps, warmup, with_other_mode=True):
        """Timed resident steps, timed end-to-end steps, the other arithmetic mode. Returns a dict of reduced numbers
        (identical on every rank for the reduced entries)."""
        torch, eng, dist = self.torch, self.eng, self.dist
        for _ in range(warmup):
            self.step_resident()
        launches0 = eng.kernel_launches()
        self.barrier()
        # Timed on the DEVICE: CUDA events on the correlation stream around the whole GPU side of each step (solve
        # kernel, result download), summed over the K steps, max over ranks. The L2 flush between steps is outside
        # the events. The host clock around the same calls is reported beside it (host_ms_per_step).
        work = kern_ms = wall = host_wall = 0.0
        with ClockSampler(self.dev.index, self.world) as clk:
            for _ in range(steps):
                self.flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                wk, ms, _ = self.step_resident()
                host_wall += time.perf_counter() - t0
                wall += 1e-3 * eng.last_step_ms()
                work += wk
                kern_ms += ms
                if self.world == 1:
                    clk.poke()
        self.barrier()
        launches = eng.kernel_launches() - launches0
        last = self.last_record()
        resident_records = self.res_buf.copy() if self.n_sectors > 1 else None
        # e2e: host buffers, copies inside the timed region
        # The loop is a pipeline: its first pair's transfer and its last solve overlap with nothing, and both are
        # inside the timed region. It runs for at least 32 steps so that this fill / drain (about one step) is
        # amortised as it would be in a frame sequence; the count is reported as e2e.steps.
        self.e2e_steps = e2e_steps = max(steps, 32)
        self.e2e_loop(3)
        self.barrier()
        t0 = time.perf_counter()
        e2e_work = self.e2e_loop(e2e_steps)
        self.barrier()
        e2e_wall = time.perf_counter() - t0
        o_val = None
        if with_other_mode:  # the other arithmetic mode, same resident inputs, for the record
            other = self.engine.MODE_PARITY if self.mode == self.engine.MODE_FAST else self.engine.MODE_FAST
            eng.set_arith_mode(other)
            for _ in range(2):
                self.step_resident()
            o_work = o_ms = 0.0
            for _ in range(max(2, steps // 2)):
                self.flush.fill_(1)
                torch.cuda.synchronize()
                wk, ms, _ = self.step_resident()
                o_work += wk
                o_ms += ms
            eng.set_arith_mode(self.mode)
            o_val = o_work / (o_ms * 1e-3) if o_ms > 0 else None
        stats = torch.tensor([wall, e2e_wall, work, e2e_work, kern_ms, host_wall], dtype=torch.float64, device=self.dev)
        my_work, my_kern_ms = work, kern_ms
        if dist is not None:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            wall, e2e_wall, host_wall = mx[0].item(), mx[1].item(), mx[5].item()
            work, e2e_work = sm[2].item(), sm[3].item()
        return dict(wall=wall, e2e_wall=e2e_wall, host_wall=host_wall, work=work, e2e_work=e2e_work, my_work=my_work,
                    my_kern_ms=my_kern_ms, launches=launches, last=last, clocks=clk.summary(), other_mode_value=o_val,
                    resident_records=resident_records)

    # -- parity
    def gather_records(self, local):
        """All ranks' records in global sector order on rank 0 (None elsewhere). Collective."""
        n_total = len(self.all_boxes)
        if self.dist is None:
            full = np.zeros(n_total, local.dtype)
            full[self.my_ids] = local
            return full
        torch = self.torch
        item = local.dtype.itemsize
        per = -(-n_total // self.world) + 64
        buf = torch.zeros(per * item, dtype=torch.uint8, device=self.dev)
        raw = torch.from_numpy(np.frombuffer(local.tobytes(), np.uint8).copy())
        buf[: raw.numel()] = raw.to(self.dev)
        out = [torch.zeros_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf)
        if self.rank != 0:
            return None
        from correlation_b200 import sharding
        d = self.w["domain"]
        full = np.zeros(n_total, local.dtype)
        for r in range(self.world):
            ids = sharding.shard_grid_rows(d[3], d[3], self.world, r)
            full[ids] = np.frombuffer(out[r].cpu().numpy().tobytes()[: len(ids) * item], local.dtype)
        return full

    def parity_subsets(self, records, n_sample):
        """records: all subsets in global order (rank 0). Three things on a stratified sample:
        vs_oracle             GPU against the restatement with fp64 accumulators (the BASELINE.json tolerances);
        reference_self_spread the restatement with the reference's OWN arithmetic (fp32 accumulators in 20 thread
                              chunks, correlation_class.cpp:131-300) against that same fp64-accumulator restatement:
                              how far the reference sits from its own noise-free version. The chi reported for a 125^2
                              subset is chi at the last accepted step, one GN step short of the reported parameters, and
                              moves by ~1e-4 relative with 1e-7-level rounding anywhere upstream (solver, summation
                              order; tools/lm_trace.py, tools/chi_diag.py) -- in the reference as much as here."""
        import oracle
        w = self.w
        und, dfm = self.und_pin.numpy(), self.dfm_pin.numpy()
        kw = dict(model=oracle.FM_AFFINE, pyramid=w["pyramid"])
        o64 = oracle.OracleEngine(n_threads=os.cpu_count() or 1, accum_double=True, real_threads=True, **kw)
        o64.set_image("und", und)
        o64.set_image("def", dfm)
        ids = stratified_sample(len(self.all_boxes), n_sample)
        inputs = []
        for i in ids:
            bx = self.all_boxes[i]
            inputs.append(((np.zeros(6, np.float32), oracle.rect_points(*bx)), ((bx[0] + bx[2]) / 2.0, (bx[1] + bx[3]) / 2.0)))
        want = [o64.correlate(*a, center=c) for a, c in inputs]
        evs = lambda rs: [r["evaluations"][:8] for r in rs]
        blk = parity_block(records["resultingParameters"][ids, :6], records["chi"][ids], records["iterations"][ids],
                           [r["params"] for r in want], [r["chi"] for r in want], [r["iterations"] for r in want],
                           gpu_evals=records["evaluationsPerLevel"][ids], want_evals=evs(want))
        blk["errors"] = [int((records["errorCode"][ids] != 0).sum()), int(sum(r["error_code"] != 0 for r in want))]
        blk["sample"] = f"{len(ids)} of {len(self.all_boxes)} subsets, stratified over the sector ids"

        def ref_subsets(chunks):
            o32 = oracle.OracleEngine(n_threads=chunks, accum_double=False, real_threads=False, **kw)
            o32.set_image("und", und)
            o32.set_image("def", dfm)
            return [o32.correlate(*a, center=c) for a, c in inputs]
        out = {"vs_oracle": blk}
        out.update(reference_spread(ref_subsets, want, blk, "same subsets"))
        return out

    def close(self):
        try:
            if self.w["domain"][0] == "rowsplit":
                self.eng.rowsplit_disconnect()
        except Exception:
            pass
        self.eng.close()
        del self.flush, self.und_t, self.dfm_t
        self.torch.cuda.empty_cache()


def roofline_block(args, w, m, peaks):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:  # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["gn_solve_tiles_kernel"].get(f"{w['key']} {args.mode}")
        if tr:
            traffic = tr["dram_bytes_read_per_launch"] + tr["dram_bytes_write_per_launch"]
    except Exception:
        pass
    achieved = ALGO_BYTES_PER_PIXEL_EVAL * m["my_work"] / (m["my_kern_ms"] * 1e-3) / 1e9 if m["my_kern_ms"] > 0 else 0.0
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_note": "bytes per launch on ONE GPU at N = 1 (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json); "
                            "algorithmic bytes per launch = 10 B x this rank's pixel_evaluations_per_step",
            "kernel": "gn_solve_tiles_kernel", "kernel_ms_per_step": m["my_kern_ms"] / max(1, args.steps_used),
            "rank_pixel_evaluations_per_step": m["my_work"] / max(1, args.steps_used),
            "fp32_note": "the kernel is FP32-issue bound, not HBM bound (DESIGN.md 4.1): parity mode replays the reference's "
                         "unfused fp32 operations per pixel; the level data stays L2-resident across evaluations. "
                         "ncu summaries under profiles/",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
            "algorithmic_bytes_per_pixel_evaluation": ALGO_BYTES_PER_PIXEL_EVAL}


def record_for(args, run, m, steps, peaks):
    """The contract's JSON fields for one measured workload (rank 0)."""
    w, world = run.w, run.world
    d = w["domain"]
    args.steps_used = steps
    last = m["last"]
    par = (f"one domain in {world} row band(s), per-evaluation all-reduce of the normal equations inside the kernel (NVLink peer mailboxes)"
           if d[0] == "rowsplit" else
           f"{d[3] * d[3]} subsets, whole rows of subsets per GPU, over {world} GPU(s), no collective"
           if d[0] == "subsets" else f"{world} independent domain(s), one per GPU")
    return {
        "metric": "domain pixel*GN-evaluations/s", "value": m["work"] / m["wall"], "unit": "pixel*evaluations/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * m["wall"] / steps,
        "host_ms_per_step": 1e3 * m["host_wall"] / steps,
        "timing": "CUDA events on the correlation stream around each step's device work, summed over the steps, max over ranks",
        "higher_is_better": True, "scaling": run.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic analytic speckle (SURVEY 8d), random phases seeded",
        "config": {"workload": w["name"], "arith_mode": args.mode, "sectors_this_rank": run.n_sectors,
                   "pixel_evaluations_per_step": m["work"] / steps,
                   "evaluations_per_level": last["evaluations"][: w["pyramid"][2] + 1],
                   "points_per_level": last["points_per_level"][: w["pyramid"][2] + 1],
                   "l2": "flushed between steps (512 MiB write); evaluations inside a step re-read the domain by design",
                   "domain_build_s": run.domain_build_s, "parallelism": par},
        "clocks": m["clocks"],
        "e2e": {"value": m["e2e_work"] / m["e2e_wall"], "unit": "pixel*evaluations/s",
                "h2d_bytes_per_step": run.h2d_bytes, "d2h_bytes_per_step": 176 * run.n_sectors,
                "ms_per_step": 1e3 * m["e2e_wall"] / run.e2e_steps, "steps": run.e2e_steps,
                "pipeline": "dic_stage_next_pair(k + 2) on the copy / image streams (two pairs ahead); dic_correlate[_batch]_wait(k); dic_advance_pair; dic_correlate[_batch]_async(k + 1): H2D of both images every step, result record(s) read every step"
                            + ("" if run.band is None else f" (this rank's row band {run.band[0]}..{run.band[1]} of {w['rows']}; bytes are per rank)")},
        "gpu_launches": m["launches"],
        "other_arith_mode": {"arith_mode": "fast" if args.mode == "parity" else "parity",
                             "kernel_value_this_rank": m["other_mode_value"],
                             "unit": "pixel*evaluations/s (kernel time, one rank)"},
        "roofline": roofline_block(args, w, m, peaks),
    }


def single_domain_parity(run, last, threads):
    """c1 / c2: the whole domain against the fp64-accumulator oracle (and the fp32 CPU engine's own spread)."""
    import oracle
    w, d, n_par = run.w, run.w["domain"], run.n_par
    und, dfm = run.und_pin.numpy(), run.dfm_pin.numpy()
    od = oracle.OracleEngine(model=oracle.FM_QUAD if w["model"] == "quad" else oracle.FM_AFFINE, n_threads=threads,
                             pyramid=w["pyramid"], accum_double=True, real_threads=True)
    od.set_image("und", und)
    od.set_image("def", dfm)
    if d[0] == "rect":
        dres = od.correlate(np.zeros(n_par, np.float32), oracle.rect_points(*d[1:]), center=((d[1] + d[3]) / 2.0, (d[2] + d[4]) / 2.0))
    else:
        dres = od.correlate(np.zeros(n_par, np.float32), oracle.annulus_points(*d[1:]))
    blk = parity_block(last["params"], last["chi"], last["iterations"], dres["params"], dres["chi"], dres["iterations"],
                       gpu_evals=[last["evaluations"][:8]], want_evals=[dres["evaluations"][:8]])
    blk["evaluations"] = [last["evaluations"][: w["pyramid"][2] + 1], dres["evaluations"][: w["pyramid"][2] + 1]]
    out = {"vs_oracle": blk}
    def ref_domain(chunks):
        orf = oracle.OracleEngine(model=oracle.FM_QUAD if w["model"] == "quad" else oracle.FM_AFFINE, n_threads=chunks,
                                  pyramid=w["pyramid"], accum_double=False)
        orf.set_image("und", und)
        orf.set_image("def", dfm)
        if d[0] == "rect":
            return [orf.correlate(np.zeros(n_par, np.float32), oracle.rect_points(*d[1:]), center=((d[1] + d[3]) / 2.0, (d[2] + d[4]) / 2.0))]
        return [orf.correlate(np.zeros(n_par, np.float32), oracle.annulus_points(*d[1:]))]
    out.update(reference_spread(ref_domain, [dres], blk, "same domain"))
    return out


def other_workload_c2(args, peaks):
    """Short run of config 2 on rank 0's GPU (N = 1 only)."""
    sub = argparse.Namespace(**vars(args))
    w = workload("c2")
    run = Run(sub, w, None, 0, 1, 0)
    steps = max(3, min(args.steps, 8))
    m = run.measure(steps, 3, with_other_mode=True)
    rec = record_for(sub, run, m, steps, peaks)
    try:
        rec["parity"] = single_domain_parity(run, m["last"], os.cpu_count() or 1)
    except Exception as ex:
        rec["parity"] = {"error": repr(ex)}
    run.close()
    return rec


def other_workload_c5(args, dist, rank, world, local_rank, peaks):
    """Short run of config 5 row-split over all ranks; rank 0 also solves the whole domain alone and a sample
    region against the oracle. Collective."""
    import torch
    sub = argparse.Namespace(**vars(args))
    w = workload("c5")
    run, err = None, None
    try:
        run = Run(sub, w, dist, rank, world, local_rank)
    except Exception as ex:  # one rank failing its set-up must not leave the others waiting in a collective
        err = repr(ex)
    ok = torch.tensor([0 if err else 1], device=torch.device("cuda", local_rank))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        if run is not None:
            run.close()
        return {"error": err or "set-up failed on another rank"}
    steps = max(3, min(args.steps, 5))
    m = run.measure(steps, 3, with_other_mode=False)
    split = m["last"]
    rec = None
    if rank == 0:
        rec = record_for(sub, run, m, steps, peaks)
    # parity_vs_single_gpu: the same domain on rank 0 alone (plain grid launch, no exchange)
    run.barrier()
    if rank == 0:
        import oracle
        d = w["domain"]
        e1 = run.engine.CudaEngine(local_rank, fitting_model=run.engine.FM_UVUxUyVxVy, arith_mode=run.mode)
        e1.resetImagePyramidsDevice(run.und_t.data_ptr(), run.dfm_t.data_ptr(), None, w["rows"], w["cols"], w["cols"], pyramid=w["pyramid"])
        e1.resetPolygon(0, d[1], d[2], d[3], d[4])
        one = e1.correlate(0, np.zeros(6, np.float32))
        blk = parity_block(split["params"], split["chi"], split["iterations"], one["params"], one["chi"], one["iterations"],
                           gpu_evals=[split["evaluations"][:8]], want_evals=[one["evaluations"][:8]])
        blk["evaluations"] = [split["evaluations"][:5], one["evaluations"][:5]]
        blk["bitwise_equal"] = bool(np.array_equal(split["params"], one["params"]) and split["chi"] == one["chi"])
        rec["parity"] = {"vs_single_gpu": blk}
        # sample region against the oracle (64-bit restatement; the reference's own cache index overflows at this size)
        try:
            cx, cy = (d[1] + d[3]) // 2, (d[2] + d[4]) // 2
            hw = (d[3] - d[1]) // 16
            box = (cx - hw, cy - hw, cx + hw, cy + hw)
            e1.resetPolygon(1, *box)
            g = e1.correlate(1, np.zeros(6, np.float32))
            o = oracle.OracleEngine(model=oracle.FM_AFFINE, n_threads=os.cpu_count() or 1, pyramid=w["pyramid"], accum_double=True, real_threads=True)
            o.set_image("und", run.und_pin.numpy())
            o.set_image("def", run.dfm_pin.numpy())
            want = o.correlate(np.zeros(6, np.float32), oracle.rect_points(*box), center=(float(cx), float(cy)))
            sblk = parity_block(g["params"], g["chi"], g["iterations"], want["params"], want["chi"], want["iterations"],
                                gpu_evals=[g["evaluations"][:8]], want_evals=[want["evaluations"][:8]])
            sblk["region"] = f"central {2 * hw + 1}^2 px of the 16384^2 pair, 5 levels, one GPU vs the oracle with fp64 accumulators"
            sblk["evaluations"] = [g["evaluations"][:5], want["evaluations"][:5]]
            rec["parity"]["sample_region_vs_oracle"] = sblk
            def ref_region(chunks):
                o32 = oracle.OracleEngine(model=oracle.FM_AFFINE, n_threads=chunks, pyramid=w["pyramid"], accum_double=False)
                o32.set_image("und", run.und_pin.numpy())
                o32.set_image("def", run.dfm_pin.numpy())
                return [o32.correlate(np.zeros(6, np.float32), oracle.rect_points(*box), center=(float(cx), float(cy)))]
            rec["parity"].update(reference_spread(ref_region, [want], sblk, "same region"))
        except Exception as ex:
            rec["parity"]["sample_region_vs_oracle"] = {"error": repr(ex)}
        e1.close()
    run.barrier()
    run.close()
    return rec


# ------------------------------------------------------------------------------ main

def bind_near_gpu(local_rank):
    """Pin this rank's threads to the CPU cores next to its GPU (nvmlDeviceGetCpuAffinity) BEFORE any pinned buffer is
    allocated, so that the buffers land on that socket's memory (first touch). With eight ranks streaming images out of
    host memory at once, round 1 measured one shared ceiling of ~165 GB/s when every rank ran wherever the scheduler
    put it. Returns a description for the bench line, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(ClockSampler._physical_index(local_rank))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus or len(cpus) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cores next to the GPU (nvmlDeviceGetCpuAffinity), set before the pinned buffers are allocated"
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="parity", choices=["parity", "fast"],
                    help="arithmetic form of the per-pixel evaluation (include/dic_b200.h dic_arith_mode). 'parity' (default, "
                         "the headline) replays the reference's fp32 operation order: w, dw/dx, dw/dy bit-identical per pixel, "
                         "every BASELINE.json tolerance met. 'fast' is the same interpolant in Catmull-Rom form (~2x fewer "
                         "instructions, more accurate than the reference's own rounding noise, which is why its chi can sit "
                         "1e-5 away from the reference's); it is reported beside the headline, labelled")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads sub-records (c2 and c3 at N = 1, c5 at N > 1)")
    ap.add_argument("--parity-sample", type=int, default=64, help="subsets compared with the oracle in the bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = workload(args.workload)
    if args.workload == "c3" and args.impl == "ours":
        if rank != 0:
            return 0
        print(json.dumps(run_c3(args, w, with_cpu=not args.no_cpu_baseline)), flush=True)
        return 0

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.impl == "reference" and rank != 0:
            return 0
        if args.impl == "ours":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.impl == "reference":
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else None
        if dev is None:
            from correlation_b200 import synth
            und = synth.make_image(w["rows"], w["cols"], w["seed"], None, w["center"])
            dfm = synth.make_image(w["rows"], w["cols"], w["seed"], w["truth"], w["center"])
        else:
            und_t, dfm_t = make_images(w, dev)
            und, dfm = und_t.cpu().numpy(), dfm_t.cpu().numpy()
        threads = os.cpu_count() or 1
        vals = []
        for i in range(args.warmup + args.steps):
            work, secs, kind, res, sample, _ = cpu_run(w, und, dfm, threads)
            if i >= args.warmup:
                vals.append((work, secs))
        tot_w, tot_s = sum(v[0] for v in vals), sum(v[1] for v in vals)
        v = tot_w / tot_s
        line = {"impl": "reference", "metric": "domain pixel*GN-evaluations/s", "value": v,
                "unit": "pixel*evaluations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True,
                "scaling": "strong" if w["domain"][0] in ("subsets", "rowsplit") else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic speckle (SURVEY 8d)",
                "config": {"workload": w["name"]},
                "cpu_baseline": {"value": v, "unit": "pixel*evaluations/s", "cores": threads, "kind": kind,
                                 "sample": sample},
                "e2e": {"value": v, "unit": "pixel*evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- our arm
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if w["domain"][0] == "rowsplit" and world == 1 and not args.no_others:
        args.no_others = True
    affinity = bind_near_gpu(local_rank) if world > 1 else None
    run = Run(args, w, dist, rank, world, local_rank)
    m = run.measure(args.steps, args.warmup)
    line = record_for(args, run, m, args.steps, peaks) if rank == 0 else None
    if rank == 0 and affinity:
        line["config"]["cpu_affinity"] = affinity
    d = w["domain"]

    # ---- parity (top level of the line)
    parity = {}
    if d[0] == "subsets":
        full = run.gather_records(m["resident_records"])  # collective
        if rank == 0:
            try:
                parity.update(run.parity_subsets(full, args.parity_sample))
            except Exception as ex:
                parity["vs_oracle"] = {"error": repr(ex)}
            if world > 1:
                # rank 0 alone, all 4096 subsets, the same CTA mapping the ranks used: every subset's arithmetic is then
                # the same instruction sequence on the same data, so the records must agree bit for bit
                try:
                    eng = run.eng
                    pair_used = eng.last_cluster_size() == 2
                    n_all = len(run.all_boxes)
                    # the end-to-end loop left only this rank's band of rows in the current slots: whole images again
                    eng.resetImagePyramidsDevice(run.und_t.data_ptr(), run.dfm_t.data_ptr(), None, w["rows"], w["cols"],
                                                 w["cols"], pyramid=w["pyramid"])
                    base = run.n_sectors
                    eng.resetPolygonRectGrid(base, np.array(run.all_boxes, np.int32))
                    eng.set_cluster_mode(2 if pair_used else 1)
                    g = np.zeros((n_all, run.n_par), np.float32)
                    one = np.zeros(n_all, run.engine.RESULT_DTYPE)
                    eng.lib.dic_correlate_batch(eng.h, base, n_all, g.ctypes.data, one.ctypes.data)
                    eng.set_cluster_mode(0)
                    same = (one.tobytes() == full.tobytes())
                    neq = int(sum(one[k].tobytes() != full[k].tobytes() for k in range(n_all)))
                    parity["vs_single_gpu"] = {
                        "records_compared": n_all, "records_bitwise_equal": n_all - neq, "bitwise_equal": bool(same),
                        "ctas_per_subset": 2 if pair_used else 1,
                        "max_abs_dparams": float(np.abs(one["resultingParameters"] - full["resultingParameters"]).max()),
                        "max_rel_dchi": float((np.abs(one["chi"] - full["chi"]) / np.maximum(np.abs(one["chi"]), 1e-30)).max()),
                        "note": "all ranks' result records gathered in sector order vs rank 0's own single-GPU batch of all subsets"}
                except Exception as ex:
                    parity["vs_single_gpu"] = {"error": repr(ex)}
    elif d[0] in ("rect", "annulus") and rank == 0 and not args.no_cpu_baseline:
        try:
            parity = single_domain_parity(run, m["last"], os.cpu_count() or 1)
        except Exception as ex:
            parity = {"vs_oracle": {"error": repr(ex)}}
    if rank == 0:
        line["parity"] = parity

    # ---- CPU baseline (N = 1, rank 0): the reference's own engine on a bounded sample of the same workload
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            threads = os.cpu_count() or 1
            cw, cs, kind, cres, sample, t_pyr = cpu_run(w, run.und_pin.numpy(), run.dfm_pin.numpy(), threads)
            line["cpu_baseline"] = {"value": cw / cs, "unit": "pixel*evaluations/s", "cores": threads, "kind": kind,
                                    "sample": sample, "seconds": cs, "pyramid_seconds": t_pyr}
        except Exception as ex:  # the baseline is reported, never allowed to sink the bench line
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    run.barrier()
    run.close()

    # ---- other workloads, short, for continuity with round 1 (each failure is recorded, never fatal)
    if not args.no_others and d[0] == "subsets":
        others = {}
        if world == 1:
            try:
                others["c2"] = other_workload_c2(args, peaks)
            except Exception as ex:
                others["c2"] = {"error": repr(ex)}
            try:  # frames/s, the second half of BASELINE's metric: the 100-frame sequence through the C++ host loop
                sub = argparse.Namespace(**vars(args))
                sub.steps, sub.warmup = max(2, min(args.steps, 4)), 1
                others["c3"] = run_c3(sub, workload("c3"), with_cpu=not args.no_cpu_baseline)
            except Exception as ex:
                others["c3"] = {"error": repr(ex)}
        else:
            try:
                rec = other_workload_c5(args, dist, rank, world, local_rank, peaks)
                if rank == 0:
                    others["c5"] = rec
            except Exception as ex:
                others["c5"] = {"error": repr(ex)}
        if rank == 0:
            line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
