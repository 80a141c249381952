#!/usr/bin/env python
"""bench.py -- domain pixel*GN-evaluations / s of the DIC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c4] [--impl ours|reference]

A *step* is one correlate() of the workload's domain from the zero initial guess: every pyramid
level, every LM iteration, reduction and solve (correlation_class.cpp:349-640 semantics), on image
pyramids already resident in HBM.  Work per step = sum over levels of N_level x evaluations_level
(SURVEY.md section 8d), reported by the engine itself.

  value     pixel*evaluations / s, device-resident inputs, all ranks (weak scaling: every rank
            solves its own independent domain, no data-path collective)
  e2e       same metric through the C-ABI with HOST (pinned) images: per step H2D of the image
            pair, both pyramid builds, correlate, D2H of the result record
  roofline  the fused GN kernel against the measured HBM copy bandwidth, with SURVEY 8d's
            10 algorithmic bytes per pixel*evaluation
  cpu_baseline / --impl reference
            the reference CPU algorithm on this box's host cores: oracle/_ref (unmodified
            reference, 6-parameter workloads) or the oracle port (12-parameter workload, which the
            reference does not have), all host threads.
"""
from __future__ import annotations

import argparse
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


# ------------------------------------------------------------------------------ workloads

def workload(name):
    two_pi = 2.0 * np.pi
    if name == "c1":  # BASELINE config 1 (the reference's CPU-runnable case)
        return dict(name="c1: 1024^2 pair, 511x511 rect, affine 6-param, bicubic, pyramid 0/1/2",
                    rows=1024, cols=1024, seed=1, model="affine", pyramid=(0, 1, 2),
                    truth=(1.75, -0.6, .004, -.003, .002, .005), center=(512.0, 512.0),
                    domain=("rect", 257, 257, 767, 767))
    if name == "c2":  # BASELINE config 2: the configuration the metric is quoted on
        return dict(name="c2: 4096^2 pair, annulus ri=600 ro=1800, quadratic 12-param, bicubic, pyramid 0..3",
                    rows=4096, cols=4096, seed=2, model="quad", pyramid=(0, 1, 3),
                    truth=(2.5, -1.75, .002, -.0015, .001, .0025, 1e-6, -5e-7, 8e-7, -1e-6, 6e-7, 4e-7),
                    center=(2048.0, 2048.0), domain=("annulus", 600.0, 1200.0, 0.0, two_pi, 2048.0, 2048.0, 1))
    if name == "c4":  # BASELINE config 4: 4096 subsets of 125^2
        return dict(name="c4: 8192^2 pair, 64x64 subsets of 125^2, affine, pyramid 0/1/2",
                    rows=8192, cols=8192, seed=4, model="affine", pyramid=(0, 1, 2),
                    truth=(1.25, -0.75, .0004, -.0003, .0002, .0005), center=(4096.0, 4096.0),
                    domain=("subsets", 64, 8128, 64))
    if name == "c3":  # BASELINE config 3: frames/s of a 100-frame sequence
        return dict(name="c3: 100-frame 2048^2 sequence, 64-vertex star blob (mean radius 700), affine, pyramid 0/1/2, "
                         "Eulerian + first-image reference, constant-velocity initial guess",
                    rows=2048, cols=2048, seed=3, model="affine", pyramid=(0, 1, 2), frames=100,
                    rate=(0.8, -0.5, 0.0, -0.0005, 0.0005, 0.0), center=(1024.0, 1024.0),
                    truth=(0.8, -0.5, 0.0, -0.0005, 0.0005, 0.0), domain=("blob", 1024.0, 1024.0, 700.0, 64))
    if name == "c5":  # BASELINE config 5: one huge domain, row-split across GPUs
        return dict(name="c5: 16384^2 pair, one 15361^2 rect domain, affine, pyramid 0..4, row-split + in-kernel all-reduce",
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
        self.period = 0.02 * max(1, world)
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

def cpu_run(w, und, dfm, threads, sample_levels=None):
    """One 'step' of the reference CPU algorithm: both pyramid builds + Newton_Raphson (cold cache).
    Returns (pixel_evaluations, seconds, kind, result)."""
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
        boxes = subset_boxes(d[1], d[2], d[3])[:: max(1, (d[3] * d[3]) // 64)]
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
        sample = f"{len(boxes)} of {d[3] * d[3]} subsets, one cold step"
    return work, secs, ("reference" if use_ref else "port"), res, sample, t_pyr


# ------------------------------------------------------------------------------ config 3: frames / s

def run_c3(args, w):
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
    secs = sum(r["seconds"] for r in runs) / len(runs)
    fps = (n - 1) / secs
    hdr, rows = host.parse_report(runs[-1]["csv"])
    last = runs[-1]["rows"][0]
    truth_last = (n - 1) * np.array(w["rate"])
    line = {"metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic speckle sequence, constant velocity",
            "config": {"workload": w["name"], "arith_mode": args.mode, "frame_pairs": n - 1,
                       "points": int(last["number_of_points"]), "errors": int(sum(int(r["error_code"]) != 0 for r in rows)),
                       "last_frame_params": [float(v) for v in last["params"][:6]],
                       "last_frame_truth": [float(v) for v in truth_last],
                       "step": "one step = the whole 99-pair sequence through dic_host_run (C++ host loop)"},
            "clocks": clk.summary(),
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": n * w["rows"] * w["cols"],
                    "d2h_bytes_per_step": 176 * (n - 1)},
            "gpu_launches": (n - 1) * 5,
            "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "frames/s is a pipeline metric (upload + pyramid + GN per frame); see the c2 / c4 lines for the kernel roofline"}}
    if not args.no_cpu_baseline:
        import oracle
        t0 = time.perf_counter()
        O = oracle.OracleEngine(n_threads=os.cpu_count() or 1, pyramid=w["pyramid"], real_threads=True)
        O.set_image("und", frames[0])
        xy = oracle.blob_points(contour)
        p = np.zeros(6, np.float32)
        p_prev = p.copy()
        nf = 4
        for k in range(nf):
            O.set_image("def", frames[k + 1])
            guess = p + (p - p_prev) if k else p
            p_prev = p
            p = O.correlate(guess, xy)["params"]
        cs = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": nf / cs, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"first {nf} frame pairs of the sequence (pyramid + Newton_Raphson each)"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="parity", choices=["parity", "fast"],
                    help="arithmetic form of the per-pixel evaluation (include/dic_b200.h dic_arith_mode). 'parity' (default, "
                         "the headline) replays the reference's fp32 operation order: w, dw/dx, dw/dy bit-identical per pixel, "
                         "every BASELINE.json tolerance met. 'fast' is the same interpolant in Catmull-Rom form (~2x fewer "
                         "instructions, more accurate than the reference's own rounding noise, which is why its chi can sit "
                         "1e-5 away from the reference's); it is reported beside the headline, labelled")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = workload(args.workload)
    n_par = 12 if w["model"] == "quad" else 6
    if args.workload == "c3" and args.impl == "ours":
        if rank != 0:
            return 0
        return run_c3(args, w)

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
        und_t, dfm_t = make_images(w, dev) if dev is not None else (None, None)
        if dev is None:
            from correlation_b200 import synth
            und = synth.make_image(w["rows"], w["cols"], w["seed"], None, w["center"])
            dfm = synth.make_image(w["rows"], w["cols"], w["seed"], w["truth"], w["center"])
        else:
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
                "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic speckle (SURVEY 8d)",
                "config": {"workload": w["name"]},
                "cpu_baseline": {"value": v, "unit": "pixel*evaluations/s", "cores": threads, "kind": kind,
                                 "sample": sample},
                "e2e": {"value": v, "unit": "pixel*evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------------------------------------------------------- our arm
    from correlation_b200 import engine
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    und_t, dfm_t = make_images(w, dev)
    und_pin = torch.empty(und_t.shape, dtype=torch.uint8, pin_memory=True)
    dfm_pin = torch.empty(dfm_t.shape, dtype=torch.uint8, pin_memory=True)
    und_pin.copy_(und_t)
    dfm_pin.copy_(dfm_t)
    torch.cuda.synchronize()
    mode = engine.MODE_PARITY if args.mode == "parity" else engine.MODE_FAST
    eng = engine.CudaEngine(local_rank, fitting_model=engine.FM_QUADRATIC if w["model"] == "quad" else engine.FM_UVUxUyVxVy,
                            arith_mode=mode)
    rows, cols = w["rows"], w["cols"]
    eng.resetImagePyramidsDevice(und_t.data_ptr(), dfm_t.data_ptr(), None, rows, cols, cols, pyramid=w["pyramid"])
    d = w["domain"]
    scaling = "weak"
    t_dom = time.perf_counter()
    if d[0] == "rect":
        eng.resetPolygon(0, *d[1:])
        n_sectors = 1
    elif d[0] == "annulus":
        eng.resetPolygon(0, *d[1:])
        n_sectors = 1
    elif d[0] == "rowsplit":
        # one domain, pixel rows in equal bands per rank, per-evaluation sum inside the kernel
        from correlation_b200 import rowsplit
        rowsplit.connect(eng, dist)
        b0, b1 = rowsplit.equal_row_bands(d[2], d[4], world)[rank]
        eng.resetPolygonRectBand(0, d[1], d[2], d[3], d[4], b0, b1)
        n_sectors = 1
        scaling = "strong"
    else:
        # independent subsets shard across ranks in contiguous blocks, images replicated
        from correlation_b200 import sharding
        boxes = subset_boxes(d[1], d[2], d[3])
        # whole rows of subsets per rank: each rank then needs only a band of image rows (e2e upload)
        boxes = [boxes[i] for i in sharding.shard_grid_rows(d[3], d[3], world, rank)]
        for k, bx in enumerate(boxes):
            eng.resetPolygon(k, *bx)
        n_sectors = len(boxes)
        scaling = "strong"
    eng.synchronize()
    t_dom = time.perf_counter() - t_dom
    zero = np.zeros((n_sectors, n_par), np.float32)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    res_buf = np.zeros(n_sectors, engine.RESULT_DTYPE)
    guess_buf = np.zeros((n_sectors, n_par), np.float32)

    one_guess = np.zeros(n_par, np.float32)
    one_result = engine.DicResult()

    def step_resident():
        if n_sectors == 1:
            one_guess[:] = 0.0
            eng.correlate_raw(0, one_guess, one_result)
            work = 0.0
            for lv in range(engine.MAX_LEVELS):
                work += float(one_result.evaluationsPerLevel[lv]) * float(one_result.pointsPerLevel[lv])
            return work, eng.last_correlate_ms(), one_result
        guess_buf[:] = 0.0
        eng.lib.dic_correlate_batch(eng.h, 0, n_sectors, guess_buf.ctypes.data, res_buf.ctypes.data)
        return None, eng.last_correlate_ms(), res_buf  # work is read from the records after the loop

    # a rank that owns a band of the image (sharded subsets, row-split domain) transfers only its rows plus
    # a halo for displacement, bicubic support and pyramid support (dic_stage_next_pair_rows)
    band = None
    if world > 1 and d[0] == "rowsplit":
        band = upload_band(b0, b1, rows, w["pyramid"][2])
    elif world > 1 and d[0] == "subsets":
        band = upload_band(min(bx[1] for bx in boxes), max(bx[3] for bx in boxes), rows, w["pyramid"][2])
    h2d_bytes = 2 * cols * ((band[1] - band[0]) if band else rows)

    def stage_pair():
        # this step's inputs: both images from pinned host memory, upload + pyramids on the image stream
        eng.stageNextPair(und_pin.data_ptr(), dfm_pin.data_ptr(), rows, cols, row_range=band)

    def e2e_loop(n):
        # double-buffered ingest (dic_stage_next_pair / dic_advance_pair): the PCIe transfer of pair k + 1
        # overlaps the solve of pair k; every step copies its own pair and reads its own result record
        tot = 0.0
        stage_pair()
        for k in range(n):
            eng.advancePair()
            if k + 1 < n:
                stage_pair()
            wk, _, _ = step_resident()
            tot += batch_work if wk is None else wk
        return tot

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    launches0 = eng.kernel_launches()
    barrier()
    # Timed on the DEVICE: CUDA events on the correlation stream around the whole GPU side of each step (guess
    # upload, solve, result download), summed over the K steps, max over ranks. The L2 flush between steps is
    # outside the events. The host clock around the same calls is reported beside it (host_ms_per_step): it
    # adds launch latency and the wake-up after the sync, and on a busy 8-rank box it jitters by 0.1 ms.
    work = kern_ms = wall = host_wall = 0.0
    with ClockSampler(local_rank, world) as clk:
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            wk, ms, last = step_resident()
            host_wall += time.perf_counter() - t0
            wall += 1e-3 * eng.last_step_ms()
            if wk is None:  # batch: identical inputs every step, count the work once per step from the records
                wk = eng.pixel_evaluations(last)
            work += wk
            kern_ms += ms
    barrier()
    launches = eng.kernel_launches() - launches0
    if n_sectors == 1:
        last = last.as_dict(n_par)
    else:
        rs = last
        last = dict(params=rs[0]["resultingParameters"][:n_par].copy(), chi=rs[0]["chi"],
                    iterations=int(rs[0]["iterations"]), evaluations=rs[0]["evaluationsPerLevel"].tolist(),
                    points_per_level=rs[0]["pointsPerLevel"].tolist(), errors=int((rs["errorCode"] != 0).sum()))
    batch_work = eng.pixel_evaluations(res_buf) if n_sectors > 1 else None
    # e2e: host buffers, copies inside the timed region
    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_work = e2e_loop(args.steps)
    barrier()
    e2e_wall = time.perf_counter() - t0

    # the other arithmetic mode, same resident inputs, for the record
    other = engine.MODE_PARITY if mode == engine.MODE_FAST else engine.MODE_FAST
    eng.set_arith_mode(other)
    for _ in range(2):
        step_resident()
    o_work = o_ms = 0.0
    for _ in range(max(2, args.steps // 2)):
        flush.fill_(1)
        torch.cuda.synchronize()
        wk, ms, _o = step_resident()
        o_work += eng.pixel_evaluations(_o) if wk is None else wk
        o_ms += ms
    eng.set_arith_mode(mode)

    stats = torch.tensor([wall, e2e_wall, work, e2e_work, kern_ms, host_wall], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, e2e_wall, host_wall = mx[0].item(), mx[1].item(), mx[5].item()
        work, e2e_work = sm[2].item(), sm[3].item()
        if d[0] == "rowsplit":  # every rank's result record already counts the whole domain
            work, e2e_work = mx[2].item(), mx[3].item()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:  # DRAM bytes of one launch of the dominant kernel, from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["gn_solve_tiles_kernel"].get(f"{args.workload} {args.mode}")
        if tr:
            traffic = tr["dram_bytes_read_per_launch"] + tr["dram_bytes_write_per_launch"]
    except Exception:
        pass
    my_work = stats[2].item()
    achieved = ALGO_BYTES_PER_PIXEL_EVAL * my_work / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    value = work / wall
    line = {
        "metric": "domain pixel*GN-evaluations/s", "value": value, "unit": "pixel*evaluations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "host_ms_per_step": 1e3 * host_wall / args.steps,
        "timing": "CUDA events on the correlation stream around each step's device work, summed over the steps, max over ranks",
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic analytic speckle (SURVEY 8d), random phases seeded",
        "config": {"workload": w["name"], "arith_mode": args.mode, "sectors": n_sectors,
                   "pixel_evaluations_per_step": my_work / args.steps,
                   "evaluations_per_level": last["evaluations"][: w["pyramid"][2] + 1],
                   "points_per_level": last["points_per_level"][: w["pyramid"][2] + 1],
                   "l2": "flushed between steps (512 MiB write); evaluations inside a step re-read the domain by design",
                   "domain_build_s": t_dom,
                   "parallelism": (f"one domain in {world} row band(s), per-evaluation all-reduce of the normal equations inside the kernel (NVLink peer mailboxes)"
                                   if d[0] == "rowsplit" else
                                   f"{d[3] * d[3]} subsets, whole rows of subsets per GPU, over {world} GPU(s), no collective"
                                   if scaling == "strong" else f"{world} independent domain(s), one per GPU")},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_work / e2e_wall, "unit": "pixel*evaluations/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 176 * n_sectors,
                "ms_per_step": 1e3 * e2e_wall / args.steps,
                "pipeline": "dic_stage_next_pair(k + 1) on the copy / image streams overlaps dic_correlate(k); H2D of both images every step"
                            + ("" if band is None else f" (this rank's row band {band[0]}..{band[1]} of {rows}; bytes are per rank)")},
        "gpu_launches": launches,
        "other_arith_mode": {"arith_mode": "parity" if other == engine.MODE_PARITY else "fast",
                             "kernel_value_this_rank": o_work / (o_ms * 1e-3) if o_ms > 0 else None,
                             "unit": "pixel*evaluations/s (kernel time, one rank)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_note": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_traffic.json); "
                                     "algorithmic bytes per launch = 10 B x pixel_evaluations_per_step",
                     "kernel": "gn_solve_tiles_kernel" if n_sectors >= 1 else "gn_solve_kernel",
                     "kernel_ms_per_step": kern_ms / args.steps,
                     "fp32_note": "the kernel is FP32-issue bound, not HBM bound (DESIGN.md 4.1): parity mode replays the reference's "
                                  "~270 unfused fp32 operations per pixel (369 issued instructions per pixel*evaluation over the whole "
                                  "launch, FMA pipe 47 % of peak, issue slots 62 % incl. the per-evaluation grid all-reduce); the level "
                                  "data stays L2-resident across evaluations (31 MB of DRAM reads for 369 MB of algorithmic traffic). "
                                  "profiles/r1_c2_gn_solve_tiles_parity_ncu_full.txt, profiles/r1_l0_pass_ncu_full.txt",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                     "algorithmic_bytes_per_pixel_evaluation": ALGO_BYTES_PER_PIXEL_EVAL},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            und, dfm = und_pin.numpy(), dfm_pin.numpy()
            threads = os.cpu_count() or 1
            cw, cs, kind, cres, sample, t_pyr = cpu_run(w, und, dfm, threads)
            line["cpu_baseline"] = {"value": cw / cs, "unit": "pixel*evaluations/s", "cores": threads, "kind": kind,
                                    "sample": sample, "seconds": cs, "pyramid_seconds": t_pyr}
            gp, cp = last["params"], cres["params"]
            if d[0] in ("rect", "annulus") and w["rows"] <= 4096:
                # the gate of BASELINE.json: against the oracle with fp64 accumulators (the CPU engine's own fp32
                # accumulation moves chi by ~2e-4 with its thread count)
                import oracle
                od = oracle.OracleEngine(model=oracle.FM_QUAD if w["model"] == "quad" else oracle.FM_AFFINE, n_threads=threads,
                                         pyramid=w["pyramid"], accum_double=True, real_threads=True)
                od.set_image("und", und)
                od.set_image("def", dfm)
                if d[0] == "rect":
                    dres = od.correlate(np.zeros(n_par, np.float32), oracle.rect_points(*d[1:]), center=((d[1] + d[3]) / 2.0, (d[2] + d[4]) / 2.0))
                else:
                    dres = od.correlate(np.zeros(n_par, np.float32), oracle.annulus_points(*d[1:]))
                dp = dres["params"]
                line["config"]["parity_vs_oracle_fp64_accumulators"] = {
                    "max_abs_duv": float(np.abs(gp[:2] - dp[:2]).max()), "max_abs_dgrad": float(np.abs(gp[2:6] - dp[2:6]).max()),
                    "rel_dchi": float(abs(last["chi"] - dres["chi"]) / max(abs(dres["chi"]), 1e-30)),
                    "iterations": [int(last["iterations"]), int(dres["iterations"])],
                    "tolerances": {"duv": 1e-4, "dgrad": 1e-6, "rel_dchi": 1e-5, "iterations": 1}}
            if d[0] in ("rect", "annulus"):  # same domain on both sides
              line["config"]["parity_vs_cpu"] = {
                "max_abs_duv": float(np.abs(gp[:2] - cp[:2]).max()), "max_abs_dgrad": float(np.abs(gp[2:6] - cp[2:6]).max()),
                "rel_dchi": float(abs(last["chi"] - cres["chi"]) / max(abs(cres["chi"]), 1e-30)),
                "iterations": [int(last["iterations"]), int(cres["iterations"])],
                "note": "CPU side accumulates in fp32 per thread (its chi moves ~2e-4 with the thread count, SURVEY H1)"}
        except Exception as ex:  # the baseline is reported, never allowed to sink the bench line
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
