#!/usr/bin/env python
"""GPU probe: where the c4 end-to-end step goes. Pure H2D of the pair, stage_pair alone (H2D + pyramids), the e2e
loop with the batch as resident CTAs (queue 1) or one CTA per sector (queue 2), and with the staging call before or
after the solve has been enqueued."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, bench
def t(f, reps=10):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def loop_new(run, n):
    eng = run.eng
    run.stage_pair()
    for k in range(n):
        eng.advancePair()
        run.guess_buf[:] = 0.0
        eng.lib.dic_correlate_batch_async(eng.h, 0, run.n_sectors, run.guess_buf.ctypes.data)
        if k + 1 < n: run.stage_pair()
        eng.lib.dic_correlate_batch_wait(eng.h, 0, run.n_sectors, None, run.res_buf.ctypes.data)
for mode in (sys.argv[1:] or ["parity", "fast"]):
    args = argparse.Namespace(mode=mode, gpus=1)
    w = bench.workload("c4")
    run = bench.Run(args, w, None, 0, 1, 0)
    nbytes = run.h2d_bytes
    if mode == "parity":
        d1 = torch.empty_like(run.und_t); d2 = torch.empty_like(run.dfm_t)
        def raw():
            d1.copy_(run.und_pin, non_blocking=True); d2.copy_(run.dfm_pin, non_blocking=True)
        a = t(raw)
        print(f"raw H2D of the pair: {a:.3f} ms = {nbytes/a/1e6:.1f} GB/s")
    for q in (1, 2):
        run.eng.set_batch_queue(q)
        for _ in range(3): run.step_resident()
        ms = []
        for _ in range(10):
            run.flush.fill_(1); torch.cuda.synchronize()
            run.step_resident(); ms.append(run.eng.last_correlate_ms())
        r = float(np.mean(ms))
        ref = run.res_buf.copy()
        run.e2e_loop(3)
        torch.cuda.synchronize(); t0 = time.perf_counter(); run.e2e_loop(12); torch.cuda.synchronize()
        e = (time.perf_counter() - t0) / 12 * 1e3
        loop_new(run, 3)
        torch.cuda.synchronize(); t0 = time.perf_counter(); loop_new(run, 12); torch.cuda.synchronize()
        e2 = (time.perf_counter() - t0) / 12 * 1e3
        same = bool((run.res_buf.tobytes() == ref.tobytes()))
        print(f"{mode} queue {q}: kernel {r:.3f} ms (flushed L2), e2e stage-then-solve {e:.3f} ms, solve-then-stage {e2:.3f} ms "
              f"= {nbytes/e2/1e6:.1f} GB/s of H2D; records equal to the resident step's: {same}")
    run.close()

# Is the H2D transfer itself slower while the solve runs? Time the raw pair copy on a side stream with CUDA events,
# alone and with a resident solve enqueued just before it.
args = argparse.Namespace(mode="parity", gpus=1)
run = bench.Run(args, bench.workload("c4"), None, 0, 1, 0)
d1 = torch.empty_like(run.und_t); d2 = torch.empty_like(run.dfm_t)
side = torch.cuda.Stream()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for with_solve in (False, True, False, True):
    ms = []
    for _ in range(6):
        torch.cuda.synchronize()
        run.guess_buf[:] = 0.0
        if with_solve:
            run.eng.lib.dic_correlate_batch_async(run.eng.h, 0, run.n_sectors, run.guess_buf.ctypes.data)
        with torch.cuda.stream(side):
            ev0.record(); d1.copy_(run.und_pin, non_blocking=True); d2.copy_(run.dfm_pin, non_blocking=True); ev1.record()
        if with_solve:
            run.eng.lib.dic_correlate_batch_wait(run.eng.h, 0, run.n_sectors, None, run.res_buf.ctypes.data)
        torch.cuda.synchronize(); ms.append(ev0.elapsed_time(ev1))
    m = float(np.median(ms))
    print(f"raw H2D of the pair, solve running: {with_solve}: {m:.3f} ms = {run.h2d_bytes/m/1e6:.1f} GB/s")
run.close()
