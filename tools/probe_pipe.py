#!/usr/bin/env python
"""GPU probe: device-time marks of the c4 staged-pair pipeline (dic_pipe_trace), both loop orders, both batch forms."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, bench
def loop_new(run, n):
    eng = run.eng
    run.stage_pair()
    for k in range(n):
        eng.advancePair()
        run.guess_buf[:] = 0.0
        eng.lib.dic_correlate_batch_async(eng.h, 0, run.n_sectors, run.guess_buf.ctypes.data)
        if k + 1 < n: run.stage_pair()
        eng.lib.dic_correlate_batch_wait(eng.h, 0, run.n_sectors, None, run.res_buf.ctypes.data)
def loop_deep(run, n):
    eng = run.eng
    run.stage_pair(); run.stage_pair()
    for k in range(n):
        eng.advancePair()
        run.guess_buf[:] = 0.0
        eng.lib.dic_correlate_batch_async(eng.h, 0, run.n_sectors, run.guess_buf.ctypes.data)
        if k + 2 < n: run.stage_pair()
        eng.lib.dic_correlate_batch_wait(eng.h, 0, run.n_sectors, None, run.res_buf.ctypes.data)
args = argparse.Namespace(mode=(sys.argv[1] if len(sys.argv) > 1 else "parity"), gpus=1)
run = bench.Run(args, bench.workload(sys.argv[2] if len(sys.argv) > 2 else "c4"), None, 0, 1, 0)
for q in (1, 2):
    run.eng.set_batch_queue(q)
    for name, loop in (("solve-then-stage", lambda n: loop_new(run, n)), ("two pairs ahead", lambda n: loop_deep(run, n))):
        loop(3)
        run.eng.pipe_trace(True, read=False)
        t0 = time.perf_counter(); loop(8); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 8 * 1e3
        st, so = run.eng.pipe_trace(False)
        print(f"queue {q} {name}: {wall:.3f} ms per step (host clock)")
        print("   k | copy may start  und landed  def landed  pyramids built | solve start  solve end (len) | period")
        for k in range(len(so)):
            s = st[k] if k < len(st) else [float('nan')] * 4
            per = so[k][1] - so[k - 1][1] if k else float('nan')
            print(f"  {k:2d} | {s[0]:9.3f} {s[1]:11.3f} {s[2]:11.3f} {s[3]:11.3f}      | {so[k][0]:9.3f} {so[k][1]:9.3f} ({so[k][1]-so[k][0]:.3f}) | {per:.3f}")
run.close()
