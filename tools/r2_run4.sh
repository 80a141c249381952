#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/chi_diag.py 128 > gpurun_out/r2_chi_diag2.log 2>&1
timeout 300 python tools/probe_tl.py c2 0 > gpurun_out/r2_tl_c2.log 2>&1
timeout 300 python tools/probe_tl.py c5 0 > gpurun_out/r2_tl_c5.log 2>&1
timeout 600 python -m pytest tests/test_gpu_host.py tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest4.log 2>&1
tail -12 gpurun_out/r2_chi_diag2.log; tail -30 gpurun_out/r2_tl_c2.log; grep -E "passed|failed|FAILED|assert" gpurun_out/r2_pytest4.log | tail
