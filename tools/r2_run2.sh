#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "error_handling_modes or rowsplit_loopback_5 or 256_subsets" 2>&1 | tail -150 > gpurun_out/r2_pytest2.log
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "error_handling_modes or rowsplit_loopback_5 or 256_subsets" 2>&1 | grep -E "^(FAILED|PASSED|ERROR)|assert|Error|c4 parity" | head -60 >> gpurun_out/r2_pytest2.log
timeout 600 python tools/chi_diag.py 64 > gpurun_out/r2_chi_diag.log 2>&1
tail -30 gpurun_out/r2_chi_diag.log
