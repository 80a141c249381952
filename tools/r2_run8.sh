#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest8.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest8.log | tail -15
python tools/probe_domain.py 2>&1 | tail -3
