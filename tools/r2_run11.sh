#!/bin/bash
mkdir -p gpurun_out
./build/ubench > gpurun_out/r2_ubench.log 2>&1; cat gpurun_out/r2_ubench.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest11.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi|c4 subsets" gpurun_out/r2_pytest11.log | tail -8
