#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest9_$i.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi" gpurun_out/r2_pytest9_$i.log | tail -8
done
