#!/bin/bash
mkdir -p gpurun_out/prof
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest12.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi|c4 subsets|c2 mode" gpurun_out/r2_pytest12.log | tail -12
for m in 0 1; do
ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 2 -c 1 -f -o gpurun_out/prof/j2_mode$m python tools/probe_batch.py 4096 $m 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof/j2_mode$m.ncu-rep 12 > gpurun_out/prof/j2_mode${m}_ncu_full.txt 2>&1
head -34 gpurun_out/prof/j2_mode${m}_ncu_full.txt
done
