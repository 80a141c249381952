#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest10_$i.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi" gpurun_out/r2_pytest10_$i.log | tail -8
done
timeout 600 python bench.py > gpurun_out/r2_bench10_default.json 2> gpurun_out/r2_bench10_default.err
cat gpurun_out/r2_bench10_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench10_ref.json 2> gpurun_out/r2_bench10_ref.err
cat gpurun_out/r2_bench10_ref.json
