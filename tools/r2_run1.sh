#!/bin/bash
# round-2 GPU pass 1: tests, default bench, other workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rA 2>&1 | tail -80 > gpurun_out/r2_pytest1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
timeout 300 python bench.py --workload c2 --steps 10 --no-others > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --workload c5 --steps 5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
timeout 300 python bench.py --workload c3 --steps 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
timeout 300 python bench.py --workload c1 --steps 10 --no-others > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
tail -5 gpurun_out/r2_pytest1.log
head -c 1500 gpurun_out/r2_bench_c4.json
