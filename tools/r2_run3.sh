#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/lm_trace.py 3250 980 455 2210 > gpurun_out/r2_lm_trace.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/r2_pytest3.log 2>&1
for v in loop1 loop2; do
  for wl in c4 c2; do
    DIC_B200_LIB=$PWD/build/ab/libdic_$v.so timeout 300 python bench.py --workload $wl --steps 10 --no-others --no-cpu-baseline > gpurun_out/r2_ab_${v}_$wl.json 2> gpurun_out/r2_ab_${v}_$wl.err
  done
done
tail -40 gpurun_out/r2_lm_trace.log
grep -E "passed|failed" gpurun_out/r2_pytest3.log | tail -3
