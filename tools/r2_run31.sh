#!/bin/bash
mkdir -p gpurun_out
for v in prio noprio prio noprio; do
  if [ $v = noprio ]; then export DIC_NO_STREAM_PRIORITY=1; else unset DIC_NO_STREAM_PRIORITY; fi
  DIC_HOST_PROFILE=1 timeout 300 python bench.py --workload c3 --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/r2_c3_$v.json 2> gpurun_out/r2_c3_$v.err
  echo "$v: sectors phase (ms) per run:" $(grep "dic_host" gpurun_out/r2_c3_$v.err | sed 's/.*sectors \([0-9.]*\),.*/\1/' | tr '\n' ' ')
done
