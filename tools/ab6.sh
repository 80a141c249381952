#!/bin/bash
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 100 python tools/probe_batch.py 4096 1 0 | tail -1
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1
  timeout 100 python tools/probe_tl.py c2 1 | head -1
  timeout 100 python tools/probe_tl.py c5 1 | head -1
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
cp build/libS.so correlation_b200/libdic_b200.so
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench15_c3.json 2> gpurun_out/r2_bench15_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench15_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['unit'], d.get('e2e',{}).get('value'))"
timeout 300 python bench.py --no-cpu-baseline --no-others --parity-sample 4 > gpurun_out/r2_bench15_c4.json 2> gpurun_out/r2_bench15_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench15_c4.json').read().strip().splitlines()[-1]); print('c4 e2e ms', d['e2e']['ms_per_step'], 'value ms', d['ms_per_step'])"
