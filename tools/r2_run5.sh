#!/bin/bash
mkdir -p gpurun_out
for v in nt256 nt128; do
  DIC_B200_LIB=$PWD/build/ab/libdic_$v.so timeout 300 python bench.py --steps 10 --no-others --no-cpu-baseline --parity-sample 8 > gpurun_out/r2_ab_${v}_c4.json 2> gpurun_out/r2_ab_${v}_c4.err
done
DIC_B200_LIB=$PWD/build/ab/libdic_nt128.so timeout 300 python tools/probe_tl.py c5 0 > gpurun_out/r2_tl_c5b.log 2>&1
DIC_B200_LIB=$PWD/build/ab/libdic_nt128.so timeout 300 python tools/probe_tl.py c2 0 > gpurun_out/r2_tl_c2b.log 2>&1
DIC_B200_LIB=$PWD/build/ab/libdic_nt128.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "cta_pair or rect_grid or batch or c4" > gpurun_out/r2_pytest5.log 2>&1
for v in nt256 nt128; do python - <<PY
import json
l=json.loads(open('gpurun_out/r2_ab_${v}_c4.json').read().strip().split('\n')[-1])
print('$v', 'value %.2f G'%(l['value']/1e9), 'kernel ms', round(l['roofline']['kernel_ms_per_step'],4), 'e2e %.2f G'%(l['e2e']['value']/1e9), round(l['e2e']['ms_per_step'],3), 'fast %.1f G'%(l['other_arith_mode']['kernel_value_this_rank']/1e9))
PY
done
tail -12 gpurun_out/r2_tl_c5b.log; grep -E "passed|failed|FAILED|c4 chi" gpurun_out/r2_pytest5.log | tail
