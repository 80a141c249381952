#!/bin/bash
mkdir -p gpurun_out
bash tools/r2_run6.sh 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-others > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-others --no-cpu-baseline --parity-sample 8 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err
python - <<PY
import json
for n in ('4','1b'):
    try:
        l=json.loads(open('gpurun_out/r2_bench_n%s.json'%n).read().strip().split('\n')[-1])
        print('N=%s value %.2f G  ms %.3f  e2e %.2f G (%.3f ms)'%(n, l['value']/1e9, l['ms_per_step'], l['e2e']['value']/1e9, l['e2e']['ms_per_step']), l['parity'].get('vs_single_gpu',{}).get('bitwise_equal'))
    except Exception as ex: print('ERR', n, ex)
PY
