#!/bin/bash
# N-GPU pass: $1 = number of GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r2_multi_check_$N.log 2>&1
tail -5 gpurun_out/r2_multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().split('\n')[-1])
    print('N=$N value %.2f G  ms %.3f  e2e %.2f G (%.3f ms)  frac %.3f'%(l['value']/1e9, l['ms_per_step'], l['e2e']['value']/1e9, l['e2e']['ms_per_step'], l['roofline']['frac']))
    print(json.dumps(l['parity'])[:1800])
    for k,v in l.get('other_workloads',{}).items():
        if 'value' in v: print(k,'value %.2f G e2e %.2f G'%(v['value']/1e9, v['e2e']['value']/1e9), json.dumps(v.get('parity'))[:1500])
        else: print(k, v)
except Exception as ex: print('ERR', ex)
PY
