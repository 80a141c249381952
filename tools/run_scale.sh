#!/bin/bash
# tools/run_scale.sh N [workloads...]: bench.py on N GPUs of this box, one JSON line per workload into gpurun_out/
N=$1; shift
mkdir -p gpurun_out
for w in "$@"; do
  port=$((29500 + RANDOM % 400))
  if [ "$N" -gt 1 ]; then
    timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench${N}_$w.err | tail -1 > gpurun_out/bench${N}_$w.json
  else
    timeout 280 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench${N}_$w.err | tail -1 > gpurun_out/bench${N}_$w.json
  fi
  python - "$w" "$N" <<'PY'
import json, sys
w, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/bench{n}_{w}.json").read().strip().splitlines()[-1])
    print(f"{w} x{n}: host {d.get('host_ms_per_step', 0):.3f} ms  value {d['value']/1e9:.1f} G  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']/1e9:.1f} G ({d['e2e'].get('ms_per_step', 0):.3f} ms, h2d {d['e2e']['h2d_bytes_per_step']/1e6:.1f} MB)  kernel {d['roofline']['kernel_ms_per_step']:.3f} ms  clocks {d['clocks'].get('sm_mhz')}")
except Exception as ex:
    print(w, n, "failed", ex)
PY
done
