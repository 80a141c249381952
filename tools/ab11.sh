#!/bin/bash
for lib in build/libdef.so build/libBIG.so; do
  cp $lib correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload c5 --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 x2 value %.1f G ms %.3f kernel %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms_per_step']))"
  timeout 100 python tools/probe_tl_c3.py 0 2>/dev/null | head -1
  timeout 100 python tools/probe_tl.py c5 0 2>/dev/null | head -1
done
cp build/libdef.so correlation_b200/libdic_b200.so
