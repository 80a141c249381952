#!/bin/bash
# tools/rep.sh <workload> [n]: repeat one bench workload on the same box and print the spread
w=${1:-c3}; n=${2:-3}
for i in $(seq $n); do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['config']['arith_mode'], 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'])"
done
