#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --no-cpu-baseline --no-others > gpurun_out/r2_bench30_n2.json 2> gpurun_out/r2_bench30_n2.err
echo "wall $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench30_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1f G (%.3f ms) e2e %.1f G (%.3f ms)" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"), d["parity"].get("chi_within_reference_self_spread"))
PY
tail -2 gpurun_out/r2_bench30_n2.err
