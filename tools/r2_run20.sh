#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 > gpurun_out/r2_bench20_n2.json 2> gpurun_out/r2_bench20_n2.err
echo "wall $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench20_n2.json").read().strip().splitlines()[-1])
print("value %.1f G e2e %.1f G" % (d["value"]/1e9, d["e2e"]["value"]/1e9), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"), d["parity"].get("chi_within_reference_self_spread"))
c5=d["other_workloads"]["c5"]
print("c5 value %.1f G" % (c5["value"]/1e9), {k:(v.get("max_rel_dchi"), v.get("within_tolerance")) if isinstance(v,dict) else v for k,v in c5["parity"].items()})
PY
tail -2 gpurun_out/r2_bench20_n2.err
