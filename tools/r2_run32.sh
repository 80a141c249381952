#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N > gpurun_out/r2_bench32_n$N.json 2> gpurun_out/r2_bench32_n$N.err
echo "wall $(( $(date +%s) - t0 )) s"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2_bench32_n{n}.json").read().strip().splitlines()[-1])
print("N=%s value %.1f G (%.3f ms) e2e %.1f G (%.3f ms) kernel %.3f" % (n, d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"]), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"), d["parity"].get("chi_within_reference_self_spread"))
c5=d.get("other_workloads",{}).get("c5")
if c5: print("c5 value %.1f G e2e %.1f G" % (c5["value"]/1e9, c5["e2e"]["value"]/1e9), c5["parity"]["vs_single_gpu"].get("bitwise_equal"), c5["parity"].get("chi_within_reference_self_spread"))
PY
tail -2 gpurun_out/r2_bench32_n$N.err | cut -c1-300
