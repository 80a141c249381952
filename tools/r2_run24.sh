#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-others > gpurun_out/r2_bench24_n1.json 2> gpurun_out/r2_bench24_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 > gpurun_out/r2_bench24_n4.json 2> gpurun_out/r2_bench24_n4.err
python - <<'PY'
import json
v={}
for n in (1,4):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench24_n{n}.json").read().strip().splitlines()[-1]); v[n]=d
        print(n, "value %.1f G ms %.3f e2e %.1f G (%.3f ms) kernel %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"]), d["parity"].get("vs_single_gpu",{}).get("bitwise_equal"))
        for k,w in d.get("other_workloads",{}).items(): print("   ", k, "value %.1f G e2e %.1f G" % (w["value"]/1e9, w["e2e"]["value"]/1e9) if "value" in w else w, (w.get("parity") or {}).get("chi_within_reference_self_spread"))
    except Exception as ex: print(n, "failed", ex)
if 1 in v and 4 in v: print("scaling x%.2f  e2e x%.2f" % (v[4]["value"]/v[1]["value"], v[4]["e2e"]["value"]/v[1]["e2e"]["value"]))
PY
