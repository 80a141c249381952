#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest13.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi|c4 subsets|c2 mode" gpurun_out/r2_pytest13.log | tail -12
timeout 600 python bench.py > gpurun_out/r2_bench13_default.json 2> gpurun_out/r2_bench13_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench13_default.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms) frac %.4f kernel ms %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"]))
print("parity", {k:v for k,v in d["parity"]["vs_oracle"].items() if k.startswith("max") or k.startswith("within") or k.startswith("units")})
c2=d["other_workloads"]["c2"]
print("c2 value %.1f G ms %.3f e2e %.1f G frac %.4f" % (c2["value"]/1e9, c2["ms_per_step"], c2["e2e"]["value"]/1e9, c2["roofline"]["frac"]), c2["parity"].get("chi_within_reference_self_spread"), c2["parity"]["vs_oracle"]["max_rel_dchi"], c2["parity"].get("reference_self_spread",{}).get("max_rel_dchi"))
print("cpu", d["cpu_baseline"]["value"]/1e6, "M")
PY
