#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2_pytest28.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r2_pytest28.log
timeout 600 python bench.py > gpurun_out/r2_bench28_default.json 2> gpurun_out/r2_bench28_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench28_default.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms over %s steps) frac %.4f kernel ms %.3f launches %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("steps"), d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["gpu_launches"]))
print("parity", d["parity"].get("chi_within_reference_self_spread"), d["parity"]["vs_oracle"]["max_rel_dchi"])
for k,w in d.get("other_workloads",{}).items():
    print("   ", k, ("value %.4g e2e %.4g (%s ms)" % (w["value"], w["e2e"]["value"], w["e2e"].get("ms_per_step"))) if "value" in w else w, (w.get("parity") or {}).get("chi_within_reference_self_spread"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d["clocks"])
PY
tail -2 gpurun_out/r2_bench28_default.err
