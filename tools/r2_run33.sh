#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2_pytest33.log 2>&1
echo "pytest rc $?"; tail -2 gpurun_out/r2_pytest33.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 400 python bench.py > gpurun_out/r2_bench33_default.json 2> gpurun_out/r2_bench33_default.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench33_ref.json 2> gpurun_out/r2_bench33_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench33_default.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench33_ref.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms over %s steps) frac %.4f kernel ms %.3f traffic %s launches %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("steps"), d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["traffic"], d["gpu_launches"]))
print("fast", d["other_arith_mode"]["kernel_value_this_rank"]/1e9, "ref arm %.2f M -> e2e ratio %.0f" % (r["value"]/1e6, d["e2e"]["value"]/r["value"]))
print("parity", d["parity"].get("chi_within_reference_self_spread"), d["parity"]["vs_oracle"]["max_rel_dchi"], d["parity"]["vs_oracle"]["max_abs_duv"])
for k,w in d.get("other_workloads",{}).items():
    print("   ", k, ("value %.4g e2e %.4g (%s ms)" % (w["value"], w["e2e"]["value"], w["e2e"].get("ms_per_step"))) if "value" in w else w, (w.get("parity") or {}).get("chi_within_reference_self_spread"), w["config"].get("ms_per_frame_spread"))
print("clocks", d["clocks"])
PY
