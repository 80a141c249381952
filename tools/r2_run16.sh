#!/bin/bash
mkdir -p gpurun_out
python tools/probe_pair_check.py | head -1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest16.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|c4 chi|c4 subsets|c2 mode" gpurun_out/r2_pytest16.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench16_default.json 2> gpurun_out/r2_bench16_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench16_ref.json 2> gpurun_out/r2_bench16_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench16_default.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench16_ref.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms) frac %.4f kernel ms %.3f traffic %s launches %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["traffic"], d["gpu_launches"]))
print("fast", d["other_arith_mode"])
print("ref arm %.2f M -> e2e ratio %.0f" % (r["value"]/1e6, d["e2e"]["value"]/r["value"]))
c2=d["other_workloads"]["c2"]
print("c2 value %.1f G ms %.3f e2e %.1f G frac %.4f" % (c2["value"]/1e9, c2["ms_per_step"], c2["e2e"]["value"]/1e9, c2["roofline"]["frac"]), c2["parity"].get("chi_within_reference_self_spread"), c2["parity"]["vs_oracle"]["max_rel_dchi"], c2["other_arith_mode"])
print("clocks", d["clocks"])
PY
