#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest18.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest18.log | tail -5
timeout 600 python bench.py --no-cpu-baseline --parity-sample 8 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench18.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms) kernel ms %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_step"]), d["clocks"])
c2=d["other_workloads"]["c2"]
print("c2 value %.1f G ms %.3f e2e %.1f G kernel %.3f" % (c2["value"]/1e9, c2["ms_per_step"], c2["e2e"]["value"]/1e9, c2["roofline"]["kernel_ms_per_step"]))
PY
python tools/probe_batch.py 512 0 0 | tail -1
