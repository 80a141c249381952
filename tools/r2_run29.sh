#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-others > gpurun_out/r2_bench29.json 2> gpurun_out/r2_bench29.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench29.json").read().strip().splitlines()[-1])
print("c4 value %.1f G ms %.3f e2e %.1f G (%.3f ms over %s steps) frac %.4f kernel ms %.3f launches %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("steps"), d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["gpu_launches"]))
PY
tail -2 gpurun_out/r2_bench29.err
