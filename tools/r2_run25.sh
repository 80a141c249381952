#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s --durations=8 > gpurun_out/r2_pytest25.log 2>&1
echo "pytest rc $? wall $(( $(date +%s) - t0 )) s"
grep -E "passed|failed|^FAILED|^ERROR|s call" gpurun_out/r2_pytest25.log | tail -14
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
