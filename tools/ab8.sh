#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "quadratic or tile_kernel or full_size_c2 or annulus or rowsplit" 2>&1 | tail -2
for w in c2 c1 c5; do timeout 100 python tools/probe_tl.py $w 0 | head -1; done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
timeout 100 python tools/probe_tl.py c2 1 2>&1 | head -1
timeout 100 python tools/probe_tl.py c2 0 2>&1 | sed -n 2,22p
