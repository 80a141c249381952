#!/bin/bash
# batch-only timing A/B: tools/ab5.sh build/libX.so ...
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 4096 1 0 | tail -1
  timeout 100 python tools/probe_batch.py 2048 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 1024 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 512 0 0 | tail -1
  timeout 100 python tools/probe_tl.py c1 0 | head -1
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
