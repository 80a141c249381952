#!/bin/bash
# timing-only A/B of libdic_b200.so builds (parity mode): tools/ab3.sh build/libX.so ...
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 100 python tools/probe_tl.py c2 0 | head -1
  timeout 100 python tools/probe_tl.py c5 0 | head -1
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 1024 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 512 0 0 | tail -1
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
