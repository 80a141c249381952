#!/bin/bash
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 4096 1 0 | tail -1
  timeout 100 python tools/probe_batch.py 1024 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 512 0 0 | tail -1
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
cp build/libdef.so correlation_b200/libdic_b200.so
