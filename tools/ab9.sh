#!/bin/bash
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  for w in c2 c5 c1; do timeout 100 python tools/probe_tl.py $w 0 | head -1; done
  timeout 100 python tools/probe_tl.py c2 1 | head -1
  timeout 100 python tools/probe_tl.py c5 1 | head -1
  timeout 100 python tools/probe_tl.py c2 0 | grep -E "^   1[5-7] |last pass"
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
cp build/libdef.so correlation_b200/libdic_b200.so
