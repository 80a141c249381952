#!/bin/bash
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  timeout 120 python tools/probe_pair_check.py 2>/dev/null | head -1
  for w in c1 c2 c5; do timeout 100 python tools/probe_tl.py $w 0 2>/dev/null | head -1; done
  timeout 100 python tools/probe_tl_c3.py 0 2>/dev/null | head -1
  timeout 100 python tools/probe_batch.py 4096 0 0 2>/dev/null | tail -1
  timeout 100 python tools/probe_batch.py 512 0 0 2>/dev/null | tail -1
  timeout 100 python tools/probe_tl.py c2 1 2>/dev/null | head -1
done
