#!/bin/bash
# A/B two builds of libdic_b200.so on the same box: tools/ab.sh build/libA.so build/libB.so
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  echo "=== $lib"
  for m in 0 1; do
    timeout 100 python tools/probe_tl.py c2 $m | head -1
    timeout 100 python tools/probe_tl.py c1 $m | head -1
    timeout 100 python tools/probe_tl.py c5 $m | head -1
  done
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1; timeout 100 python tools/probe_batch.py 4096 1 0 | tail -1
done
