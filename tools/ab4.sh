#!/bin/bash
# parity-subset + timing A/B: tools/ab4.sh build/libX.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  cp "$lib" correlation_b200/libdic_b200.so
  tag=$(basename $lib .so)
  echo "=== $lib"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "single_evaluation or tile_kernel or full_size_c2 or full_size_c1 or c4_256 or rowsplit_loopback or quadratic or annulus or blob" > gpurun_out/ab_$tag.pytest.log 2>&1
  tail -3 gpurun_out/ab_$tag.pytest.log
  timeout 100 python tools/probe_tl.py c2 0 | head -1
  timeout 100 python tools/probe_tl.py c2 1 | head -1
  timeout 100 python tools/probe_tl.py c5 0 | head -1
  timeout 100 python tools/probe_tl.py c1 0 | head -1
  timeout 100 python tools/probe_batch.py 4096 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 4096 1 0 | tail -1
  timeout 100 python tools/probe_batch.py 1024 0 0 | tail -1
  timeout 100 python tools/probe_batch.py 512 0 0 | tail -1
done 2>&1 | grep -v "Traceback\|File \"\|print(f\|BrokenPipe"
