#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_pipe.py parity 2>&1 | tee gpurun_out/r2_probe_pipe.log | tail -50
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "staged or batch_forms or double_buffer or row_band" 2>&1 | tail -3
