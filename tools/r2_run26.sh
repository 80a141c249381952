#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_e2e.py 2>&1 | tee gpurun_out/r2_probe_e2e.log | tail -12
