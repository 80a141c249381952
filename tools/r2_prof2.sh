#!/bin/bash
# ncu capture of one library variant on the c4 batch probe: tools/r2_prof2.sh build/libC.so tag
mkdir -p gpurun_out/prof
cp "$1" correlation_b200/libdic_b200.so
tag=$2
python tools/probe_batch.py 4096 0 0 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 2 -c 1 -f -o gpurun_out/prof/$tag python tools/probe_batch.py 4096 0 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof/$tag.ncu-rep 40 > gpurun_out/prof/${tag}_ncu_full.txt 2>&1
python tools/sass_hist.py gpurun_out/prof/$tag.ncu-rep 267295560 > gpurun_out/prof/${tag}_sass_hist.txt 2>&1
head -60 gpurun_out/prof/${tag}_ncu_full.txt
