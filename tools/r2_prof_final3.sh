#!/bin/bash
# round-2 final ncu captures of the committed build (each only after the same command has run clean without ncu)
mkdir -p gpurun_out/prof3
B="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline --parity-sample 2"
$B > gpurun_out/prof3/plain_c4.json 2> gpurun_out/prof3/plain_c4.err || exit 1
$B --workload c2 > gpurun_out/prof3/plain_c2.json 2> gpurun_out/prof3/plain_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gn_solve|pyramid_level|rect_grid|compact|tiles_|rect_fill|copy_rows" -c 900 --csv --log-file gpurun_out/prof3/launches_c4.csv $B > /dev/null 2>&1
python tools/ncu_launch_list.py gpurun_out/prof3/launches_c4.csv > gpurun_out/prof3/r2_c4_launch_list.txt 2>&1
for m in parity fast; do
  ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 3 -c 1 -f -o gpurun_out/prof3/c4_$m $B --mode $m > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 3 -c 1 -f -o gpurun_out/prof3/c2_$m $B --workload c2 --mode $m > /dev/null 2>&1
done
rm -f gpurun_out/prof3/r2_traffic_raw.txt
for r in c4_parity c4_fast c2_parity c2_fast; do
  [ -f gpurun_out/prof3/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof3/$r.ncu-rep 30 > gpurun_out/prof3/r2_${r}_ncu_full.txt 2>&1
  ncu -i gpurun_out/prof3/$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
d=dict(zip(h,zip(u,v)))
for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'): print('$r', k, d[k])
" >> gpurun_out/prof3/r2_traffic_raw.txt 2>&1
done
WORK=$(python -c "import json;print(json.loads(open('gpurun_out/prof3/plain_c4.json').read().strip().split(chr(10))[-1])['config']['pixel_evaluations_per_step'])")
python tools/sass_hist.py gpurun_out/prof3/c4_parity.ncu-rep $WORK > gpurun_out/prof3/r2_c4_parity_sass_hist.txt 2>&1
python tools/sass_hist.py gpurun_out/prof3/c4_fast.ncu-rep $WORK > gpurun_out/prof3/r2_c4_fast_sass_hist.txt 2>&1
WORK2=$(python -c "import json;print(json.loads(open('gpurun_out/prof3/plain_c2.json').read().strip().split(chr(10))[-1])['config']['pixel_evaluations_per_step'])")
python tools/sass_hist.py gpurun_out/prof3/c2_parity.ncu-rep $WORK2 > gpurun_out/prof3/r2_c2_parity_sass_hist.txt 2>&1
rm -f gpurun_out/prof3/*.ncu-rep
cat gpurun_out/prof3/r2_traffic_raw.txt
head -8 gpurun_out/prof3/r2_c4_parity_sass_hist.txt
