#!/bin/bash
# ncu captures of round 2 (each only after the same command has run clean without ncu)
mkdir -p gpurun_out/prof
B="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline --parity-sample 2"
$B > gpurun_out/prof/plain_c4.json 2> gpurun_out/prof/plain_c4.err || exit 1
$B --workload c2 > gpurun_out/prof/plain_c2.json 2> gpurun_out/prof/plain_c2.err || exit 1
python tools/probe_domain.py > gpurun_out/prof/domain_build.txt 2>&1 || exit 1
cat gpurun_out/prof/domain_build.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gn_solve|pyramid_level|rect_grid|compact|tiles_|rect_fill|copy_rows" -c 400 --csv --log-file gpurun_out/prof/launches_c4.csv $B > /dev/null 2>&1
python tools/ncu_launch_list.py gpurun_out/prof/launches_c4.csv > gpurun_out/prof/r2_c4_launch_list.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 3 -c 1 -f -o gpurun_out/prof/c4 $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_solve_tiles -s 3 -c 1 -f -o gpurun_out/prof/c2 $B --workload c2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"compact_kernel<.*AnnulusPred, 1>" -s 0 -c 1 -f -o gpurun_out/prof/compact python tools/probe_domain.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rect_grid -s 2 -c 2 -f -o gpurun_out/prof/rectgrid python tools/probe_domain.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pyramid_level -s 6 -c 1 -f -o gpurun_out/prof/pyr $B > /dev/null 2>&1
for r in c4 c2 compact rectgrid pyr; do
  [ -f gpurun_out/prof/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof/$r.ncu-rep 30 > gpurun_out/prof/r2_${r}_ncu_full.txt 2>&1
done
WORK=$(python -c "import json;print(json.loads(open('gpurun_out/prof/plain_c4.json').read().strip().split(chr(10))[-1])['config']['pixel_evaluations_per_step'])")
python tools/sass_hist.py gpurun_out/prof/c4.ncu-rep $WORK > gpurun_out/prof/r2_c4_sass_hist.txt 2>&1
WORK2=$(python -c "import json;print(json.loads(open('gpurun_out/prof/plain_c2.json').read().strip().split(chr(10))[-1])['config']['pixel_evaluations_per_step'])")
python tools/sass_hist.py gpurun_out/prof/c2.ncu-rep $WORK2 > gpurun_out/prof/r2_c2_sass_hist.txt 2>&1
ncu -i gpurun_out/prof/compact.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:60], 'dur', d.get('gpu__time_duration.sum'), 'rd', d.get('dram__bytes_read.sum'), 'wr', d.get('dram__bytes_write.sum'))
" > gpurun_out/prof/r2_compact_raw.txt 2>&1
rm -f gpurun_out/prof/rectgrid.ncu-rep gpurun_out/prof/pyr.ncu-rep gpurun_out/prof/compact.ncu-rep
ls -la gpurun_out/prof
head -40 gpurun_out/prof/r2_c4_ncu_full.txt
