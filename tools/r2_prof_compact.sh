#!/bin/bash
mkdir -p gpurun_out/prof2
python tools/probe_domain.py > gpurun_out/prof2/domain_build.txt 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"compact|tiles_scatter|expand_spans|rect_fill|rect_tiles|bbox" -c 60 --csv --log-file gpurun_out/prof2/launches_domain.csv python tools/probe_domain.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"compact_kernel" -s 1 -c 1 -f -o gpurun_out/prof2/compact python tools/probe_domain.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof2/compact.ncu-rep 10 > gpurun_out/prof2/r2_compact_ncu_full.txt 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/prof2/launches_domain.csv")))
hi=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
h=rows[hi]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    d.setdefault((int(r[ii]), r[ik][:70]),{})[r[im]]=float(r[iv].replace(",",""))
for (i,k),m in sorted(d.items())[:40]:
    print(i, k, {a.split("__")[1][:16]:round(b,1) for a,b in m.items()})
PY
cat gpurun_out/prof2/domain_build.txt
head -30 gpurun_out/prof2/r2_compact_ncu_full.txt
rm -f gpurun_out/prof2/compact.ncu-rep
