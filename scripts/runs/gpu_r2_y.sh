#!/bin/bash
# round 2: drift probe (per-tile pacing vs free running): times, then DRAM bytes per launch under ncu
mkdir -p gpurun_out
( timeout 300 python scripts/drift_probe.py 3 ) > gpurun_out/y_drift_times.log 2>&1
cat gpurun_out/y_drift_times.log
( timeout 400 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:count_umma3 --csv --log-file gpurun_out/y_drift_ncu.csv python scripts/drift_probe.py 1 ) > gpurun_out/y_drift_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/y_drift_ncu.csv') if l.startswith('"'))]
h=rows[0]; iM=h.index('Metric Name'); iV=h.index('Metric Value'); iI=h.index('ID')
d={}
for r in rows[1:]: d.setdefault(int(r[iI]),{})[r[iM]]=float(r[iV].replace(',',''))
for k in sorted(d):
    v=d[k]
    if v['gpu__time_duration.sum']>5e6: print(k//2, 'dram GB %.1f  ms %.2f  hit %.1f' % (v['dram__bytes_read.sum']/1e9, v['gpu__time_duration.sum']/1e6, v['lts__t_sector_hit_rate.pct']))
PY
