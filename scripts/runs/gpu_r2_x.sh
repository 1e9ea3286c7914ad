#!/bin/bash
# round 2: statistics kernel with the shared-memory log-factorial table + staged record stores (GPU suite, C1 A/B),
# C2 sweep of tile order / L2 eviction hints (times, then DRAM bytes per launch under ncu)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -n 8 gpurun_out/x_pytest.log
for t in 0 1; do
  echo "== TWKB_STATS_TAB_SMEM=$t" >> gpurun_out/x_c1.log
  ( TWKB_STATS_TAB_SMEM=$t timeout 200 python scripts/profile_cfg.py c1 ) >> gpurun_out/x_c1.log 2>&1
done
cat gpurun_out/x_c1.log
( timeout 400 python scripts/l2_sweep.py 3 ) > gpurun_out/x_l2_sweep_times.log 2>&1
cat gpurun_out/x_l2_sweep_times.log
( timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:count_umma3 --csv --log-file gpurun_out/x_l2_sweep_ncu.csv python scripts/l2_sweep.py 1 ) > gpurun_out/x_l2_sweep_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/x_l2_sweep_ncu.csv') if l.startswith('"'))]
h=rows[0]; iM=h.index('Metric Name'); iV=h.index('Metric Value'); iI=h.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iI],{})[r[iM]]=r[iV]
for k,v in d.items(): print(k, v)
PY
