#!/bin/bash
# round 2: paced count kernel + staged statistics kernel as the product default: GPU suite, smoke, bench N=1, launch list and
# ncu --set full of the C2 count kernel, statistics kernel (C1) and planes kernel (C3)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest.log
tail -n 6 gpurun_out/z_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/z_smoke.log 2>&1; tail -n 2 gpurun_out/z_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err
echo "bench rc=$?"; tail -n 3 gpurun_out/z_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z_bench_n1.json').read().strip().splitlines()[-1])
print('primary value %.4g ms %.3f count %.3f stats %.3f e2e %.4g (%.3f ms) frac %.4f clocks %s' % (d['value'], d['ms_per_step'], d['config']['ms_count_kernel_per_step'], d['config']['ms_stats_kernel_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks']))
print('parity', d.get('parity'))
for x in d.get('extra_configs',[]): print(x['baseline_config'], 'value %.4g ms %.2f count %.2f stats %.2f e2e %.4g frac %.3f parity %s' % (x['value'], x['ms_per_step'], x['ms_count_kernel_per_step'], x['ms_stats_kernel_per_step'], x['e2e']['value'], x['roofline']['frac'], (x.get('parity') or {}).get('checked')))
PY
B="--no-extra --no-cpu-baseline --no-mma-ceiling"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches_c2.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/z_launches_c2.out 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_umma3 -s 2 -c 1 -f -o gpurun_out/z_c2_count python bench.py --steps 2 --warmup 1 $B > gpurun_out/z_c2_count.out 2>&1
ncu -i gpurun_out/z_c2_count.ncu-rep --page raw --csv > gpurun_out/z_c2_count_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stats_kernel -s 1 -c 1 -f -o gpurun_out/z_c1_stats python scripts/profile_cfg.py c1 > gpurun_out/z_c1_stats.out 2>&1
ncu -i gpurun_out/z_c1_stats.ncu-rep --page raw --csv > gpurun_out/z_c1_stats_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_umma3 -s 1 -c 1 -f -o gpurun_out/z_c3_planes python scripts/profile_cfg.py c3 30000 > gpurun_out/z_c3_planes.out 2>&1
ncu -i gpurun_out/z_c3_planes.ncu-rep --page raw --csv > gpurun_out/z_c3_planes_raw.csv 2>/dev/null
rm -f gpurun_out/z_c3_planes.ncu-rep gpurun_out/z_c1_stats.ncu-rep
grep -h "count_ms" gpurun_out/z_c1_stats.out gpurun_out/z_c3_planes.out
ls -la gpurun_out/z_*
