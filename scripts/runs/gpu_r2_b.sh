#!/bin/bash
# round 2, GPU call B: screen ablations (LDS vs FMA), ncu of the statistics kernel at C1, tools tests
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_tools.py -m gpu -x -q ) > gpurun_out/b_pytest_tools.log 2>&1
timeout 600 python scripts/ceiling2.py > gpurun_out/b_ceiling2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stats_kernel -s 1 -c 1 -f -o gpurun_out/b_stats_c1 \
  python bench.py --variants 10000 --min-r2 0 --steps 1 --warmup 1 --no-cpu-baseline --no-mma-ceiling > gpurun_out/b_ncu_stats.log 2>&1
ncu -i gpurun_out/b_stats_c1.ncu-rep --page raw --csv > gpurun_out/b_stats_c1_raw.csv 2>/dev/null
tail -5 gpurun_out/b_pytest_tools.log; cat gpurun_out/b_ceiling2.log; tail -3 gpurun_out/b_ncu_stats.log
