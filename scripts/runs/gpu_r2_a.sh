#!/bin/bash
# round 2, GPU call A: parity suite + epilogue ablations + TMEM read rate + C1 statistics-kernel A/B + default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 120 scripts/tmem_bw > gpurun_out/a_tmem_bw.log 2>&1
timeout 600 python scripts/ceiling.py > gpurun_out/a_ceiling.log 2>&1
for occ in 2 3; do
  TWKB_STATS_OCC=$occ timeout 300 python bench.py --variants 10000 --min-r2 0 --steps 5 --no-cpu-baseline --no-mma-ceiling > gpurun_out/a_bench_c1_occ$occ.json 2> gpurun_out/a_bench_c1_occ$occ.err
done
timeout 600 python bench.py > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_tmem_bw.log gpurun_out/a_ceiling.log; cat gpurun_out/a_bench_c1_occ2.json gpurun_out/a_bench_c1_occ3.json gpurun_out/a_bench_c2.json
