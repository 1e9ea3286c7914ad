#!/bin/bash
# One gpurun call: parity tests, bench (ours + reference arm), ncu launch list, one --set full capture
# of the dominant kernel. Usage: bash scripts/gpu_round.sh [tag] ; outputs land in gpurun_out/<tag>_*
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/${TAG}_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
fi
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_umma3 -s 2 -c 1 -o gpurun_out/${TAG}_ncu_count_full -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_count_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_count_full_raw.csv 2>/dev/null
fi
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json
