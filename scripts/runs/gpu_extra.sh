#!/bin/bash
# Extra measurements of one gpurun call: super-tile edge sweep of the tensor kernel and the
# file-to-file timing of twkb_calc_file. Usage: bash scripts/gpu_extra.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
for S in 16 24 32 48; do
  echo "TWKB_SUPER=$S" >> gpurun_out/${TAG}_super.log
  TWKB_SUPER=$S timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline >> gpurun_out/${TAG}_super.log 2>> gpurun_out/${TAG}_super.err
done
timeout 900 python scripts/file_bench.py --variants 200000 --reference 30000 > gpurun_out/${TAG}_file_bench.jsonl 2> gpurun_out/${TAG}_file_bench.err
cat gpurun_out/${TAG}_file_bench.jsonl
grep -o '"ms_count_kernel_per_step": [0-9.]*\|TWKB_SUPER=[0-9]*\|"value": [0-9.]*' gpurun_out/${TAG}_super.log | head -40
