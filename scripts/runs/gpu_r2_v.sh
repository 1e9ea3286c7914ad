#!/bin/bash
# round 2: full GPU suite (device sort, aggregate, shards), memcheck of the new kernels, sorted-output file bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -n 12 gpurun_out/v_pytest.log
( timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_consumers.py ) > gpurun_out/v_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/v_memcheck.log
tail -n 12 gpurun_out/v_memcheck.log
( time timeout 600 python scripts/file_bench.py --variants 200000 --sorted ) > gpurun_out/v_file_bench.jsonl 2> gpurun_out/v_file_bench.err
echo "file_bench rc=$?"; cat gpurun_out/v_file_bench.jsonl; tail -n 4 gpurun_out/v_file_bench.err
