#!/bin/bash
# round 2: ncu / compute-sanitizer evidence for profiles/
mkdir -p gpurun_out
B="--no-extra --no-cpu-baseline --no-mma-ceiling"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches_c2.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/p_launches_c2.out 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_umma3 -s 2 -c 1 -f -o gpurun_out/p_c2_count python bench.py --steps 2 --warmup 1 $B > gpurun_out/p_c2_count.out 2>&1
ncu -i gpurun_out/p_c2_count.ncu-rep --page raw --csv > gpurun_out/p_c2_count_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stats_kernel -s 1 -c 1 -f -o gpurun_out/p_c1_stats python scripts/profile_cfg.py c1 > gpurun_out/p_c1_stats.out 2>&1
ncu -i gpurun_out/p_c1_stats.ncu-rep --page raw --csv > gpurun_out/p_c1_stats_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_umma3 -s 1 -c 1 -f -o gpurun_out/p_c3_planes python scripts/profile_cfg.py c3 30000 > gpurun_out/p_c3_planes.out 2>&1
ncu -i gpurun_out/p_c3_planes.ncu-rep --page raw --csv > gpurun_out/p_c3_planes_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_sparse -s 1 -c 1 -f -o gpurun_out/p_c4_sparse python scripts/profile_cfg.py c4 16000 > gpurun_out/p_c4_sparse.out 2>&1
ncu -i gpurun_out/p_c4_sparse.ncu-rep --page raw --csv > gpurun_out/p_c4_sparse_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:count_umma3 -s 1 -c 1 -f -o gpurun_out/p_c4_count python scripts/profile_cfg.py c4 16000 > gpurun_out/p_c4_count.out 2>&1
ncu -i gpurun_out/p_c4_count.ncu-rep --page raw --csv > gpurun_out/p_c4_count_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py --popc > gpurun_out/p_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/p_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/p_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/p_racecheck.log
rm -f gpurun_out/p_c4_count.ncu-rep gpurun_out/p_c4_sparse.ncu-rep
ls -la gpurun_out/p_*; tail -5 gpurun_out/p_memcheck.log gpurun_out/p_racecheck.log; cat gpurun_out/p_c3_planes.out gpurun_out/p_c4_count.out gpurun_out/p_c1_stats.out | grep -v "==PROF"
