#!/bin/bash
TAG=${1:-c3}
mkdir -p gpurun_out
timeout 300 python scripts/c3_probe.py 20000 > gpurun_out/${TAG}_probe.log 2>&1
cat gpurun_out/${TAG}_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_umma3 -s 1 -c 1 -o gpurun_out/${TAG}_ncu_planes -f python scripts/c3_probe.py 20000 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_planes.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_planes_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_ncu_planes.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Issue\|Eligible\|No Elig" | head -40
