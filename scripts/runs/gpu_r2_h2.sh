#!/bin/bash
# round 2, final 8-GPU call: bench at N = 8 / 4 (N = 2 and N = 1 were run on smaller boxes)
mkdir -p gpurun_out
bash scripts/runs/gpu_r2_g.sh 8
bash scripts/runs/gpu_r2_g.sh 4
