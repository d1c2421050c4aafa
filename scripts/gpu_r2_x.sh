#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2x_x python scripts/profile_step.py --mode exact > gpurun_out/r2x_ncu.log 2>&1
tail -1 gpurun_out/r2x_ncu.log
bash scripts/gpu_r2_t.sh 2>&1 | tail -12
