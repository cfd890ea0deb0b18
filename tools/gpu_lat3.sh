#!/bin/bash
mkdir -p gpurun_out
TEBGPU_LAT_TIMING=1 TEBGPU_WARP_SOLVER=3 timeout 300 python tools/kernel_times.py C3 1 1 2>&1 | grep "k_solve_lat" | sort | uniq -c | sort -rn | head -2
TEBGPU_WARP_SOLVER=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_solve_lat --launch-skip 4 -c 1 -f -o gpurun_out/prof_lat python tools/kernel_times.py C3 1 1 > gpurun_out/ncu_lat.log 2>&1
ls -la gpurun_out/prof_lat.ncu-rep
