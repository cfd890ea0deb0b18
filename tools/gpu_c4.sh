#!/bin/bash
mkdir -p gpurun_out
for K in k_solve_lat k_trial_eval3; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 6 -c 1 -f -o gpurun_out/prof_single_$K python tools/kernel_times.py C3 1 1 > gpurun_out/ncu_single_$K.log 2>&1
done
ls -la gpurun_out/prof_single_k_solve_lat.ncu-rep gpurun_out/prof_single_k_trial_eval3.ncu-rep
