# usage: bash tools/gpu_ncu.sh TAG  (on the GPU box): launch list + full captures of the main kernels on the default bench
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-single-request > gpurun_out/ncu_${TAG}_l.log 2>&1
for K in k_linearize2 k_solve_tpb k_trial_eval2; do
  TEBGPU_OVERLAP=0 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 2 -c 1 -f -o gpurun_out/prof_${TAG}_$K python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-single-request > gpurun_out/ncu_${TAG}_$K.log 2>&1
done
ls -la gpurun_out/ | tail -8
