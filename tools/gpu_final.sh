#!/bin/bash
# round-2 final evidence on one B200: GPU test suite, smoke, both bench arms, other workloads, ncu launch list + captures.
# usage (on the GPU box): bash tools/gpu_final.sh bench | ncu | ncu_single ; everything lands in gpurun_out/final_*
# (three calls: gpurun merges at most 64 MiB back, one ncu report is ~14 MB)
mkdir -p gpurun_out
O=gpurun_out
PART=${1:-bench}
if [ "$PART" = bench ]; then
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/final_gputest.log; cat $O/final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > $O/final_bench_reference.json 2> $O/final_bench_reference.err; tail -c 300 $O/final_bench_reference.err
timeout 900 python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err; tail -c 300 $O/final_bench_n1.err
for W in C1 C2 C4; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/final_bench_$W.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final_bench_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], d.get('impl', 'b200'), d['config']['workload'][:28], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3),
                  'e2e', round(d['e2e']['value'], 1), 'single', (d.get('single_request') or {}).get('ms_per_request'), 'roof', (d.get('roofline') or {}).get('frac'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-single-request > $O/final_ncu_l.log 2>&1
TEBGPU_LAT_TIMING=1 timeout 300 python tools/kernel_times.py C3 1 1 2>&1 | grep "k_solve_lat" | sort | uniq -c | sort -rn | head -1 > $O/final_lat_timing.log; cat $O/final_lat_timing.log
timeout 300 python tools/kernel_times.py C3 1 5 > $O/final_single_kernels.json 2>/dev/null
timeout 300 python tools/kernel_times.py C3 256 3 > $O/final_c3_kernels.json 2>/dev/null
fi
if [ "$PART" = ncu ]; then
for K in k_linearize2 k_solve_tpb k_trial_eval2; do
  TEBGPU_OVERLAP=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 2 -c 1 -f -o $O/final_prof_$K python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-single-request > $O/final_ncu_$K.log 2>&1
done
fi
if [ "$PART" = ncu_single ]; then
for K in k_solve_lat k_trial_eval3 k_linearize2; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 6 -c 1 -f -o $O/final_prof_single_$K python tools/kernel_times.py C3 1 1 > $O/final_ncu_single_$K.log 2>&1
done
fi
ls -la $O/ | grep final_ | awk '{print $5, $9}'
