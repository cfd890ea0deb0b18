#!/bin/bash
# round-2 checkpoint: latency-solver timing, whole GPU suite, default bench, ncu of the single-request kernels
mkdir -p gpurun_out
TEBGPU_LAT_TIMING=1 timeout 300 python tools/kernel_times.py C3 1 1 2>&1 | grep "k_solve_lat" | sort | uniq -c | sort -rn | head -2
timeout 300 python tools/kernel_times.py C3 1 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('single', d['e2e_ms_per_call_unprofiled'], {k: (round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 600 gpurun_out/c1_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/c1_bench.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
        print('single_request', d.get('single_request'))
        print('parity_sample', d.get('parity_sample'))
PY
for K in k_trial_eval2 k_linearize2 k_solve_lat; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 6 -c 1 -f -o gpurun_out/prof_single_$K python tools/kernel_times.py C3 1 1 > gpurun_out/ncu_single_$K.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
