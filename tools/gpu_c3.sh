#!/bin/bash
mkdir -p gpurun_out
TEBGPU_LAT_TIMING=1 timeout 300 python tools/kernel_times.py C3 1 1 2>&1 | grep "k_solve_lat" | sort | uniq -c | sort -rn | head -2
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/kernel_times.py C3 1 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('single', d['e2e_ms_per_call_unprofiled'], {k: (round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('single_request', json.dumps(d.get('single_request'))); print('value', d['value'])
"
