#!/bin/bash
for W in C1 C2 C4; do
for cfg in "TEBGPU_OVERLAP=0" "TEBGPU_OVERLAP=1 TEBGPU_SPLIT=0" "TEBGPU_OVERLAP=1 TEBGPU_SPLIT=1"; do
env $cfg timeout 600 python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$W $cfg value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches_per_step'])
"
done
done
