#!/bin/bash
timeout 300 python tools/kernel_times.py C3 256 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('C3 8192', round(d['e2e_ms_per_call_unprofiled'],2), {k: (round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
timeout 300 python tools/kernel_times.py C2 256 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('C2 8192', round(d['e2e_ms_per_call_unprofiled'],2), {k: (round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches_per_step'])
"
