#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -x -q -k "per_trial or latency_solver or graph" 2>&1 | tail -5
for m in 0 2; do
TEBGPU_EVAL3=$m timeout 300 python tools/kernel_times.py C3 1 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('eval3=$m single', d['e2e_ms_per_call_unprofiled'], {k: (round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('single_request', json.dumps(d.get('single_request'))); print('value', d['value'])
"
