#!/bin/bash
# latency solver: parity tests, then per-kernel times of one 32-candidate C3 request with each solver
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q -k "latency_solver or warp_solver or short_bands" -s 2>&1 | grep -E "short bands|passed|failed|Error|assert" | head -40
for mode in 0 3; do
  for req in 1 4; do
    TEBGPU_WARP_SOLVER=$mode timeout 300 python tools/kernel_times.py C3 $req 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('mode $mode req $req', d['e2e_ms_per_call_unprofiled'], {k: (round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
  done
done
TEBGPU_WARP_SOLVER=3 timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('single_request', json.dumps(d.get('single_request'))); print('value', d['value'])
"
