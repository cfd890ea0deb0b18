#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for m in 0 1; do
TEBGPU_SPLIT=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('split=$m value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['best_candidate_of_request0'])
"
done
