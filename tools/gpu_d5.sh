#!/bin/bash
for T in 128 256 512; do
python bench.py --workload C4 --total-bands $T --steps 10 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('C4 bands $T value', round(d['value']), 'ms', round(d['ms_per_step'],3))
"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
