#!/bin/bash
TEBGPU_LAT_TIMING=1 TEBGPU_WARP_SOLVER=3 timeout 300 python tools/kernel_times.py C3 1 1 2>&1 | grep "k_solve_lat" | sort | uniq -c | sort -rn | head -3
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q -k "short_bands or latency_solver" 2>&1 | tail -3
TEBGPU_WARP_SOLVER=3 timeout 300 python tools/kernel_times.py C3 1 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['e2e_ms_per_call_unprofiled'], {k: (round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k, v in d['kernels'].items()}, d['cost_checksum'])"
