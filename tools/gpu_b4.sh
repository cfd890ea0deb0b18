python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for W in 0 1; do
TEBGPU_WARP_SOLVER=$W python tools/kernel_times.py C3 1 30 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 single warp_solver=$W', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
TEBGPU_WARP_SOLVER=$W python tools/kernel_times.py C2 1 30 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 single warp_solver=$W', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
done
TEBGPU_WARP_SOLVER=1 python tools/kernel_times.py C3 4 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 4 requests warp=1', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
TEBGPU_WARP_SOLVER=0 python tools/kernel_times.py C3 4 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 4 requests warp=0', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
TEBGPU_WARP_SOLVER=1 python tools/kernel_times.py C3 16 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 16 requests warp=1', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
TEBGPU_WARP_SOLVER=0 python tools/kernel_times.py C3 16 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 16 requests warp=0', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['single_request'])"
