python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for V in 0 1; do
TEBGPU_EVAL_V1=$V python tools/kernel_times.py C3 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 evalV1=$V', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
done
TEBGPU_SPEC_K=4 python tools/kernel_times.py C3 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 K=4', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
python tools/kernel_times.py C2 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
python tools/kernel_times.py C3 1 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 single request', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
TEBGPU_OVERLAP=0 ncu --set full --import-source on --clock-control none -k regex:k_linearize2 --launch-skip 6 -c 1 -f -o gpurun_out/prof_r2c_k_linearize2 python tools/kernel_times.py C3 256 1 > gpurun_out/ncu_r2c.log 2>&1
TEBGPU_OVERLAP=0 ncu --set full --import-source on --clock-control none -k regex:k_trial_eval2 --launch-skip 6 -c 1 -f -o gpurun_out/prof_r2c_k_trial_eval2 python tools/kernel_times.py C3 256 1 > gpurun_out/ncu_r2c2.log 2>&1
