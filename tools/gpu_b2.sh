python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for R in 0 10; do
TEBGPU_RING=$R python tools/kernel_times.py C3 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 ring=$R', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
done
for R in 0 10; do
TEBGPU_RING=$R python tools/kernel_times.py C3 1 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 single ring=$R', round(d['e2e_ms_per_call_unprofiled'],3), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
done
python tools/kernel_times.py C2 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 400 gpurun_out/r2_bench_b.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_b.json"))
for k in ("value","ms_per_step","e2e","kernel_time_share","roofline_kernel_a","roofline_kernel_b","single_request","parity_sample"):
    print(k, json.dumps(d.get(k))[:500])
print("cpu", json.dumps(d["cpu_baseline"])[:700])
PY
cat gpurun_out/parity_report.json | python -c "import sys,json; d=json.load(sys.stdin); print({k:(v['fraction_within_1e-4_of_reference'], v['max_abs_pose_diff']) for k,v in d.items()})"
