python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/kernel_times.py C3 256 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', round(d['e2e_ms_per_call_unprofiled'],2), d['K'], {k:(round(v['avg_ms'],4), round(v['ms_per_call'],2)) for k,v in d['kernels'].items()}, d['cost_checksum'])"
for G in 0 1; do
TEBGPU_GRAPH=$G python tools/kernel_times.py C3 1 30 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 single graph=$G', round(d['e2e_ms_per_call_unprofiled'],3))"
done
for S in 1 0; do
python tools/kernel_times.py C3 1 20 set_solver=$S | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 single solver=$S', round(d['e2e_ms_per_call_unprofiled'],3), {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
python tools/kernel_times.py C2 1 20 set_solver=$S | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 single solver=$S', round(d['e2e_ms_per_call_unprofiled'],3), {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
done
python tools/kernel_times.py C2 1 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 single solver=2', round(d['e2e_ms_per_call_unprofiled'],3), {k:(round(v['avg_ms'],4), round(v['ms_per_call'],3)) for k,v in d['kernels'].items()})"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 300 gpurun_out/r2_bench_c.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_c_reference.json 2> gpurun_out/r2_bench_c_ref.err; tail -c 300 gpurun_out/r2_bench_c_ref.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_c.json"))
for k in ("value","ms_per_step","e2e","kernel_time_share","single_request"):
    print(k, json.dumps(d.get(k))[:400])
r=json.load(open("gpurun_out/r2_bench_c_reference.json"))
print("ref", r["value"], r["cpu_baseline"]["effective_cores"], r["cpu_baseline"]["port_value"])
PY
bash tools/gpu_ncu.sh r2d > /dev/null 2>&1; ls gpurun_out | grep r2d
python tools/c5_sweep.py gpurun_out/r2_c5_sweep.json > gpurun_out/r2_c5_sweep.log 2>&1; tail -3 gpurun_out/r2_c5_sweep.log | cut -c1-300
