python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_a.json"))
for k in ("value","ms_per_step","e2e","kernel_time_share","roofline_kernel_a","roofline_kernel_b","single_request","parity_sample","clocks"):
    print(k, json.dumps(d.get(k))[:700])
print("cpu", json.dumps(d["cpu_baseline"])[:900])
PY
