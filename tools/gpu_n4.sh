#!/bin/bash
# BASELINE config 4 on 4 GPUs: C4 (150 poses, 32 moving obstacles, 4 via-points), 512 bands in total, STRONG scaling,
# one NCCL all-gather of the costs per step behind the C-ABI; plus the same job on 1 GPU of the box and the default
# weak-scaling line at N = 4
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
python bench.py --workload C4 --total-bands 512 --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | grep '^{' > $O/final_bench_c4_strong_n1.json
for N in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload C4 --total-bands 512 --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | grep '^{' > $O/final_bench_c4_strong_n$N.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > $O/final_bench_n4.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final_bench_c4_strong_n*.json')) + ['gpurun_out/final_bench_n4.json']:
    for l in open(f):
        d = json.loads(l)
        print(f.split('/')[-1], d['n_gpus'], d['scaling'], d['config'].get('bands_per_gpu'), 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
PY
