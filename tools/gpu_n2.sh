#!/bin/bash
# default workload on 2 GPUs (weak scaling, NCCL cost all-gather behind the C-ABI)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-single-request 2>/dev/null | grep '^{' > gpurun_out/final_bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/final_bench_n2.json').read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"
