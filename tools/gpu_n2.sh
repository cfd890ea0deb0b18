nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 600 gpurun_out/r2_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print('N=2', d['value'], d['n_gpus'], d['e2e'], d['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
teb_local_planner_b200/host/test/test_dropin gpu | grep -E "SHARDED|RESULT"
python - <<'PY'
# two ranks in one process group is covered above; here: gather through the C-ABI with host buffers across 2 ranks
PY
