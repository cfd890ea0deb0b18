"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding of the batch axis, ONE all-gather of the
per-candidate costs, selectBestTeb on the gathered vector. The per-band optimisation is replaced by a deterministic
stand-in cost so the test needs neither a GPU nor the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_costs(B):
    rng = np.random.default_rng(123)
    return rng.uniform(1.0, 100.0, B)


def _worker(rank, world, port, B, candidates, out_dir):
    sys.path.insert(0, ROOT)
    from teb_local_planner_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(B, rank, world)
    local = torch.from_numpy(_fake_costs(B)[lo:hi].copy())
    allc = D.gather_costs(local, world)
    best = D.select_best_per_request(allc.numpy(), candidates)
    np.save(os.path.join(out_dir, f"best_{rank}.npy"), best)
    np.save(os.path.join(out_dir, f"all_{rank}.npy"), allc.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shard_gather_select_world2(tmp_path):
    world, B, candidates = 2, 64, 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, candidates, str(tmp_path)), nprocs=world, join=True)
    ref_cost = _fake_costs(B)
    ref_best = ref_cost.reshape(-1, candidates).argmin(axis=1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"all_{r}.npy"), ref_cost)
        assert np.array_equal(np.load(tmp_path / f"best_{r}.npy"), ref_best)
