"""world_size-2 test on CPU (gloo) of the N > 1 path: the batch axis (requests x candidates) is sharded over two
processes, every rank REALLY optimises its shard (with the CPU oracle standing in for the GPU kernels, which need a
device), the per-candidate costs are all-gathered once, and selectBestTeb runs on the gathered vector. Both ranks must
end with the costs and winners of a single-process run over the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES, REQUESTS = 4, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    from teb_local_planner_b200 import abi, scenes
    p, hb = scenes.make_config_batch("C1", requests=REQUESTS, seed=17, candidates=CANDIDATES)
    args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                         p.selection_viapoint_cost_scale, False)
    return p, hb, args


def _shard(hb, lo, hi):
    """bands [lo, hi) with their scenes (requests are not split: CANDIDATES divides the shard bounds)"""
    from teb_local_planner_b200 import abi
    s_lo, s_hi = lo // CANDIDATES, hi // CANDIDATES
    return abi.HostBatch(hb.poses[lo:hi], hb.n[lo:hi], hb.obstacles[s_lo:s_hi], hb.obst_count[s_lo:s_hi],
                         hb.scene_id[lo:hi] - s_lo)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from teb_local_planner_b200 import distributed as D
    from tests import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p, hb, args = _problem()
    r_lo, r_hi = D.shard_requests(REQUESTS, rank, world)
    mine = _shard(hb, r_lo * CANDIDATES, r_hi * CANDIDATES)
    ob.optimize_batch(p, mine, args, jac_mode=ob.JAC_ANALYTIC, threads=2)      # the rank's share of optimizeAllTEBs
    allc = D.gather_costs_torch(torch.from_numpy(mine.cost.copy()), world)       # the ONE collective
    best = D.select_best_per_request(allc.numpy(), CANDIDATES, p)
    np.save(os.path.join(out_dir, f"best_{rank}.npy"), best)
    np.save(os.path.join(out_dir, f"all_{rank}.npy"), allc.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_shard_optimize_gather_select_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from teb_local_planner_b200 import distributed as D
    from tests import oracle_binding as ob
    p, hb, args = _problem()
    ob.optimize_batch(p, hb, args, jac_mode=ob.JAC_ANALYTIC, threads=2)          # single process, whole batch
    ref_best = D.select_best_per_request(hb.cost, CANDIDATES, p)
    assert np.all(np.isfinite(hb.cost)) and len(set(ref_best.tolist())) > 1       # a non-trivial selection
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"all_{r}.npy"), hb.cost)
        assert np.array_equal(np.load(tmp_path / f"best_{r}.npy"), ref_best)


def test_shard_ranges_cover_the_batch():
    from teb_local_planner_b200 import distributed as D
    for B in (1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            got = [D.shard_range(B, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == B
            assert all(got[k][1] == got[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in got) - min(h - l for l, h in got) <= 1
