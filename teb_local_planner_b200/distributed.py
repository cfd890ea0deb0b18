"""Multi-GPU plumbing: one process per GPU, the batch axis (candidates x requests) sharded contiguously, ONE
all-gather of the per-candidate costs per plan, then HomotopyClassPlanner::selectBestTeb on the gathered costs
(reference src/homotopy_class_planner.cpp:466-493 fan-out, :564-616 selection). Bands never exchange data during
optimisation, so there is no data-path collective.
"""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous split of the batch axis: rank g owns bands [lo, hi). Sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_costs(local_cost, world=None, group=None):
    """Single all-gather of the local cost vector (torch tensor, cuda for NCCL / cpu for gloo).

    Shards must have equal length (pad with +inf on the caller side otherwise). Returns the [world * len] tensor.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if world is None else world
    out = torch.empty(world * local_cost.numel(), dtype=local_cost.dtype, device=local_cost.device)
    dist.all_gather_into_tensor(out, local_cost.contiguous(), group=group)
    return out


def select_best(cost, last_best=-1, initial_plan=-1, hysteresis=1.0, prefer_initial=0.95):
    """selectBestTeb on one request's candidate costs (strict '<', first minimum wins)."""
    best, min_cost = -1, np.finfo(np.float64).max
    for i, c in enumerate(cost):
        if i == last_best:
            c = c * hysteresis
        elif i == initial_plan:
            c = c * prefer_initial
        if c < min_cost:
            best, min_cost = i, c
    return best


def select_best_per_request(all_cost, candidates, params=None, last_best=None, initial_plan=None):
    """all_cost: [requests * candidates] gathered costs. Returns the winning candidate index per request."""
    all_cost = np.asarray(all_cost, dtype=np.float64).reshape(-1, candidates)
    hyst = 1.0 if params is None else params.selection_cost_hysteresis
    pref = 0.95 if params is None else params.selection_prefer_initial_plan
    out = np.empty(all_cost.shape[0], np.int32)
    for r in range(all_cost.shape[0]):
        lb = -1 if last_best is None else int(last_best[r])
        ip = -1 if initial_plan is None else int(initial_plan[r])
        out[r] = select_best(all_cost[r], lb, ip, hyst, pref)
    return out
