"""Generates tests/golden/golden_v1.npz and golden_v2.npz: oracle outputs (analytic-Jacobian mode, deterministic) for small cases.

The reference holds no numeric golden vectors for this path (only test/teb_basics.cpp property
tests) and cannot be run here, so these vectors are the ORACLE's outputs, committed to (a) guard the oracle against
drift and (b) give the GPU tests a fixture that does not need the oracle library at all.
Run:  python -m tests.golden.make_golden
"""
import os

import numpy as np

from teb_local_planner_b200 import abi, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    out = []
    for cfg, cand, autosize in (("C1", 3, True), ("C2", 4, False), ("C3", 3, True), ("C4", 3, False)):
        p, hb = scenes.make_config_batch(cfg, candidates=cand, seed=11, autosize=autosize)
        args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                             p.selection_viapoint_cost_scale, False)
        out.append((f"{cfg}_{'auto' if autosize else 'fixed'}", hb, p, args))
    return out


def cases_v2():
    """holonomic robot (the scene of tests/test_gpu_parity.py::test_holonomic_edges) and vertex-list shapes with a polygon
    footprint (oracle drift guard only: at a closest-feature switch the GPU may legitimately take another LM branch)"""
    out = []
    p, hb = scenes.make_config_batch("C3", candidates=6, seed=8)
    p.max_vel_y, p.acc_lim_y, p.max_vel_trans = 0.3, 0.5, 0.45
    p.weight_kinematics_nh, p.weight_max_vel_y, p.weight_acc_lim_y = 1.0, 2.0, 1.5
    rng = np.random.default_rng(3)
    for b in range(hb.B):
        n = hb.n[b]
        hb.poses[b, 1:n - 1, 2] += rng.normal(0, 0.35, n - 2)
        hb.poses[b, :n - 1, 3] *= 0.8
    hb.vel_start[:, 0], hb.vel_start[:, 1], hb.vel_start[:, 2] = 0.25, 0.1, -0.1
    hb.vel_goal[1, 3] = 0.0
    hb.vel_goal[2, 1] = 0.15
    out.append(("holonomic", hb, p, abi.make_args(5, 4, True, 100.0, 1.0, False), True))
    p2, hb2 = scenes.make_config_batch("C4", candidates=4, seed=6)
    hb2 = scenes.add_shape_obstacles(hb2, seed=1)
    hb2.obstacles["dynamic"][0, ::3] = 0
    scenes.set_polygon_footprint(p2)
    out.append(("shapes_polygon_footprint", hb2, p2, abi.make_args(5, 4, True, 100.0, 1.0, False), False))
    return out


def hsig_cases():
    """candidate classification inputs: (name, params, batch)"""
    out = []
    for cfg in ("C1", "C4"):
        p, hb = scenes.make_config_batch(cfg, candidates=12, seed=13)
        if cfg != "C4":
            p.include_dynamic_obstacles = 0
        out.append((f"hsig_{cfg}", p, hb))
    return out


def main_v2():
    from tests import oracle_binding as ob
    data = {}
    for name, hb_in, p, args, _gpu_exact in cases_v2():
        hb = hb_in.copy()
        ob.optimize_batch(p, hb, args, jac_mode=ob.JAC_ANALYTIC)
        data[f"{name}_poses"], data[f"{name}_n"], data[f"{name}_cost"], data[f"{name}_status"] = hb.poses, hb.n, hb.cost, hb.status
    for name, p, hb in hsig_cases():
        obst = hb.obstacles[0][:hb.obst_count[0]]
        vals = [ob.h_signature(p, hb.poses[b], hb.n[b], obst) for b in range(hb.B)]
        if p.include_dynamic_obstacles:
            data[name] = np.array(vals)
        else:
            data[name] = np.array([[v.real, v.imag] for v in vals])
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **data)
    print("wrote", os.path.join(HERE, "golden_v2.npz"), {k: v.shape for k, v in data.items()})


def main():
    from tests import oracle_binding as ob
    data = {}
    for name, hb_in, p, args in cases():
        hb = hb_in.copy()
        ob.optimize_batch(p, hb, args, jac_mode=ob.JAC_ANALYTIC)
        data[f"{name}_poses"] = hb.poses
        data[f"{name}_n"] = hb.n
        data[f"{name}_cost"] = hb.cost
        data[f"{name}_status"] = hb.status
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **data)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
    main_v2()
