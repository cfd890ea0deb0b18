"""Generates tests/golden/golden_v1.npz: oracle outputs (analytic-Jacobian mode, deterministic) for small cases.

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
