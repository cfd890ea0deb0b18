"""Generates tests/golden/golden_ref_v1.npz from oracle/_ref/libteb_ref.so, i.e. from the REFERENCE'S OWN code
(src/optimal_planner.cpp, src/timed_elastic_band.cpp, src/obstacles.cpp + headers; only Eigen / boost / ROS / the g2o
optimizer are stand-ins, see oracle/ref_driver.cpp). Run where /root/reference exists:

    python -m tests.golden.make_golden_ref

Inputs are the seeded scenarios of tests/scenarios.py (regenerated, not stored). Stored per scenario NAME:
  NAME/n, NAME/poses   final band of TebOptimalPlanner::optimizeTEB (defaults: 4 x 5 LM iterations, cost computed)
  NAME/cost, NAME/trials, NAME/terminated
  NAME/chi2_0, NAME/b_0   activeChi2 and the right-hand side of the normal equations at the INITIAL state (multiplier 1)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from teb_local_planner_b200 import abi  # noqa: E402
from tests import ref_binding as rb, scenarios  # noqa: E402

OUT = os.path.join(HERE, "golden_ref_v1.npz")


def generate(only=None):
    out = {}
    for name in scenarios.ALL:
        if only is not None and name not in only:
            continue
        p, hb = scenarios.scenario(name)
        args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                             p.selection_viapoint_cost_scale, False)
        n_out, poses, cost, trials, term, chi0, b0 = [], [], [], [], [], [], []
        for b in range(hb.B):
            kw = scenarios.band_kwargs(hb, b)
            n = int(hb.n[b])
            H, rhs, c2 = rb.build_system(p, hb.poses[b], n, **kw)
            rec, c, st, ok = rb.optimize_band(p, hb.poses[b], n, args=args, n_cap=hb.n_cap, **kw)
            assert ok
            full = np.zeros((hb.n_cap, 4))
            full[:len(rec)] = rec
            n_out.append(len(rec)); poses.append(full); cost.append(c); trials.append(st["lm_trials"])
            term.append(st["terminated"]); chi0.append(c2)
            bb = np.zeros(4 * hb.n_cap)
            bb[:len(rhs)] = rhs
            b0.append(bb)
        nmax = max(n_out)
        out[name + "/n"] = np.array(n_out, np.int32)
        out[name + "/poses"] = np.array(poses)[:, :nmax]
        out[name + "/cost"] = np.array(cost)
        out[name + "/trials"] = np.array(trials, np.int32)
        out[name + "/terminated"] = np.array(term, np.bool_)
        out[name + "/chi2_0"] = np.array(chi0)
        out[name + "/b_0"] = np.array(b0)[:, :4 * int(hb.n.max()) - 7]
    return out


if __name__ == "__main__":
    data = generate()
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes,", len(data), "arrays")
