"""GPU vs closed-form oracle on the vertex-list scenarios: where does a band first deviate? (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi
from tests import oracle_binding as ob, scenarios

for name in ("shapes_polygon", "shapes_line", "shapes_two_circles", "shapes_point"):
    p, hb = scenarios.scenario(name, candidates=24)
    g = T.TebGpu(hb.B, hb.n_cap, hb.S, max(hb.M_cap, 1), hb.V_cap, max_obst_vertices=hb.PV_cap)
    g.set_params(p)
    print("==", name)
    for (inner, outer) in ((1, 1), (2, 1), (5, 1), (5, 2), (5, 4)):
        args = abi.make_args(inner, outer, True, 100.0, 1.0, False)
        hg = hb.copy(); g.optimize(hg, args)
        ha = hb.copy(); ob.optimize_batch(p, ha, args, jac_mode=ob.JAC_ANALYTIC, threads=8)
        d = np.array([np.abs(hg.poses[k, :hg.n[k]] - ha.poses[k, :ha.n[k]]).max() for k in range(hb.B)])
        bad = np.where(d > 1e-9)[0]
        print(f" inner={inner} outer={outer}: max {d.max():.2e} median {np.median(d):.2e} bands>1e-9 {bad.tolist()} "
              f"iters_equal {np.array_equal(hg.lm_iters, ha.lm_iters)} cost_rel {np.abs(hg.cost-ha.cost).max()/np.abs(ha.cost).max():.2e} "
              f"status_equal {np.array_equal(hg.status, ha.status)}")
        if len(bad) and inner == 1 and outer == 1:
            k = int(bad[0])
            i = int(np.abs(hg.poses[k, :hg.n[k]] - ha.poses[k, :ha.n[k]]).max(axis=1).argmax())
            print("   first bad band", k, "pose", i, hg.poses[k, i], ha.poses[k, i], "status", hg.status[k], ha.status[k], "iters", hg.lm_iters[k], ha.lm_iters[k])
    # the linear system at the initial state and after one accepted step
    Hb, chi2 = g.build_system(hb, 0)
    worst = 0
    for k in range(hb.B):
        kw = scenarios.band_kwargs(hb, k)
        Hd, bd, c2 = ob.build_system(p, hb.poses[k], int(hb.n[k]), jac_mode=ob.JAC_ANALYTIC, **kw)
        worst = max(worst, abs(chi2[k] - c2) / max(c2, 1))
    print(" chi2 rel diff at the initial state:", worst)
    g.close()
