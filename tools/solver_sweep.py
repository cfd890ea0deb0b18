"""Throughput of the three LM solvers over batch sizes (host-buffer entry, wall clock around synchronous calls).
usage: python tools/solver_sweep.py [workload]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
for requests in (1, 4, 16, 64):
    p, hb0 = scenes.make_config_batch(wl, requests=requests, seed=1, candidates=32)
    args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, 100.0, 1.0, False)
    g = T.TebGpu(hb0.B, hb0.n_cap, hb0.S, max(hb0.M_cap, 1), hb0.V_cap)
    g.set_params(p)
    row = [f"{wl} B={hb0.B:5d}"]
    for name, solver, k in (("spec auto", 2, 0), ("spec4", 2, 4), ("spec8", 2, 8), ("bcr", 1, 0), ("seq", 0, 0)):
        if solver == 1 and hb0.n_cap > 256:
            continue
        g.set_solver(solver)
        g.set_speculation(k)
        for _ in range(2):
            g.optimize(hb0.copy(), args)
        ts = []
        for _ in range(5):
            h = hb0.copy()
            t0 = time.perf_counter()
            g.optimize(h, args)
            ts.append(time.perf_counter() - t0)
        row.append(f"{name} {1e3 * min(ts):7.2f} ms")
    g.close()
    print(" | ".join(row), flush=True)
