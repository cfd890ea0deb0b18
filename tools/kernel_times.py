"""Per-kernel CUDA-event times of one workload (profiled pass of the library): python tools/kernel_times.py [C3] [requests] [reps]
Environment switches of the library (TEBGPU_*) apply. Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
req = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, M, cand, inflated, moving, via_pts = scenes.CONFIG_SHAPES[wl]
cand = 32
p = scenes.config_params(wl)
p.teb_autosize = 0
hb = scenes.make_batch(n, M, cand, req, seed=1000, inflated=inflated, moving=moving, via_points=via_pts)
args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale, p.selection_viapoint_cost_scale, False)
g = T.TebGpu(hb.B, hb.n_cap, hb.S, max(hb.M_cap, 1), hb.V_cap)
g.set_params(p)
for k in sys.argv[4:]:
    key, val = k.split("=")
    getattr(g, key)(int(val))
h = hb.copy(); g.optimize(h, args)          # warm-up
import time
t0 = time.perf_counter()
for _ in range(reps):
    h = hb.copy(); g.optimize(h, args)
wall = (time.perf_counter() - t0) / reps
g.set_profiling(True)
for _ in range(reps):
    h = hb.copy(); g.optimize(h, args)
kt = g.kernel_times()
out = {k: {"avg_ms": v[0] / v[1], "launches_per_call": v[1] / reps, "ms_per_call": v[0] / reps} for k, v in kt.items() if v[1]}
print(json.dumps({"workload": wl, "bands": hb.B, "e2e_ms_per_call_unprofiled": wall * 1e3, "K": g.speculation_width(), "kernels": out,
                  "cost_checksum": float(np.nansum(h.cost[np.isfinite(h.cost)]))}))
g.close()
