"""BASELINE config C5: sweep batch 64-4096 x poses 50-400 x obstacles 8-256 on one GPU; trajectories/s, LM iterations/s and
the HBM GB/s of kernel A (algorithmic bytes of SURVEY 8(d) / CUDA-event launch time) against the measured peak.
usage: python tools/c5_sweep.py [out.json]   (rank-local; run under gpurun)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_c5_sweep.json"
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    peak = 6650.0
rows = []
for carlike in (False, True):
    for n in (50, 100, 200, 400):
        for M in (8, 64, 256):
            for B in (64, 512, 4096):
                if carlike and (M == 256 or n == 50):
                    continue   # half grid for the second kinematics
                p = scenes.config_params("C3" if carlike else "C2")
                p.teb_autosize = 0
                cand = 32
                hb = scenes.make_batch(n, M, cand, B // cand, seed=77, inflated=carlike)
                args = abi.make_args(5, 4, True, 100.0, 1.0, False)
                g = T.TebGpu(hb.B, hb.n_cap, hb.S, max(hb.M_cap, 1), hb.V_cap)
                g.set_params(p)
                g.set_graph(0)
                h = hb.copy(); g.optimize(h, args)
                reps = 3 if B >= 4096 else 6
                import torch
                ev = []
                for _ in range(reps):
                    h = hb.copy()
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    g.optimize(h, args)
                    torch.cuda.synchronize(); ev.append(time.perf_counter() - t0)
                sec = float(np.median(ev))
                g.set_profiling(True)
                h2 = hb.copy(); g.optimize(h2, args)
                kt = g.kernel_times()
                a_ms = kt["k_linearize"][0] / max(1, kt["k_linearize"][1])
                N = 4 * n - 7
                bytes_a = (32 * n + 64 * M + 8 * 12 * N + 8) * hb.B
                ws_mb = hb.B * 4 * n * 96 / 1e6
                rows.append({"kinematics": "car-like" if carlike else "diff-drive", "bands": hb.B, "poses": n, "obstacles": M,
                             "ms_per_call_host_api": sec * 1e3, "trajectories_per_s": hb.B / sec,
                             "lm_iterations_per_s": float(h.lm_iters.sum()) / sec, "kernel_a_ms": a_ms,
                             "kernel_a_gbs": bytes_a / (a_ms * 1e-3) / 1e9, "kernel_a_frac_of_measured_peak": bytes_a / (a_ms * 1e-3) / 1e9 / peak,
                             "working_set_mb": ws_mb, "regime": "HBM" if ws_mb > 126 else "L2-resident / latency-bound"})
                g.close()
                print(json.dumps(rows[-1]), flush=True)
json.dump({"peak_gbs": peak, "note": "one B200, host-buffer API (H2D + D2H inside), 4 x 5 LM iterations, teb_autosize off", "rows": rows},
          open(out_path, "w"), indent=1)
