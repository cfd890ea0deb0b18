"""Dev probe (run under gpurun): kernel-A system vs oracle, full optimize vs oracle, on small batches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes
from tests import oracle_binding as ob

def padded_from_oracle(Hd, bd, n):
    """oracle dense (g2o order, N=4n-7) -> padded band layout [4n][12]"""
    N = 4*n-7
    Hb = np.zeros((4*n, 12))
    for r in range(4*n):
        Hb[r,0] = 1.0
    for r in range(N):
        R = r+3
        for k in range(11):
            q = r-k
            if q < 0: break
            Hb[R,k] = Hd[r,q]
        Hb[R,11] = bd[r]
    return Hb

def main():
    for cfg in ("C1","C2","C3","C4"):
        p, hb = scenes.make_config_batch(cfg, candidates=6, autosize=False)
        g = T.TebGpu(hb.B, hb.n_cap, hb.S, hb.M_cap, hb.V_cap)
        g.set_params(p)
        for outer_index in (0, 2):
            Hb, chi2 = g.build_system(hb, outer_index)
            worstH = worstb = 0; worstc=0
            for b in range(hb.B):
                n = hb.n[b]
                Hd, bd, c2 = ob.build_system(p, hb.poses[b], n, hb.obstacles[0][:hb.obst_count[0]],
                                             via=hb.via[b] if hb.V_cap else None, weight_multiplier=2.0**outer_index, jac_mode=1)
                ref = padded_from_oracle(Hd, bd, n)
                got = Hb[b,:4*n]
                sH = np.abs(ref[:,:11]).max(); sb = np.abs(ref[:,11]).max()
                dH = np.abs(got[:,:11]-ref[:,:11]).max()/sH; dbb = np.abs(got[:,11]-ref[:,11]).max()/max(sb,1e-300)
                worstH=max(worstH,dH); worstb=max(worstb,dbb); worstc=max(worstc,abs(chi2[b]-c2)/max(c2,1e-300))
                if dH > 1e-9 or dbb > 1e-9:
                    idx = np.unravel_index(np.abs(got-ref).argmax(), got.shape)
                    print("  MISMATCH", cfg, "band", b, "at row,col", idx, "got", got[idx], "ref", ref[idx])
            print(cfg, "outer", outer_index, "kernelA rel err H %.2e b %.2e chi2 %.2e" % (worstH, worstb, worstc))
        args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, 100.0, 1.0, False)
        for autosize in (False, True):
            p.teb_autosize = int(autosize)
            _, hb0 = scenes.make_config_batch(cfg, candidates=6, autosize=autosize)
            g.close()
            g = T.TebGpu(hb0.B, hb0.n_cap, hb0.S, hb0.M_cap, hb0.V_cap); g.set_params(p)
            hg = hb0.copy(); t=time.time(); g.optimize(hg, args); tg=time.time()-t
            ha = hb0.copy(); ob.optimize_batch(p, ha, args, jac_mode=1)
            hn = hb0.copy(); t=time.time(); ob.optimize_batch(p, hn, args, jac_mode=0); tc=time.time()-t
            for name, href in (("analytic", ha), ("g2o-numeric", hn)):
                same_n = (hg.n == href.n)
                d = [np.abs(hg.poses[b,:hg.n[b]] - href.poses[b,:href.n[b]]).max() if same_n[b] else np.inf for b in range(hg.B)]
                dc = np.abs(hg.cost-href.cost)/np.maximum(np.abs(href.cost),1e-300)
                print(cfg, "autosize", autosize, "vs", name, "pose diff", ["%.1e"%x for x in d], "cost rel", "%.1e"%dc.max(),
                      "n", hg.n.tolist(), href.n.tolist(), "iters", hg.lm_iters.tolist(), href.lm_iters.tolist(), "status", hg.status.tolist(), href.status.tolist())
            print("   gpu %.1f ms (incl copies), cpu g2o-mode 1 thread %.1f ms, launches %d" % (tg*1e3, tc*1e3, g.launch_count()))
        g.close()

if __name__ == "__main__":
    main()
