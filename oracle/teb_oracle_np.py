"""Second, independent restatement of the reference TEB optimisation path in numpy (TEST INFRASTRUCTURE ONLY).

Purpose: cross-check oracle/teb_oracle.c on small bands (n <= ~30). It is written differently on purpose:
one stacked weighted residual vector r(x) over the whole band, a dense numeric Jacobian of that vector
(central differences, delta = 1e-9, through the same [+] operator), dense normal equations, numpy Cholesky,
and the g2o Levenberg-Marquardt schedule. PARITY UNPINNED, see oracle/teb_oracle.h.

Reference bodies followed (relative to the reference checkout):
  g2o_types/penalties.h:57-117, misc.h:95, g2o_types/edge_velocity.h:97-114, edge_acceleration.h:103-145,
  :316-341, :408-433, edge_kinematics.h:94-101, :203-215, edge_obstacle.h:90-103, :212-229,
  edge_dynamic_obstacle.h:98-101, edge_via_point.h:86, edge_time_optimal.h:93,
  src/optimal_planner.cpp:182-231 (outer loop), :444-548 (association), :646-718, :1041-1094 (cost),
  edge_velocity.h:236-273 and edge_acceleration.h:487-712 (holonomic variants), h_signature.h:97-186, :282-353.
Supported subset: point / circular footprint, point / circular obstacles, no autoResize.
"""
import math

import numpy as np

PI = math.pi


def nt(t):
    if -PI <= t < PI:
        return t
    m = math.floor(t / (2 * PI))
    t = t - m * 2 * PI
    if t >= PI:
        t -= 2 * PI
    if t < -PI:
        t += 2 * PI
    return t


def p_int(v, a, eps):
    if v < -a + eps:
        return -v - (a - eps)
    if v <= a - eps:
        return 0.0
    return v - (a - eps)


def p_int2(v, a, b, eps):
    if v < a + eps:
        return -v + (a + eps)
    if v <= b - eps:
        return 0.0
    return v - (b - eps)


def p_below(v, a, eps):
    return 0.0 if v >= a + eps else -v + (a + eps)


def seg(c, pa, pb, dt):
    dx, dy = pb[0] - pa[0], pb[1] - pa[1]
    dist = math.sqrt(dx * dx + dy * dy)
    ad = nt(pb[2] - pa[2])
    if c.exact_arc_length and ad != 0:
        dist = abs(ad * (dist / (2 * math.sin(ad / 2))))
    u = 100 * (dx * math.cos(pa[2]) + dy * math.sin(pa[2]))
    return dist / dt * (u / (1 + abs(u))), ad / dt


def fp_dist(c, pose, ox, oy, orad):
    d = math.hypot(pose[0] - ox, pose[1] - oy) - orad
    if c.footprint_type == 1:
        d -= c.footprint_radius
    return d


class Band:
    def __init__(self, c, rec, obst, via=None, vel_start=(0, 0, 0, 1), vel_goal=(0, 0, 0, 1)):
        self.c, self.rec, self.obst = c, np.array(rec, dtype=float), obst
        self.via = np.zeros((0, 2)) if via is None else np.asarray(via, dtype=float)
        self.vs, self.vg = vel_start, vel_goal
        self.n = len(self.rec)

    # ---- graph structure frozen per outer iteration
    def build(self, mult):
        c, n, P = self.c, self.n, self.rec
        self.mult = mult
        self.assoc = [[] for _ in range(n)]
        if c.weight_obstacle != 0 and mult != 0:
            for i in range(1, n - 1):
                left = right = None
                lmin = rmin = float("inf")
                ox, oy = math.cos(P[i, 2]), math.sin(P[i, 2])
                for m, ob in enumerate(self.obst):
                    if c.include_dynamic_obstacles and ob["dynamic"]:
                        continue
                    d = fp_dist(c, P[i], ob["x"], ob["y"], ob["radius"])
                    if d < c.min_obstacle_dist * c.obstacle_association_force_inclusion_factor:
                        self.assoc[i].append(m)
                        continue
                    if d > c.min_obstacle_dist * c.obstacle_association_cutoff_factor:
                        continue
                    if ox * (ob["y"] - P[i, 1]) - (ob["x"] - P[i, 0]) * oy > 0:
                        if d < lmin:
                            lmin, left = d, m
                    elif d < rmin:
                        rmin, right = d, m
                if left is not None:
                    self.assoc[i].append(left)
                if right is not None:
                    self.assoc[i].append(right)
        self.dyn_t = np.zeros(n)
        t = P[0, 3]
        for i in range(1, n - 1):
            self.dyn_t[i] = t
            t += P[i, 3]
        self.via_idx = []
        if c.weight_viapoint != 0 and n >= 3:
            for v in self.via:
                d2 = (P[:, 0] - v[0]) ** 2 + (P[:, 1] - v[1]) ** 2
                idx = int(np.argmin(d2))
                idx = min(idx, n - 2)
                if idx < 1:
                    continue
                self.via_idx.append((idx, v))

    # ---- stacked weighted residuals, tagged by family (0 obstacle, 1 via, 2 time, 3 other)
    def residuals(self, P):
        c, n = self.c, self.n
        r, tag = [], []

        def add(e, w, t):
            r.append(math.sqrt(w) * e)
            tag.append(t)

        infl = c.inflation_dist > c.min_obstacle_dist
        for i in range(1, n - 1):
            for m in self.assoc[i]:
                ob = self.obst[m]
                d = fp_dist(c, P[i], ob["x"], ob["y"], ob["radius"])
                add(p_below(d, c.min_obstacle_dist, c.penalty_epsilon), c.weight_obstacle * self.mult, 0)
                if infl:
                    add(p_below(d, c.inflation_dist, 0.0), c.weight_inflation, 0)
        if c.include_dynamic_obstacles and c.weight_obstacle != 0:
            for ob in self.obst:
                if not ob["dynamic"]:
                    continue
                for i in range(1, n - 1):
                    t = self.dyn_t[i]
                    d = fp_dist(c, P[i], ob["x"] + t * ob["vx"], ob["y"] + t * ob["vy"], ob["radius"])
                    add(p_below(d, c.min_obstacle_dist, c.penalty_epsilon), c.weight_dynamic_obstacle, 0)
                    add(p_below(d, c.dynamic_obstacle_inflation_dist, 0.0), c.weight_dynamic_obstacle_inflation, 0)
        for idx, v in self.via_idx:
            add(math.hypot(P[idx, 0] - v[0], P[idx, 1] - v[1]), c.weight_viapoint, 1)
        vw = [seg(c, P[i], P[i + 1], P[i, 3]) for i in range(n - 1)]
        holo = c.max_vel_y != 0

        def holo_vel(a, b, dt):          # velocity in the frame of the first pose (edge_velocity.h:247-254)
            dx, dy = b[0] - a[0], b[1] - a[1]
            cs, sn = math.cos(a[2]), math.sin(a[2])
            return (cs * dx + sn * dy) / dt, (-sn * dx + cs * dy) / dt, nt(b[2] - a[2]) / dt

        hv = [holo_vel(P[i], P[i + 1], P[i, 3]) for i in range(n - 1)] if holo else None
        if holo:
            for i in range(n - 1):
                vx, vy, om = hv[i]
                rem_y = math.sqrt(max(0.0, c.max_vel_trans ** 2 - vx * vx))
                rem_x = math.sqrt(max(0.0, c.max_vel_trans ** 2 - vy * vy))
                add(p_int2(vx, -min(rem_x, c.max_vel_x_backwards), min(rem_x, c.max_vel_x), 0.0), c.weight_max_vel_x, 3)
                add(p_int(vy, min(rem_y, c.max_vel_y), 0.0), c.weight_max_vel_y, 3)
                add(p_int(om, c.max_vel_theta, c.penalty_epsilon), c.weight_max_vel_theta, 3)
        else:
            for i in range(n - 1):
                add(p_int2(vw[i][0], -c.max_vel_x_backwards, c.max_vel_x, c.penalty_epsilon), c.weight_max_vel_x, 3)
                add(p_int(vw[i][1], c.max_vel_theta, c.penalty_epsilon), c.weight_max_vel_theta, 3)
        if holo and c.acc_lim_y != 0:    # EdgeAccelerationHolonomic / Start / Goal (edge_acceleration.h:487-712)
            lim = (c.acc_lim_x, c.acc_lim_y, c.acc_lim_theta)
            wts = (c.weight_acc_lim_x, c.weight_acc_lim_y, c.weight_acc_lim_theta)
            if self.vs[3]:
                for k in range(3):
                    add(p_int((hv[0][k] - self.vs[k]) / P[0, 3], lim[k], c.penalty_epsilon), wts[k], 3)
            for i in range(n - 2):
                T = P[i, 3] + P[i + 1, 3]
                for k in range(3):
                    add(p_int((hv[i + 1][k] - hv[i][k]) * 2 / T, lim[k], c.penalty_epsilon), wts[k], 3)
            if self.vg[3]:
                for k in range(3):
                    add(p_int((self.vg[k] - hv[n - 2][k]) / P[n - 2, 3], lim[k], c.penalty_epsilon), wts[k], 3)
            return self._tail(P, r, tag, add)
        if self.vs[3]:
            dt = P[0, 3]
            add(p_int((vw[0][0] - self.vs[0]) / dt, c.acc_lim_x, c.penalty_epsilon), c.weight_acc_lim_x, 3)
            add(p_int((vw[0][1] - self.vs[2]) / dt, c.acc_lim_theta, c.penalty_epsilon), c.weight_acc_lim_theta, 3)
        for i in range(n - 2):
            T = P[i, 3] + P[i + 1, 3]
            add(p_int((vw[i + 1][0] - vw[i][0]) * 2 / T, c.acc_lim_x, c.penalty_epsilon), c.weight_acc_lim_x, 3)
            add(p_int((vw[i + 1][1] - vw[i][1]) * 2 / T, c.acc_lim_theta, c.penalty_epsilon), c.weight_acc_lim_theta, 3)
        if self.vg[3]:
            dt = P[n - 2, 3]
            add(p_int((self.vg[0] - vw[n - 2][0]) / dt, c.acc_lim_x, c.penalty_epsilon), c.weight_acc_lim_x, 3)
            add(p_int((self.vg[2] - vw[n - 2][1]) / dt, c.acc_lim_theta, c.penalty_epsilon), c.weight_acc_lim_theta, 3)
        return self._tail(P, r, tag, add)

    def _tail(self, P, r, tag, add):
        c, n = self.c, self.n
        for i in range(n - 1):
            add(P[i, 3], c.weight_optimaltime, 2)
        carlike = not (c.min_turning_radius == 0 or c.weight_kinematics_turning_radius == 0)
        for i in range(n - 1):
            a, b = P[i], P[i + 1]
            dx, dy = b[0] - a[0], b[1] - a[1]
            add(abs((math.cos(a[2]) + math.cos(b[2])) * dy - (math.sin(a[2]) + math.sin(b[2])) * dx),
                c.weight_kinematics_nh, 3)
            if not carlike:
                add(p_below(dx * math.cos(a[2]) + dy * math.sin(a[2]), 0, 0), c.weight_kinematics_forward_drive, 3)
            else:
                ad = nt(b[2] - a[2])
                e = 0.0 if ad == 0 else p_below(math.hypot(dx, dy) / abs(ad), c.min_turning_radius, 0.0)
                add(e, c.weight_kinematics_turning_radius, 3)
        return np.array(r), np.array(tag)

    # ---- free-variable bookkeeping: g2o order dt_0, pose_1, dt_1, ...
    def free_index(self):
        idx = []
        for i in range(self.n - 1):
            if i >= 1:
                idx += [(i, 0), (i, 1), (i, 2)]
            idx.append((i, 3))
        return idx

    def oplus(self, P, idx, dx):
        Q = P.copy()
        for (i, cidx), d in zip(idx, dx):
            Q[i, cidx] = nt(Q[i, cidx] + d) if cidx == 2 else Q[i, cidx] + d
        return Q

    def lm(self, inner):
        idx = self.free_index()
        N = len(idx)
        lam, ni = 0.0, 2.0
        last_r, last_tag = None, None
        for it in range(inner):
            r0, tag = self.residuals(self.rec)
            last_r, last_tag = r0, tag
            cur = float(r0 @ r0)
            J = np.zeros((len(r0), N))
            for k in range(N):
                d = np.zeros(N)
                d[k] = 1e-9
                rp, _ = self.residuals(self.oplus(self.rec, idx, d))
                rm, _ = self.residuals(self.oplus(self.rec, idx, -d))
                J[:, k] = (rp - rm) / 2e-9
            H, b = J.T @ J, -J.T @ r0
            if it == 0:
                lam, ni = 1e-5 * np.abs(np.diag(H)).max(), 2.0
            rho, q = 0.0, 0
            while True:
                try:
                    L = np.linalg.cholesky(H + lam * np.eye(N))
                    dx = np.linalg.solve(L.T, np.linalg.solve(L, b))
                    ok = True
                except np.linalg.LinAlgError:
                    dx, ok = b.copy(), False
                trial = self.oplus(self.rec, idx, dx)
                rt, tagt = self.residuals(trial)
                last_r, last_tag = rt, tagt
                tmp = float(rt @ rt) if ok else float("inf")
                scale = float(dx @ (lam * dx + b)) + 1e-3
                rho = (cur - tmp) / scale
                if rho > 0 and math.isfinite(tmp):
                    alpha = min(1 - (2 * rho - 1) ** 3, 2 / 3)
                    lam *= max(1 / 3, alpha)
                    ni = 2.0
                    cur = tmp
                    self.rec = trial
                else:
                    lam *= ni
                    ni *= 2
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0 or not math.isfinite(lam):
                break
        return last_r, last_tag


def optimize(c, rec, obst, via=None, inner=5, outer=4, obst_scale=1.0, via_scale=1.0, vel_start=(0, 0, 0, 1),
             vel_goal=(0, 0, 0, 1)):
    """optimizeTEB without autoResize; returns (rec, cost)."""
    band = Band(c, rec, obst, via, vel_start, vel_goal)
    mult, cost = 1.0, float("inf")
    for o in range(outer):
        band.build(mult)
        r, tag = band.lm(inner)
        if o == outer - 1:
            sq = r * r
            cost = obst_scale * sq[tag == 0].sum() + via_scale * sq[tag == 1].sum() + sq[tag == 2].sum() + sq[tag == 3].sum()
        mult *= c.weight_adapt_factor
    return band.rec, float(cost)


# ---------------------------------------------------------------------------------------------------------------------
# H-signatures (h_signature.h), independent of oracle/teb_oracle.c: numpy longdouble / vector algebra
def h_signature_2d(c, rec, obst):
    """HSignature::calculateHSignature h_signature.h:97-186"""
    M = len(obst)
    if M == 0:
        return 0j
    ld = np.longdouble
    z = np.array([complex(p[0], p[1]) for p in rec], dtype=np.clongdouble)
    o = np.array([complex(ob["x"], ob["y"]) for ob in obst], dtype=np.clongdouble)
    m = max(M - 1, 5)
    a = int(math.ceil(m / 2.0))
    b = m - a
    delta = z[-1] - z[0]
    normal = np.clongdouble(complex(-float(delta.imag), float(delta.real)))
    if abs(delta) < 3.0:
        bl, tr = z[0] + np.clongdouble(-3j), z[0] + np.clongdouble(3 + 3j)
    else:
        bl, tr = z[0] - normal, z[0] + delta + normal
    H = np.clongdouble(0)
    A = []
    for l in range(M):
        Al = ld(c.h_signature_prescaler) * ld(a) * (o[l] - bl) * ld(b) * (o[l] - tr)
        for j in range(M):
            if j == l:
                continue
            diff = o[l] - o[j]
            if abs(diff) < 0.05:
                continue
            Al = Al / diff
        A.append(Al)
    for k in range(len(z) - 1):
        for l in range(M):
            d2, d1 = float(abs(z[k + 1] - o[l])), float(abs(z[k] - o[l]))
            if d2 == 0 or d1 == 0:
                continue
            log_real = math.log(d2) - math.log(d1)
            w2, w1 = z[k + 1] - o[l], z[k] - o[l]
            arg_diff = float(np.arctan2(w2.imag, w2.real) - np.arctan2(w1.imag, w1.real))
            cand = [arg_diff, arg_diff + 2 * PI, arg_diff - 2 * PI, arg_diff + 4 * PI, arg_diff - 4 * PI]
            log_imag = min(cand, key=abs)
            H = H + A[l] * np.clongdouble(complex(log_real, log_imag))
    return complex(H)


def h_signature_3d(c, rec, obst, use_timediffs=True):
    """HSignature3d::calculateHSignature h_signature.h:282-353 (Biot-Savart integral, 10 steps per segment)"""
    rec = np.asarray(rec, float)
    t = np.zeros(len(rec))
    for k in range(len(rec) - 1):
        t[k + 1] = t[k] + (rec[k, 3] if use_timediffs else math.hypot(*(rec[k + 1, :2] - rec[k, :2])) / c.max_vel_x)
    out = []
    for ob in obst:
        s1 = np.array([ob["x"], ob["y"], 0.0])
        s2 = np.array([ob["x"] + 120 * ob["vx"], ob["y"] + 120 * ob["vy"], 120.0])
        ds = s2 - s1
        H = 0.0
        for k in range(len(rec) - 1):
            direction = np.array([rec[k + 1, 0] - rec[k, 0], rec[k + 1, 1] - rec[k, 1], t[k + 1] - t[k]])
            if np.linalg.norm(direction) < 1e-15:
                continue
            r = np.array([rec[k, 0], rec[k, 1], t[k]])
            dl = 0.1 * direction
            for _ in range(10):
                p1, p2 = s1 - r, s2 - r
                d = np.cross(ds, np.cross(p1, p2)) / ds.dot(ds)
                phi = (np.cross(d, p2) / np.linalg.norm(p2) - np.cross(d, p1) / np.linalg.norm(p1)) / d.dot(d)
                H += phi.dot(dl)
                r = r + dl
        out.append(H / (4 * PI))
    return np.array(out)
