"""Seeded synthetic planning scenes (SURVEY.md §8d / BASELINE.md §3).

A *request* is one planning problem (start, goal, obstacle table); a *candidate* is one initial band
of that request in its own homotopy-like class (lateral sinusoid of amplitude A_k). The batch axis of
the optimizer is candidates x requests.
"""
import numpy as np

from . import abi


def demo_scene_obstacles():
    """The three obstacles of test_optim_node (reference src/test_optim_node.cpp:106-117)."""
    o = np.zeros(3, abi.OBST_DTYPE)
    o["x"], o["y"] = [-3.0, 6.0, 0.0], [1.0, 2.0, 0.1]
    o["vx"], o["vy"] = [0.1, -0.3, 0.0], [-0.3, -0.2, 0.0]
    o["dynamic"] = [1, 1, 0]
    return o


def make_band(n, amplitude, max_vel_x=0.4, dt_ref=0.3, length=None):
    """Straight line start(-L/2,0,0) -> goal(L/2,0,0) plus lateral offset A*sin(pi s)."""
    L = max_vel_x * dt_ref * (n - 1) if length is None else length
    s = np.linspace(0.0, 1.0, n)
    x = -L / 2 + L * s
    y = amplitude * np.sin(np.pi * s)
    th = np.zeros(n)
    th[1:-1] = np.arctan2(y[2:] - y[1:-1], x[2:] - x[1:-1])
    seg = np.hypot(np.diff(x), np.diff(y))
    rec = np.zeros((n, 4))
    rec[:, 0], rec[:, 1], rec[:, 2] = x, y, th
    rec[:-1, 3] = seg / max_vel_x
    return rec


def make_obstacles(rng, M, L, inflated=False, moving=False, M_cap=None):
    M_cap = M if M_cap is None else M_cap
    o = np.zeros(M_cap, abi.OBST_DTYPE)
    k = 0
    lo, hi = -L / 2 + 1.0, L / 2 - 1.0
    if hi <= lo:
        lo, hi = -L / 2, L / 2
    while k < M:
        x, y = rng.uniform(lo, hi), rng.uniform(-3.0, 3.0)
        if min(np.hypot(x + L / 2, y), np.hypot(x - L / 2, y)) < 0.3:
            continue
        o[k]["x"], o[k]["y"] = x, y
        if inflated:
            o[k]["radius"] = rng.uniform(0.0, 0.3)
            o[k]["type"] = abi.TEB_OBST_CIRCULAR
        if moving:
            o[k]["vx"], o[k]["vy"] = rng.uniform(-0.3, 0.3, 2)
            o[k]["dynamic"] = 1
        k += 1
    return o


def make_batch(n, M, candidates, requests=1, seed=0, inflated=False, moving=False, via_points=0,
               n_cap=None, amp=2.0, max_vel_x=0.4, dt_ref=0.3):
    """Build a HostBatch of `candidates * requests` bands; scene s = request s."""
    n_cap = n if n_cap is None else n_cap
    B = candidates * requests
    L = max_vel_x * dt_ref * (n - 1)
    poses = np.zeros((B, n_cap, 4))
    obst = np.zeros((requests, max(M, 1)), abi.OBST_DTYPE)
    via = np.zeros((B, via_points, 2)) if via_points else None
    via_count = np.full(B, via_points, np.int32) if via_points else None
    scene_id = np.zeros(B, np.int32)
    for r in range(requests):
        rng = np.random.default_rng(seed * 100003 + r)
        obst[r, :M] = make_obstacles(rng, M, L, inflated, moving)[:M]
        for k in range(candidates):
            b = r * candidates + k
            brng = np.random.default_rng((seed * 100003 + r) * 1000003 + k)
            A = brng.uniform(-amp, amp)
            poses[b, :n] = make_band(n, A, max_vel_x, dt_ref)
            scene_id[b] = r
            if via_points:
                sv = np.linspace(0.2, 0.8, via_points)
                via[b, :, 0] = -L / 2 + L * sv
                via[b, :, 1] = A * np.sin(np.pi * sv)
    hb = abi.HostBatch(poses, np.full(B, n, np.int32), obst, np.full(requests, M, np.int32), scene_id,
                       via, via_count)
    return hb


def config_params(name):
    """TebParams for the BASELINE.md §3 configs (ctor defaults + per-config overrides)."""
    p = abi.default_params()
    if name == "C1":
        p.no_inner_iterations, p.no_outer_iterations = 4, 3
    elif name == "C2":
        p.min_turning_radius = 0.0
    elif name == "C3":
        p.min_turning_radius = 0.5
    elif name == "C4":
        p.include_dynamic_obstacles = 1
        p.weight_viapoint = 1.0
    else:
        raise ValueError(name)
    return p


CONFIG_SHAPES = {
    # name: (n poses, M obstacles, candidates, inflated, moving, via-points)
    "C1": (50, 5, 1, False, False, 0),
    "C2": (100, 20, 32, False, False, 0),
    "C3": (200, 64, 128, True, False, 0),
    "C4": (150, 32, 512, False, True, 4),
}


def make_config_batch(name, requests=1, seed=0, candidates=None, autosize=False):
    n, M, cand, inflated, moving, via_pts = CONFIG_SHAPES[name]
    cand = cand if candidates is None else candidates
    p = config_params(name)
    p.teb_autosize = int(autosize)
    n_cap = n if not autosize else min(2 * n, 512)
    hb = make_batch(n, M, cand, requests, seed, inflated, moving, via_pts, n_cap=n_cap)
    return p, hb


def polygon_centroid(v):
    """PolygonObstacle::calcCentroid (reference src/obstacles.cpp:56-119) for non-degenerate polygons."""
    v = np.asarray(v, float)
    if len(v) == 1:
        return v[0].copy()
    if len(v) == 2:
        return 0.5 * (v[0] + v[1])
    w = np.roll(v, -1, axis=0)
    aux = v[:, 0] * w[:, 1] - w[:, 0] * v[:, 1]
    A = 0.5 * aux.sum()
    return ((v + w) * aux[:, None]).sum(0) / (6 * A)


def add_shape_obstacles(hb, seed=0, kinds=(abi.TEB_OBST_LINE, abi.TEB_OBST_PILL, abi.TEB_OBST_POLYGON), every=2):
    """Turn every `every`-th obstacle of each scene into a Line / Pill / Polygon obstacle around its old position
    (vertices go to the scene's vertex pool, (x, y) becomes getCentroid()). Returns a new HostBatch."""
    obst = hb.obstacles.copy()
    pools = []
    for s in range(hb.S):
        rng = np.random.default_rng(seed * 7919 + s)
        pool = []
        for m in range(0, int(hb.obst_count[s]), every):
            kind = kinds[(m // every) % len(kinds)]
            c = np.array([obst[s, m]["x"], obst[s, m]["y"]])
            if kind == abi.TEB_OBST_POLYGON:
                nv = int(rng.integers(3, 6))
                ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
                v = c + np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.15, 0.5, (nv, 1))
            else:
                a = rng.uniform(0, np.pi)
                h = rng.uniform(0.2, 0.6)
                v = np.array([c - h * np.array([np.cos(a), np.sin(a)]), c + h * np.array([np.cos(a), np.sin(a)])])
            ctr = polygon_centroid(v)
            obst[s, m]["x"], obst[s, m]["y"] = ctr
            obst[s, m]["type"] = kind
            obst[s, m]["radius"] = rng.uniform(0.05, 0.25) if kind == abi.TEB_OBST_PILL else 0.0
            obst[s, m]["vertex_begin"], obst[s, m]["vertex_count"] = len(pool), len(v)
            pool.extend(v.tolist())
        pools.append(pool)
    pv = max(1, max(len(p) for p in pools))
    verts = np.zeros((hb.S, pv, 2))
    for s, pool in enumerate(pools):
        if pool:
            verts[s, :len(pool)] = pool
    return abi.HostBatch(hb.poses.copy(), hb.n.copy(), obst, hb.obst_count.copy(), hb.scene_id.copy(),
                         hb.via.copy() if hb.V_cap > 0 else None, hb.via_count.copy() if hb.V_cap > 0 else None,
                         hb.vel_start.copy(), hb.vel_goal.copy(), hb.prefer_rotdir.copy(), verts)


def set_line_footprint(p, start=(-0.3, 0.0), end=(0.4, 0.0)):
    p.footprint_type = abi.TEB_FOOTPRINT_LINE
    p.footprint_line[0], p.footprint_line[1], p.footprint_line[2], p.footprint_line[3] = start[0], start[1], end[0], end[1]
    return p


def set_polygon_footprint(p, vertices=((-0.25, -0.2), (0.35, -0.2), (0.35, 0.2), (-0.25, 0.2))):
    p.footprint_type = abi.TEB_FOOTPRINT_POLYGON
    p.footprint_vertex_count = len(vertices)
    for k, (x, y) in enumerate(vertices):
        p.footprint_vertices[2 * k], p.footprint_vertices[2 * k + 1] = x, y
    return p
