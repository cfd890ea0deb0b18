"""CPU restatement of the candidate exploration of HomotopyClassPlanner (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Follows the reference line by line, sequentially and without batching:
  * lrKeyPointGraph::createGraph        src/graph_search.cpp:92-216
  * ProbRoadmapGraph::createGraph       src/graph_search.cpp:220-342 (boost::mt19937 default seed 5489 +
                                         boost::random::uniform_real_distribution<double>: draw / 2^32 * (b - a) + a)
  * GraphSearchInterface::DepthFirst    src/graph_search.cpp:45-88
  * addAndInitNewTeb (path version)     include/teb_local_planner/homotopy_class_planner.hpp:67-100
  * initTrajectoryToGoal (path version) include/teb_local_planner/timed_elastic_band.hpp:46-185
  * addEquivalenceClassIfNew, isEqual   src/homotopy_class_planner.cpp:178-211, h_signature.h:191-207, :366-388
H-signature values come from oracle/teb_oracle.c (teb_oracle_h_signature, long double). Pure Python: small cases only.
PARITY PINNED (tests/test_reference_pin.py::test_candidate_exploration_matches_the_reference_planner): the reference's own
homotopy_class_planner.cpp / graph_search.cpp / h_signature.h, compiled into oracle/_ref against stand-in headers, propose
the same candidates in the same order with the same initial bands (1e-12) on random scenes - both graphs, both signature
kinds, two cycles. Restated on BOTH sides (boost is not in the image): mt19937 is the standard engine, the
uniform_real_distribution follows boost's published one-draw algorithm. The pin caught one thing the restatement had
wrong: the roadmap draws its y sample before its x sample (unspecified argument evaluation order, right-to-left with GCC).
The Planner class below (whole plan() cycles) is pinned the same way (test_planning_cycles_match_the_reference_planner).
"""
import math

import numpy as np


class MT19937:
    """mt19937 with init_genrand(seed) seeding (boost::random::mt19937 / std::mt19937 default constructor: 5489)."""

    def __init__(self, seed=5489):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for k in range(624):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def uniform(self, lo, hi):
        while True:
            r = self() / 4294967296.0 * (hi - lo) + lo
            if r < hi:
                return r


def _segments_intersect(a0, a1, b0, b1):
    """check_line_segments_intersection_2d distance_calculations.h:97-127"""
    l1 = a1 - a0
    l2 = b1 - b0
    denom = l1[0] * l2[1] - l2[0] * l1[1]
    if denom == 0:
        return False
    pos = denom > 0
    aux = a0 - b0
    s_numer = l1[0] * aux[1] - l1[1] * aux[0]
    if (s_numer < 0) == pos:
        return False
    t_numer = l2[0] * aux[1] - l2[1] * aux[0]
    if (t_numer < 0) == pos:
        return False
    if ((s_numer > denom) == pos) or ((t_numer > denom) == pos):
        return False
    return True


class Obst:
    """kind: 'point' | 'circle' | 'line' | 'pill' | 'polygon'; centroid as the reference computes it"""

    def __init__(self, kind, centroid, radius=0.0, vertices=None):
        self.kind, self.c, self.r = kind, np.asarray(centroid, float), radius
        self.v = None if vertices is None else np.asarray(vertices, float)

    def check_line_intersection(self, a, b, min_dist):
        if self.kind in ("point", "circle"):   # obstacles.h:339-355, :483-499
            d = b - a
            t = d.dot(self.c - a) / d.dot(d)
            t = 0.0 if t < 0 else (1.0 if t > 1 else t)
            nearest = a + d * t
            return (np.linalg.norm(nearest - self.c) - self.r) < min_dist
        if self.kind in ("line", "pill"):      # :647-650, :794-797 (min_dist ignored)
            return _segments_intersect(a, b, self.v[0], self.v[1])
        k = len(self.v)                        # obstacles.cpp:176-191
        for i in range(k - 1):
            if _segments_intersect(a, b, self.v[i], self.v[i + 1]):
                return True
        if k == 2:
            return False
        return _segments_intersect(a, b, self.v[-1], self.v[0])


def _nt(t):
    if -math.pi <= t < math.pi:
        return t
    m = math.floor(t / (2 * math.pi))
    t = t - m * 2 * math.pi
    if t >= math.pi:
        t -= 2 * math.pi
    if t < -math.pi:
        t += 2 * math.pi
    return t


def init_from_path(path, max_vel_x, acc_lim_x, start_orient, goal_orient, min_samples, guess_backwards):
    """timed_elastic_band.hpp:46-185 -> records [n][4] (x, y, theta, dt)"""
    path = [np.asarray(p, float) for p in path]
    start, goal = path[0], path[-1]
    backwards = guess_backwards and (goal - start).dot(np.array([math.cos(start_orient), math.sin(start_orient)])) < 0

    def seg_time(length):
        tv = length / max_vel_x
        ta = math.sqrt(2 * length / acc_lim_x)
        return ta if tv < ta else tv

    poses = [[start[0], start[1], start_orient]]
    dts = []
    for k in range(1, len(path) - 1):
        diff = path[k] - np.array(poses[-1][:2])
        ts = seg_time(math.hypot(diff[0], diff[1]))
        if ts <= 0:
            ts = 0.2
        yaw = math.atan2(diff[1], diff[0])
        if backwards:
            yaw = _nt(yaw + math.pi)
        poses.append([path[k][0], path[k][1], yaw])
        dts.append(ts)
    d = goal - np.array(poses[-1][:2])
    ts = seg_time(math.hypot(d[0], d[1]))
    while len(poses) < min_samples - 1:
        ts /= 2
        b = poses[-1]
        th = math.atan2(math.sin(b[2]) + math.sin(goal_orient), math.cos(b[2]) + math.cos(goal_orient))
        poses.append([(b[0] + goal[0]) / 2, (b[1] + goal[1]) / 2, th])     # PoseSE2::average pose_se2.h:262-268
        dts.append(ts)
    poses.append([goal[0], goal[1], goal_orient])
    dts.append(ts)
    rec = np.zeros((len(poses), 4))
    rec[:, :3] = poses
    rec[:-1, 3] = dts
    return rec


class Explorer:
    """equivalence-class bookkeeping of a fresh HomotopyClassPlanner (no best band yet) + the graph searches"""

    def __init__(self, params, hcp, oracle_binding, obst_rows, obstacles):
        self.p, self.hcp, self.ob, self.rows, self.obstacles = params, hcp, oracle_binding, obst_rows, obstacles
        self.classes, self.tebs = [], []
        self.rng = MT19937()

    # --- equivalence classes
    def _is_equal(self, a, b):
        thr = self.p.h_signature_threshold
        if self.p.include_dynamic_obstacles:
            for x, y in zip(a, b):
                if abs(x) < thr or abs(y) < thr:
                    continue
                if np.sign(x) != np.sign(y):
                    return False
            return True
        return abs(a.real - b.real) <= thr and abs(a.imag - b.imag) <= thr

    def _add_if_new(self, h):
        valid = np.all(np.isfinite(h)) if self.p.include_dynamic_obstacles else (math.isfinite(h.real) and math.isfinite(h.imag))
        if not valid:
            return False
        if any(self._is_equal(h, c) for c in self.classes):
            return False           # no best band: isInBestTebClass is false
        self.classes.append(h)
        return True

    def add_and_init(self, path, start_orient, goal_orient):
        if len(self.tebs) >= self.hcp["max_number_classes"]:
            return
        rec = init_from_path(path, self.p.max_vel_x, self.p.acc_lim_x, start_orient, goal_orient, self.p.min_samples,
                             bool(self.p.allow_init_with_backwards_motion))
        h = self.ob.h_signature(self.p, rec, len(rec), self.rows, use_timediffs=True)
        if self._add_if_new(h):
            self.tebs.append(rec)

    # --- graph_search.cpp:45-88
    def _depth_first(self, pos, adj, visited, goal, so, go):
        if len(self.tebs) >= self.hcp["max_number_classes"]:
            return
        back = visited[-1]
        for v in adj[back]:
            if v in visited:
                continue
            if v == goal:
                self.add_and_init([pos[u] for u in visited] + [pos[goal]], so, go)
                break
        for v in adj[back]:
            if v in visited or v == goal:
                continue
            visited.append(v)
            self._depth_first(pos, adj, visited, goal, so, go)
            visited.pop()

    def _edges(self, pos, diff, thr, min_dist, start=None, nearest=None):
        nv = len(pos)
        adj = [[] for _ in range(nv)]
        for i in range(nv - 1):
            for j in range(nv):
                if i == j:
                    continue
                dij = pos[j] - pos[i]
                z = dij.dot(dij)
                if z > 0:
                    dij = dij / math.sqrt(z)
                if dij.dot(diff) <= thr:
                    continue
                if nearest is not None and thr and i == 0 and j in nearest:
                    kd = pos[j] - start[:2]
                    z = kd.dot(kd)
                    if z > 0:
                        kd = kd / math.sqrt(z)
                    if np.array([math.cos(start[2]), math.sin(start[2])]).dot(kd) <= thr:
                        continue
                if any(o.check_line_intersection(pos[i], pos[j], min_dist) for o in self.obstacles):
                    continue
                adj[i].append(j)
        return adj

    def lr_key_point_graph(self, start, goal, dist_to_obst):
        start, goal = np.asarray(start, float), np.asarray(goal, float)
        thr = self.hcp["obstacle_heading_threshold"]
        if len(self.tebs) >= self.hcp["max_number_classes"]:
            return
        diff = goal[:2] - start[:2]
        normal = np.array([-diff[1], diff[0]])
        normal = normal / math.sqrt(normal.dot(normal)) * dist_to_obst
        pos = [start[:2].copy()]
        diff = diff / math.sqrt(diff.dot(diff))
        nearest, min_dist = None, float("inf")
        for o in self.obstacles:
            s2o = o.c - start[:2]
            dist = math.sqrt(s2o.dot(s2o))
            if s2o.dot(diff) / dist < 0.1:
                continue
            pos.append(o.c + normal)
            pos.append(o.c - normal)
            if thr and dist < min_dist:
                min_dist, nearest = dist, (len(pos) - 2, len(pos) - 1)
        pos.append(goal[:2].copy())
        adj = self._edges(pos, diff, thr, 0.5 * dist_to_obst, start, nearest)
        self._depth_first(pos, adj, [0], len(pos) - 1, start[2], goal[2])

    def prob_roadmap_graph(self, start, goal, dist_to_obst):
        start, goal = np.asarray(start, float), np.asarray(goal, float)
        thr = self.hcp["obstacle_heading_threshold"]
        if len(self.tebs) >= self.hcp["max_number_classes"]:
            return
        diff = goal[:2] - start[:2]
        sg = math.sqrt(diff.dot(diff))
        normal = np.array([-diff[1], diff[0]])
        normal = normal / math.sqrt(normal.dot(normal))
        width, scale = self.hcp["roadmap_graph_area_width"], self.hcp["roadmap_graph_area_length_scale"]
        phi = math.atan2(diff[1], diff[0])
        c, s = math.cos(phi), math.sin(phi)
        if scale != 1.0:
            origin = start[:2] + 0.5 * (1.0 - scale) * sg * (diff / sg) - 0.5 * width * normal
        else:
            origin = start[:2] - 0.5 * width * normal
        pos = [start[:2].copy()]
        diff = diff / sg
        for _ in range(self.hcp["roadmap_graph_no_samples"]):
            # Eigen::Vector2d(distribution_x(gen), distribution_y(gen)) (graph_search.cpp:274): the order in which the two
            # arguments are evaluated is unspecified in C++; GCC - what the reference is built with - evaluates them right
            # to left, so the y sample is drawn FIRST (pinned against the reference compiled here, tests/test_reference_pin.py)
            sy = self.rng.uniform(0, width)
            sx = self.rng.uniform(0, sg * scale)
            pos.append(origin + np.array([c * sx - s * sy, s * sx + c * sy]))
        pos.append(goal[:2].copy())
        adj = self._edges(pos, diff, thr, dist_to_obst)
        self._depth_first(pos, adj, [0], len(pos) - 1, start[2], goal[2])


# ---------------------------------------------------------------------------------------------------------------------
# whole planning cycles of HomotopyClassPlanner::plan (src/homotopy_class_planner.cpp:107-125), sequential restatement:
# updateAllTEBs -> exploreEquivalenceClassesAndInitTebs (renewAndAnalyzeOldTebs, initial plan, graph search) ->
# optimizeAllTEBs -> selectBestTeb. The optimisation itself is oracle/teb_oracle.c (closed-form-Jacobian mode).
class Planner(Explorer):
    def __init__(self, params, hcp, oracle_binding, obst_rows, obstacles, simple_exploration=True):
        super().__init__(params, hcp, oracle_binding, obst_rows, obstacles)
        self.simple = simple_exploration
        self.obst_vertices = None                         # vertex pool of Line / Pill / Polygon obstacle rows (for the optimisation)
        self.jac_mode = None                              # None: closed-form Jacobians; oracle_binding.JAC_G2O = the reference's numeric ones
        self.costs, self.optimized = [], []
        self.best, self.best_class = None, None          # index into self.tebs / its class at classification time

    # homotopy_class_planner.cpp:189-211 with a best band present
    def _add_if_new(self, h):
        valid = np.all(np.isfinite(h)) if self.p.include_dynamic_obstacles else (math.isfinite(h.real) and math.isfinite(h.imag))
        if not valid:
            return False
        if any(self._is_equal(h, c) for c in self.classes):
            in_best = self.best_class is not None and self._is_equal(self.best_class, h)
            n_best = sum(1 for c in self.classes if self.best_class is not None and self._is_equal(self.best_class, c))
            if not in_best or n_best >= self.hcp.get("max_number_plans_in_current_class", 1):
                return False
        self.classes.append(h)
        return True

    def _signature(self, rec):
        return self.ob.h_signature(self.p, rec, len(rec), self.rows, use_timediffs=True)

    # :214-256 (detour deletion :766-801 included)
    def renew_and_analyze(self):
        self.classes = []
        order = list(range(len(self.tebs)))
        if self.best is not None:       # std::iter_swap(tebs_.begin(), it_best_teb): a swap, not a rotation
            order[0], order[self.best] = order[self.best], order[0]
        tebs = [self.tebs[i] for i in order]
        costs = [self.costs[i] for i in order]
        opt = [self.optimized[i] for i in order]
        keep = []
        if self.best is not None:
            self.best_class = self._signature(tebs[0])
            self._add_if_new(self.best_class)
            keep.append(0)
        for k in range(1 if self.best is not None else 0, len(tebs)):
            if self._add_if_new(self._signature(tebs[k])):
                keep.append(k)
            # else: band dropped, its class is taken
        self.tebs, self.costs, self.optimized = [tebs[k] for k in keep], [costs[k] for k in keep], [opt[k] for k in keep]
        self.best = 0 if self.best is not None else None
        if self.hcp.get("delete_detours_backwards", True):
            self._delete_detours()

    def _start_orientation(self, rec, length):
        for q in rec:
            v = rec[0, :2] - q[:2]
            if math.hypot(v[0], v[1]) > length:
                return math.atan2(v[1], v[0])
        return None

    def _delete_detours(self):
        if len(self.tebs) < 2 or self.best is None or len(self.tebs[self.best]) < 2:
            return
        length = self.hcp.get("length_start_orientation_vector", 0.4)
        thr = self.hcp.get("detours_orientation_tolerance", math.pi / 2)
        ratio = self.hcp.get("max_ratio_detours_duration_best_duration", 3.0)
        cur = self._start_orientation(self.tebs[self.best], length)
        if cur is None:
            return
        best_dur = max(self.tebs[self.best][:, 3].sum(), 1.0)
        keep = []
        for k, rec in enumerate(self.tebs):
            if k == self.best:
                keep.append(k)
                continue
            o = self._start_orientation(rec, length) if len(rec) >= 2 else None
            if o is None or abs(_nt(o - cur)) > thr or not self.optimized[k] or rec[:, 3].sum() / best_dur > ratio:
                continue
            keep.append(k)
        # parallel containers
        self.classes = [self.classes[k] for k in keep] if len(self.classes) == len(self.tebs) else self.classes
        self.best = keep.index(self.best)
        self.tebs, self.costs, self.optimized = [self.tebs[k] for k in keep], [self.costs[k] for k in keep], [self.optimized[k] for k in keep]

    def add_and_init(self, path, start_orient, goal_orient):
        n0 = len(self.tebs)
        super().add_and_init(path, start_orient, goal_orient)
        if len(self.tebs) > n0:
            self.costs.append(float("inf"))
            self.optimized.append(False)

    # timed_elastic_band.cpp:555-597
    @staticmethod
    def _prune(rec, new_start, new_goal, min_samples=3):
        rec = rec.copy()
        n = len(rec)
        d_cache = math.hypot(new_start[0] - rec[0, 0], new_start[1] - rec[0, 1])
        nearest = 0
        for i in range(1, min(n - min_samples, 10) + 1):
            d = math.hypot(new_start[0] - rec[i, 0], new_start[1] - rec[i, 1])
            if d < d_cache:
                d_cache, nearest = d, i
            else:
                break
        if nearest > 0:
            rec = np.concatenate([rec[:1], rec[1 + nearest:]])
        rec[0, :3] = new_start
        rec[-1, :3] = new_goal
        return rec

    def plan(self, start, goal, args, abi_module):
        """one HomotopyClassPlanner::plan(start, goal) cycle; returns the index of the best band"""
        start, goal = np.asarray(start, float), np.asarray(goal, float)
        # updateAllTEBs :443-463
        if self.tebs:
            back = self.tebs[0][-1]
            if (math.hypot(goal[0] - back[0], goal[1] - back[1]) >= self.p.force_reinit_new_goal_dist or
                    abs(_nt(goal[2] - back[2])) >= self.p.force_reinit_new_goal_angular):
                self.tebs, self.costs, self.optimized, self.classes, self.best = [], [], [], [], None
        self.tebs = [self._prune(r, start, goal, self.p.min_samples) for r in self.tebs]
        # explore :337-357
        self.renew_and_analyze()
        if self.simple:
            self.lr_key_point_graph(start, goal, self.p.min_obstacle_dist)
        else:
            self.prob_roadmap_graph(start, goal, self.p.min_obstacle_dist)
        # optimizeAllTEBs :466-493
        for k, rec in enumerate(self.tebs):
            out, cost, st = self.ob.optimize_band(self.p, rec, len(rec), self.rows, args=args,
                                                 jac_mode=self.ob.JAC_ANALYTIC if self.jac_mode is None else self.jac_mode, n_cap=512,
                                                 obst_vertices=self.obst_vertices)
            self.tebs[k], self.costs[k] = out, cost
            self.optimized[k] = bool(st.status & abi_module.TEB_STATUS_OPTIMIZED)
        # selectBestTeb :564-616 (no initial-plan band in this entry point, no switching blocking period)
        best, min_cost = None, float("inf")
        for k, c in enumerate(self.costs):
            cc = c * self.p.selection_cost_hysteresis if (self.best is not None and k == self.best) else c
            if cc < min_cost:
                best, min_cost = k, cc
        self.best = best
        return best
