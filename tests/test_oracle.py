"""CPU tests of the oracle (the checker): reference known answers, self-consistency, second opinion."""
import math
import os

import numpy as np
import pytest

from teb_local_planner_b200 import abi, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _straight_band(dts):
    """test/teb_basics.cpp: poses (i,0,0), i = 0..len(dts)."""
    n = len(dts) + 1
    rec = np.zeros((n, 4))
    rec[:, 0] = np.arange(n)
    rec[:-1, 3] = dts
    return rec


# --- the reference's own unit tests (test/teb_basics.cpp:5-68), the only reference-owned known answers ---
@pytest.mark.parametrize("case", ["large_at_end", "small_at_end", "middle_and_end"])
def test_autoresize_reference_gtests(oracle, case):
    dt, hyst = 0.1, 0.1 / 3.0
    dts = [dt] * 9
    if case == "large_at_end":       # TEBBasic.autoResizeLargeValueAtEnd
        dts.append(dt + 2 * hyst)
    elif case == "small_at_end":     # TEBBasic.autoResizeSmallValueAtEnd
        dts.append(dt - 2 * hyst)
    else:                            # TEBBasic.autoResize
        dts[5] = dt + 2 * hyst
        dts.append(dt - 2 * hyst)
    out = oracle.auto_resize(_straight_band(dts), len(dts) + 1, dt, hyst, 3, 100, False)
    d = out[:-1, 3]
    assert np.all(d <= dt + hyst + 1e-3)
    assert np.all(dt - hyst - 1e-3 <= d)
    assert out[0, 0] == 0.0 and out[-1, 0] == 10.0      # start / goal untouched


def test_autoresize_split_inserts_average_pose(oracle):
    rec = np.array([[0, 0, 0.2, 1.0], [2, 2, 0.6, 0]], float)
    out = oracle.auto_resize(rec, 2, 0.3, 0.1, 3, 500, False, n_cap=64)
    assert len(out) >= 4
    assert np.all(out[:-1, 3] <= 0.3 + 0.1 + 1e-12)
    assert abs(out[:-1, 3].sum() - 1.0) < 1e-12
    # inserted poses lie on the segment and headings are between the two (PoseSE2::average)
    assert np.allclose(out[:, 0], out[:, 1])
    assert np.all((out[:, 2] >= 0.2 - 1e-12) & (out[:, 2] <= 0.6 + 1e-12))


def test_helpers_known_answers(oracle):
    L = oracle.lib()
    assert L.teb_oracle_normalize_theta(0.5) == 0.5
    assert abs(L.teb_oracle_normalize_theta(math.pi) + math.pi) < 1e-15     # pi -> -pi (half-open interval)
    assert abs(L.teb_oracle_normalize_theta(3 * math.pi + 0.1) - (-math.pi + 0.1)) < 1e-12
    assert abs(L.teb_oracle_normalize_theta(-3 * math.pi - 0.1) - (math.pi - 0.1)) < 1e-12
    assert abs(L.teb_oracle_average_angle(0.2, 0.6) - 0.4) < 1e-15
    # penalties.h:57-117
    assert L.teb_oracle_penalty_interval(0.3, 0.4, 0.05) == 0.0
    assert abs(L.teb_oracle_penalty_interval(0.5, 0.4, 0.05) - 0.15) < 1e-15
    assert abs(L.teb_oracle_penalty_interval(-0.5, 0.4, 0.05) - 0.15) < 1e-15
    assert abs(L.teb_oracle_penalty_interval2(-0.3, -0.2, 0.4, 0.05) - 0.15) < 1e-15
    assert L.teb_oracle_penalty_interval2(0.0, -0.2, 0.4, 0.05) == 0.0
    assert abs(L.teb_oracle_penalty_below(0.4, 0.5, 0.05) - 0.15) < 1e-15
    assert L.teb_oracle_penalty_below(0.56, 0.5, 0.05) == 0.0


def test_init_trajectory_plumbing_scene(oracle):
    """plan(start, goal) cold start: diststep = 0 -> start, forced mid samples up to min_samples, goal
    (timed_elastic_band.cpp:325-387, test_optim_node.cpp:168)."""
    rec = oracle.init_trajectory([-4, 0, 0], [4, 0, 0], 0.0, 0.4, 3)
    assert len(rec) == 3
    assert np.allclose(rec[:, 0], [-4, 0, 4]) and np.allclose(rec[:, 1:3], 0)
    assert np.allclose(rec[:2, 3], [4 / 0.4, 4 / 0.4])


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_analytic_vs_numeric_system(oracle, cfg):
    """closed-form Jacobians agree with g2o-style central differences (delta 1e-9) at 1e-6 relative"""
    p, hb = scenes.make_config_batch(cfg, candidates=3, seed=3)
    for b in range(hb.B):
        via = hb.via[b] if hb.V_cap else None
        obst = hb.obstacles[0][:hb.obst_count[0]]
        for mult in (1.0, 4.0):
            Hn, bn, c2n = oracle.build_system(p, hb.poses[b], hb.n[b], obst, via=via, weight_multiplier=mult, jac_mode=0)
            Ha, ba, c2a = oracle.build_system(p, hb.poses[b], hb.n[b], obst, via=via, weight_multiplier=mult, jac_mode=1)
            assert c2n == c2a
            assert np.abs(Hn - Ha).max() <= 1e-6 * np.abs(Ha).max()
            assert np.abs(bn - ba).max() <= 1e-6 * max(np.abs(ba).max(), 1.0)
            # structure: symmetric, half bandwidth <= 10, SPD after damping
            assert np.allclose(Ha, Ha.T)
            i, j = np.nonzero(Ha)
            assert np.abs(i - j).max() <= 10
            np.linalg.cholesky(Ha + 1e-6 * np.abs(np.diag(Ha)).max() * np.eye(len(Ha)))


def test_two_circles_footprint_and_exponent_jacobians(oracle):
    p, hb = scenes.make_config_batch("C3", candidates=2, seed=5)
    p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
    p.footprint_front_offset, p.footprint_front_radius = 0.3, 0.2
    p.footprint_rear_offset, p.footprint_rear_radius = 0.2, 0.25
    p.obstacle_cost_exponent = 2.0
    p.exact_arc_length = 1
    p.weight_shortest_path = 0.5
    obst = hb.obstacles[0][:hb.obst_count[0]]
    Hn, bn, _ = oracle.build_system(p, hb.poses[0], hb.n[0], obst, rotdir=abi.TEB_ROTDIR_LEFT, jac_mode=0)
    Ha, ba, _ = oracle.build_system(p, hb.poses[0], hb.n[0], obst, rotdir=abi.TEB_ROTDIR_LEFT, jac_mode=1)
    assert np.abs(Hn - Ha).max() <= 2e-6 * np.abs(Ha).max()
    assert np.abs(bn - ba).max() <= 2e-6 * np.abs(ba).max()


@pytest.mark.parametrize("acc_lim_y,max_vel_trans", [(0.5, 0.45), (0.0, 0.0)])
def test_holonomic_edges_jacobians_and_known_answers(oracle, acc_lim_y, max_vel_trans):
    """EdgeVelocityHolonomic / EdgeAccelerationHolonomic*: closed-form Jacobians vs central differences, and the
    residual values against a hand evaluation of edge_velocity.h:250-269 on one segment"""
    p, hb = scenes.make_config_batch("C3", candidates=2, seed=8)
    p.max_vel_y, p.acc_lim_y, p.max_vel_trans = 0.3, acc_lim_y, max_vel_trans
    p.weight_kinematics_nh, p.weight_max_vel_y, p.weight_acc_lim_y = 1.0, 2.0, 1.5
    rng = np.random.default_rng(3)
    n = hb.n[0]
    hb.poses[0, 1:n - 1, 2] += rng.normal(0, 0.35, n - 2)
    hb.poses[0, :n - 1, 3] *= 0.8
    obst = hb.obstacles[0][:hb.obst_count[0]]
    vs = [0.25, 0.1, -0.1, 1.0]
    Hn, bn, c2n = oracle.build_system(p, hb.poses[0], n, obst, vel_start=vs, jac_mode=0)
    Ha, ba, c2a = oracle.build_system(p, hb.poses[0], n, obst, vel_start=vs, jac_mode=1)
    assert c2n == c2a
    assert np.abs(Hn - Ha).max() <= 2e-6 * np.abs(Ha).max()
    assert np.abs(bn - ba).max() <= 2e-6 * max(np.abs(ba).max(), 1.0)
    # chi2 of a 3-pose band with only the holonomic velocity edges switched on, by hand
    q = abi.default_params()
    q.max_vel_y, q.max_vel_trans, q.acc_lim_y = 0.3, max_vel_trans, acc_lim_y
    for w in ("weight_acc_lim_x", "weight_acc_lim_theta", "weight_kinematics_nh", "weight_kinematics_forward_drive",
              "weight_optimaltime", "weight_obstacle", "weight_shortest_path", "weight_viapoint"):
        setattr(q, w, 0.0)
    q.include_dynamic_obstacles = 0
    band = np.array([[0, 0, 0.3, 1.0], [0.5, 0.4, 0.6, 1.2], [1.1, 0.5, 0.2, 0.0]])
    _, _, c2 = oracle.build_system(q, band, 3, np.zeros(0, abi.OBST_DTYPE), jac_mode=1)
    want = 0.0
    for i in range(2):
        dx, dy = band[i + 1, :2] - band[i, :2]
        c, s_ = np.cos(band[i, 2]), np.sin(band[i, 2])
        vx, vy = (c * dx + s_ * dy) / band[i, 3], (-s_ * dx + c * dy) / band[i, 3]
        om = (band[i + 1, 2] - band[i, 2]) / band[i, 3]
        mvx, mvxb, mvy = q.max_vel_x, q.max_vel_x_backwards, q.max_vel_y
        # max_vel_trans == 0 (the header default, teb_config.h:280) collapses all three bounds to 0, as in the reference
        mvy = min(mvy, np.sqrt(max(0.0, max_vel_trans ** 2 - vx ** 2)))
        rem = np.sqrt(max(0.0, max_vel_trans ** 2 - vy ** 2))
        mvx, mvxb = min(mvx, rem), min(mvxb, rem)
        e0 = max(0.0, vx - mvx) if vx >= 0 else max(0.0, -vx - mvxb)          # penaltyBoundToInterval(vx, -b, a, 0)
        e1 = max(0.0, abs(vy) - mvy)
        e2 = max(0.0, abs(om) - (q.max_vel_theta - q.penalty_epsilon))
        want += q.weight_max_vel_x * e0 ** 2 + q.weight_max_vel_y * e1 ** 2 + q.weight_max_vel_theta * e2 ** 2
    assert want > 0 and abs(c2 - want) <= 1e-12 * want


def test_banded_vs_dense_solver(oracle):
    p, hb = scenes.make_config_batch("C1", candidates=2, seed=1)
    obst = hb.obstacles[0][:hb.obst_count[0]]
    args = abi.make_args(4, 3, True, 100.0, 1.0, False)
    a, ca, sa = oracle.optimize_band(p, hb.poses[0], hb.n[0], obst, args=args, jac_mode=1, solver=oracle.SOLVER_BANDED)
    d, cd, sd = oracle.optimize_band(p, hb.poses[0], hb.n[0], obst, args=args, jac_mode=1, solver=oracle.SOLVER_DENSE)
    assert sa.lm_trials == sd.lm_trials
    assert np.abs(a - d).max() < 1e-8
    assert abs(ca - cd) <= 1e-8 * abs(cd)


def test_fixed_endpoints_and_descent(oracle):
    p, hb = scenes.make_config_batch("C2", candidates=4, seed=2)
    obst = hb.obstacles[0][:hb.obst_count[0]]
    for b in range(hb.B):
        _, _, chi0 = oracle.build_system(p, hb.poses[b], hb.n[b], obst, jac_mode=1)
        out, cost, st = oracle.optimize_band(p, hb.poses[b], hb.n[b], obst, args=abi.make_args(5, 1, True), jac_mode=0)
        assert np.array_equal(out[0, :3], hb.poses[b, 0, :3]) and np.array_equal(out[-1, :3], hb.poses[b, hb.n[b] - 1, :3])
        assert st.chi2_final <= chi0 + 1e-12       # accepted LM steps never increase chi2
        assert st.status & abi.TEB_STATUS_OPTIMIZED
        assert np.all(np.isfinite(out))


def test_numpy_second_opinion(oracle):
    """independent numpy restatement (dense numeric J, numpy Cholesky) agrees with the C oracle on a small band"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("teb_oracle_np", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "teb_oracle_np.py"))
    npo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(npo)
    p = abi.default_params()
    p.teb_autosize = 0
    hb = scenes.make_batch(14, 4, 2, seed=7, moving=False, via_points=2)
    rng = np.random.default_rng(0)
    for b in range(2):
        obst = hb.obstacles[0][:4].copy()
        obst["x"] = rng.uniform(-0.5, 0.5, 4)      # close to the band so that obstacle edges are active
        obst["y"] = rng.uniform(-0.6, 0.6, 4)
        obst[3]["vx"], obst[3]["vy"], obst[3]["dynamic"] = 0.1, -0.05, 1
        args = abi.make_args(3, 2, True, 100.0, 2.0, False)
        ref, cref, st = oracle.optimize_band(p, hb.poses[b], 14, obst, via=hb.via[b], args=args, jac_mode=0,
                                             solver=oracle.SOLVER_DENSE)
        got, cgot = npo.optimize(p, hb.poses[b, :14], obst, via=hb.via[b], inner=3, outer=2, obst_scale=100.0, via_scale=2.0)
        assert np.abs(got - ref).max() < 1e-5
        assert abs(cgot - cref) <= 1e-5 * max(abs(cref), 1.0)


def _load_np_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("teb_oracle_np", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "teb_oracle_np.py"))
    npo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(npo)
    return npo


@pytest.mark.parametrize("acc_lim_y", [0.5, 0.0])
def test_numpy_second_opinion_holonomic(oracle, acc_lim_y):
    """holonomic velocity / acceleration edges: the numpy restatement (stacked residuals, dense numeric Jacobian) and the C
    oracle reach the same band"""
    npo = _load_np_oracle()
    p = abi.default_params()
    p.teb_autosize = 0
    p.max_vel_y, p.acc_lim_y, p.max_vel_trans, p.weight_kinematics_nh = 0.3, acc_lim_y, 0.45, 1.0
    hb = scenes.make_batch(12, 3, 2, seed=11, moving=False)
    rng = np.random.default_rng(2)
    for b in range(2):
        rec = hb.poses[b, :12].copy()
        rec[1:-1, 2] += rng.normal(0, 0.3, 10)
        rec[:-1, 3] *= 0.8
        obst = hb.obstacles[0][:3].copy()
        obst["x"], obst["y"] = rng.uniform(-0.5, 0.5, 3), rng.uniform(-0.6, 0.6, 3)
        vs = [0.2, 0.05, -0.1, 1.0]
        args = abi.make_args(3, 2, True, 100.0, 1.0, False)
        ref, cref, _ = oracle.optimize_band(p, rec, 12, obst, vel_start=vs, args=args, jac_mode=0, solver=oracle.SOLVER_DENSE)
        got, cgot = npo.optimize(p, rec, obst, inner=3, outer=2, obst_scale=100.0, via_scale=1.0, vel_start=vs)
        # both sides use central differences with delta = 1e-9: their ~1e-7 relative Jacobian noise is amplified by the
        # stiff holonomic problem; the north-star tolerance (1e-4 per pose component) is the bar
        assert np.abs(got - ref).max() < 1e-4
        assert abs(cgot - cref) <= 1e-3 * max(abs(cref), 1.0)   # cost gradient ~1e2 x pose difference ~1e-5


def test_numpy_second_opinion_h_signatures(oracle):
    """both signature kinds against an independent numpy evaluation (numpy longdouble complex arithmetic / numpy cross
    products)"""
    npo = _load_np_oracle()
    p, hb = scenes.make_config_batch("C4", candidates=3, seed=5)
    obst = hb.obstacles[0][:12]
    for b in range(hb.B):
        rec = hb.poses[b, :40].copy()
        p.include_dynamic_obstacles = 0
        h2 = oracle.h_signature(p, rec, 40, obst)
        assert abs(h2 - npo.h_signature_2d(p, rec, obst)) <= 1e-13 * abs(h2)
        p.include_dynamic_obstacles = 1
        for use_dt in (True, False):
            h3 = oracle.h_signature(p, rec, 40, obst, use_timediffs=use_dt)
            assert np.abs(h3 - npo.h_signature_3d(p, rec, obst, use_dt)).max() < 1e-13


def test_golden_fixtures_pin_oracle(oracle):
    """the committed golden vectors (tests/golden/make_golden.py) guard the oracle against drift"""
    path = os.path.join(GOLDEN, "golden_v1.npz")
    g = np.load(path, allow_pickle=False)
    from tests.golden import make_golden
    for name, hb_in, p, args in make_golden.cases():
        hb = hb_in.copy()
        oracle.optimize_batch(p, hb, args, jac_mode=1)
        assert np.array_equal(hb.n, g[f"{name}_n"])
        for b in range(hb.B):
            assert np.abs(hb.poses[b, :hb.n[b]] - g[f"{name}_poses"][b, :hb.n[b]]).max() < 1e-9
        assert np.allclose(hb.cost, g[f"{name}_cost"], rtol=1e-9)


def test_golden_v2_pins_oracle(oracle):
    """second fixture file: holonomic robot, vertex-list shapes with a polygon footprint, both H-signature kinds"""
    g = np.load(os.path.join(GOLDEN, "golden_v2.npz"), allow_pickle=False)
    from tests.golden import make_golden
    for name, hb_in, p, args, _ in make_golden.cases_v2():
        hb = hb_in.copy()
        oracle.optimize_batch(p, hb, args, jac_mode=1)
        assert np.array_equal(hb.n, g[f"{name}_n"])
        for b in range(hb.B):
            assert np.abs(hb.poses[b, :hb.n[b]] - g[f"{name}_poses"][b, :hb.n[b]]).max() < 1e-9
        assert np.allclose(hb.cost, g[f"{name}_cost"], rtol=1e-9)
    for name, p, hb in make_golden.hsig_cases():
        obst = hb.obstacles[0][:hb.obst_count[0]]
        for b in range(hb.B):
            v = oracle.h_signature(p, hb.poses[b], hb.n[b], obst)
            want = g[name][b]
            if p.include_dynamic_obstacles:
                assert np.array_equal(v, want[:len(obst)])
            else:
                assert v == complex(want[0], want[1])


# ---------------------------------------------------------------------------------------------------------------------
# Line / Pill / Polygon obstacles and Line / Polygon footprints (distance_calculations.h, obstacles.h:597-1045,
# robot_footprint_model.h:439-760)
def _ob(type_, x=0.0, y=0.0, radius=0.0, vx=0.0, vy=0.0, begin=0, count=0, dynamic=0):
    o = np.zeros(1, abi.OBST_DTYPE)
    o["x"], o["y"], o["vx"], o["vy"], o["radius"] = x, y, vx, vy, radius
    o["type"], o["dynamic"], o["vertex_begin"], o["vertex_count"] = type_, dynamic, begin, count
    return o


def _footprint(kind, line=None, poly=None):
    p = abi.default_params()
    p.footprint_type = kind
    if line is not None:
        for k, v in enumerate(line):
            p.footprint_line[k] = v
    if poly is not None:
        p.footprint_vertex_count = len(poly)
        for k, (x, y) in enumerate(poly):
            p.footprint_vertices[2 * k], p.footprint_vertices[2 * k + 1] = x, y
    return p


def test_shape_distances_known_answers(oracle):
    pt = _footprint(abi.TEB_FOOTPRINT_POINT)
    line_v = [[1, -1], [1, 1]]
    d, g = oracle.distance(pt, [0, 0, 0.3], _ob(abi.TEB_OBST_LINE, 1, 0, count=2), line_v, want_grad=True)
    assert d == 1.0 and np.allclose(g, [-1, 0, 0])
    assert oracle.distance(pt, [0, 0, 0], _ob(abi.TEB_OBST_PILL, 1, 0, radius=0.2, count=2), line_v) == 0.8
    assert abs(oracle.distance(pt, [0, 3, 0], _ob(abi.TEB_OBST_LINE, 1, 0, count=2), line_v) - np.hypot(1, 2)) < 1e-15
    sq = [[1, -1], [3, -1], [3, 1], [1, 1]]
    assert oracle.distance(pt, [0, 0, 0], _ob(abi.TEB_OBST_POLYGON, 2, 0, count=4), sq) == 1.0
    # inside a polygon the reference still returns the distance to the nearest edge (no interior test,
    # distance_calculations.h:172-199)
    assert oracle.distance(pt, [2, 0.5, 0], _ob(abi.TEB_OBST_POLYGON, 2, 0, count=4), sq) == 0.5
    # vertex pool offset, 1-vertex and 2-vertex polygons (point / line case)
    pool = [[9, 9], [9, 9], [1, -1], [1, 1], [4, 0]]
    assert oracle.distance(pt, [0, 0, 0], _ob(abi.TEB_OBST_POLYGON, 1, 0, begin=2, count=2), pool) == 1.0
    assert oracle.distance(pt, [0, 0, 0], _ob(abi.TEB_OBST_POLYGON, 4, 0, begin=4, count=1), pool) == 4.0
    # constant-velocity prediction shifts every vertex (obstacles.h:676-696, predictVertices :994)
    assert oracle.distance(pt, [0, 0, 0], _ob(abi.TEB_OBST_LINE, 1, 0, vx=1.0, count=2, dynamic=1), line_v, t=2.0) == 3.0
    # circular footprint subtracts its radius
    cf = _footprint(abi.TEB_FOOTPRINT_CIRCULAR)
    cf.footprint_radius = 0.25
    assert oracle.distance(cf, [0, 0, 0], _ob(abi.TEB_OBST_PILL, 1, 0, radius=0.25, count=2), line_v) == 0.5
    # line footprint (-0.5,0)-(0.5,0) turned by 90 degrees
    lf = _footprint(abi.TEB_FOOTPRINT_LINE, line=[-0.5, 0, 0.5, 0])
    pose = [0, 0, np.pi / 2]
    assert abs(oracle.distance(lf, pose, _ob(abi.TEB_OBST_POINT, 1, 0.2)) - 1.0) < 1e-15
    assert abs(oracle.distance(lf, pose, _ob(abi.TEB_OBST_CIRCULAR, 1, 0.2, radius=0.3)) - 0.7) < 1e-15
    assert abs(oracle.distance(lf, pose, _ob(abi.TEB_OBST_POINT, 0, 2.0)) - 1.5) < 1e-15          # beyond the end point
    d, g = oracle.distance(lf, pose, _ob(abi.TEB_OBST_LINE, 0, 0, count=2), [[-1, 0.1], [1, 0.1]], want_grad=True)
    assert d == 0.0 and np.all(g == 0)                                                              # crossing segments
    assert abs(oracle.distance(lf, pose, _ob(abi.TEB_OBST_LINE, 0, 0, count=2), [[1, -3], [1, 3]]) - 1.0) < 1e-15
    # polygon footprint: unit square turned by 45 degrees, corner towards the obstacle
    pf = _footprint(abi.TEB_FOOTPRINT_POLYGON, poly=[(-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5)])
    d, g = oracle.distance(pf, [0, 0, np.pi / 4], _ob(abi.TEB_OBST_POINT, 2, 0), want_grad=True)
    assert abs(d - (2 - np.sqrt(0.5))) < 1e-15 and np.allclose(g, [-1, 0, 0], atol=1e-15)
    assert abs(oracle.distance(pf, [0, 0, 0], _ob(abi.TEB_OBST_POLYGON, 2, 0, count=4), sq) - 0.5) < 1e-15
    assert oracle.distance(pf, [0.8, 0, 0], _ob(abi.TEB_OBST_POLYGON, 2, 0, count=4), sq) == 0.0    # overlapping polygons
    # two-circles footprint against a line obstacle: min over the two circles
    tc = _footprint(abi.TEB_FOOTPRINT_TWO_CIRCLES)
    tc.footprint_front_offset, tc.footprint_front_radius, tc.footprint_rear_offset, tc.footprint_rear_radius = 0.4, 0.1, 0.3, 0.2
    assert abs(oracle.distance(tc, [0, 0, 0], _ob(abi.TEB_OBST_LINE, 1, 0, count=2), line_v) - 0.5) < 1e-15


def _seg_seg_ericson(p1, q1, p2, q2):
    """independent closest distance of two segments (clamped quadratic minimisation), not the reference's scheme"""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-300 and e <= 1e-300:
        return np.linalg.norm(r)
    if a <= 1e-300:
        s, t = 0.0, np.clip(f / e, 0, 1)
    else:
        c = d1 @ r
        if e <= 1e-300:
            t, s = 0.0, np.clip(-c / a, 0, 1)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = np.clip((b * f - c * e) / den, 0, 1) if den > 1e-300 else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, np.clip(-c / a, 0, 1)
            elif t > 1:
                t, s = 1.0, np.clip((b - c) / a, 0, 1)
    return np.linalg.norm((p1 + d1 * s) - (p2 + d2 * t))


def _edges(v):
    v = np.asarray(v, float)
    if len(v) == 1:
        return [(v[0], v[0])]
    if len(v) == 2:
        return [(v[0], v[1])]
    return [(v[i], v[(i + 1) % len(v)]) for i in range(len(v))]


@pytest.mark.parametrize("seed", range(6))
def test_shape_distances_second_opinion_and_gradients(oracle, seed):
    """every footprint x obstacle pair: value against an independent segment-segment routine, closed-form gradient
    against central differences (skipping the kinks, where forward and backward differences disagree)"""
    rng = np.random.default_rng(seed)
    fps = [(_footprint(abi.TEB_FOOTPRINT_POINT), [(0, 0)], 0.0),
           (_footprint(abi.TEB_FOOTPRINT_LINE, line=[-0.4, 0.1, 0.6, -0.05]), [(-0.4, 0.1), (0.6, -0.05)], 0.0),
           (_footprint(abi.TEB_FOOTPRINT_POLYGON, poly=[(-0.3, -0.25), (0.5, -0.2), (0.45, 0.3), (-0.35, 0.2)]),
            [(-0.3, -0.25), (0.5, -0.2), (0.45, 0.3), (-0.35, 0.2)], 0.0)]
    c = _footprint(abi.TEB_FOOTPRINT_CIRCULAR)
    c.footprint_radius = 0.2
    fps.append((c, [(0, 0)], 0.2))
    checked = 0
    for p, local, rrad in fps:
        for otype in (abi.TEB_OBST_POINT, abi.TEB_OBST_CIRCULAR, abi.TEB_OBST_LINE, abi.TEB_OBST_PILL, abi.TEB_OBST_POLYGON):
            for _ in range(12):
                pose = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-np.pi, np.pi)])
                ctr = rng.uniform(-2.5, 2.5, 2)
                nv = {abi.TEB_OBST_LINE: 2, abi.TEB_OBST_PILL: 2, abi.TEB_OBST_POLYGON: int(rng.integers(3, 7))}.get(otype, 0)
                ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
                verts = ctr + np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(0.3, 1.0, (nv, 1)) if nv else np.zeros((0, 2))
                orad = rng.uniform(0.05, 0.3) if otype in (abi.TEB_OBST_CIRCULAR, abi.TEB_OBST_PILL) else 0.0
                vel = rng.uniform(-0.3, 0.3, 2)
                t = float(rng.uniform(0, 3)) if rng.random() < 0.5 else 0.0
                ob = _ob(otype, ctr[0], ctr[1], radius=orad, vx=vel[0], vy=vel[1], count=nv)
                d, g = oracle.distance(p, pose, ob, verts, t=t, want_grad=True)
                # second opinion
                cs, sn = np.cos(pose[2]), np.sin(pose[2])
                R = np.array([[cs, -sn], [sn, cs]])
                rw = [pose[:2] + R @ np.array(v, float) for v in local]
                ow = (verts if nv else ctr[None]) + t * vel
                best = min(_seg_seg_ericson(a0, a1, b0, b1) for a0, a1 in _edges(rw) for b0, b1 in _edges(ow))
                # the reference returns 0 for crossing segments and the Ericson routine as well
                assert abs(d - (best - rrad - orad)) < 1e-12, (p.footprint_type, otype)
                # gradient
                h = 1e-6
                fd = np.zeros(3)
                smooth = True
                for k in range(3):
                    e = np.zeros(3)
                    e[k] = h
                    dp, dm = oracle.distance(p, pose + e, ob, verts, t=t), oracle.distance(p, pose - e, ob, verts, t=t)
                    fd[k] = (dp - dm) / (2 * h)
                    if abs((dp - d) - (d - dm)) > 1e-9:
                        smooth = False
                if smooth:
                    assert np.abs(fd - g).max() < 1e-6, (p.footprint_type, otype, fd, g)
                    checked += 1
    assert checked > 150


@pytest.mark.parametrize("footprint", ["point", "line", "polygon", "two_circles"])
@pytest.mark.parametrize("legacy", [0, 1])
def test_shape_obstacles_system_jacobians(oracle, footprint, legacy):
    """buildGraph with Line / Pill / Polygon obstacles (static and moving) and Line / Polygon footprints: closed-form
    system vs central differences; the association sees the shapes through calculateDistance and getCentroid"""
    p, hb = scenes.make_config_batch("C4", candidates=2, seed=6)
    hb = scenes.add_shape_obstacles(hb, seed=1)
    p.legacy_obstacle_association, p.obstacle_poses_affected = legacy, 8
    if footprint == "line":
        scenes.set_line_footprint(p)
    elif footprint == "polygon":
        scenes.set_polygon_footprint(p)
    elif footprint == "two_circles":
        p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
        p.footprint_front_offset, p.footprint_front_radius, p.footprint_rear_offset, p.footprint_rear_radius = 0.3, 0.15, 0.2, 0.2
    obst = hb.obstacles[0][:hb.obst_count[0]].copy()
    obst["dynamic"][::3] = 0           # a mix of static and moving shapes
    for b in range(hb.B):
        Hn, bn, c2n = oracle.build_system(p, hb.poses[b], hb.n[b], obst, via=hb.via[b], jac_mode=0, obst_vertices=hb.obst_vertices[0])
        Ha, ba, c2a = oracle.build_system(p, hb.poses[b], hb.n[b], obst, via=hb.via[b], jac_mode=1, obst_vertices=hb.obst_vertices[0])
        assert c2n == c2a and c2a > 0
        assert np.abs(Hn - Ha).max() <= 5e-6 * np.abs(Ha).max()
        assert np.abs(bn - ba).max() <= 5e-6 * max(np.abs(ba).max(), 1.0)
    # the shapes matter: the same scene with the vertex obstacles collapsed to points gives another system
    flat = obst.copy()
    flat["type"], flat["radius"] = abi.TEB_OBST_POINT, 0.0
    Hp, _, c2p = oracle.build_system(p, hb.poses[b], hb.n[b], flat, via=hb.via[b], jac_mode=1)
    assert abs(c2p - c2a) > 1e-6 or np.abs(Hp - Ha).max() > 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# H-signatures (h_signature.h): which candidates are the same homotopy class
def _path(points):
    rec = np.zeros((len(points), 4))
    rec[:, :2] = points
    rec[:-1, 3] = np.hypot(*np.diff(np.asarray(points, float), axis=0).T) / 0.4
    return rec


def test_h_signature_2d_known_answer(oracle):
    """one obstacle: H = A (ln|z_end - o| - ln|z_start - o| + i * swept angle), A = prescaler * a (o - BL) * b (o - TR) with
    a = 3, b = 2 (h_signature.h:119-123, :160); passing above sweeps -pi, below +pi"""
    p = abi.default_params()
    p.include_dynamic_obstacles = 0
    ob = _ob(abi.TEB_OBST_POINT, 0.5, 0.2)
    s = np.linspace(0, 1, 41)
    above = _path(np.stack([-4 + 8 * s, 1.5 * np.sin(np.pi * s)], 1))
    below = _path(np.stack([-4 + 8 * s, -1.5 * np.sin(np.pi * s)], 1))
    o = complex(0.5, 0.2)
    start, end = complex(-4, 0), complex(4, 0)
    normal = complex(0, 8)
    A = 1.0 * 3 * (o - (start - normal)) * 2 * (o - (start + (end - start) + normal))
    lr = np.log(abs(end - o)) - np.log(abs(start - o))
    for path, swept in ((above, None), (below, None)):
        ang = np.unwrap(np.angle((path[:, 0] + 1j * path[:, 1]) - o))
        want = A * complex(lr, ang[-1] - ang[0])
        got = oracle.h_signature(p, path, len(path), ob)
        assert abs(got - want) <= 1e-12 * abs(want)
    ha, hb = oracle.h_signature(p, above, len(above), ob), oracle.h_signature(p, below, len(below), ob)
    assert abs(ha - hb) > p.h_signature_threshold                   # different classes
    wiggle = above.copy()
    wiggle[1:-1, 1] += 0.2
    assert abs(oracle.h_signature(p, wiggle, len(wiggle), ob) - ha) < 1e-9 * abs(ha)   # same class: same signature
    assert oracle.h_signature(p, above, len(above), np.zeros(0, abi.OBST_DTYPE)) == 0


def test_h_signature_3d_properties(oracle):
    """x-y-t signature (h_signature.h:282-353): sign = side of the obstacle (positive: obstacle on the left), |H| < 1 without
    a loop, far obstacles ~ 0; the sign test of isEqual (:366-388) separates above / below"""
    p = abi.default_params()
    p.include_dynamic_obstacles = 1
    obs = np.concatenate([_ob(abi.TEB_OBST_POINT, 0.0, 0.0), _ob(abi.TEB_OBST_POINT, 1.0, 40.0),
                          _ob(abi.TEB_OBST_POINT, -1.0, 0.5, vx=0.05, vy=-0.02, dynamic=1)])
    s = np.linspace(0, 1, 41)
    above = _path(np.stack([-4 + 8 * s, 1.5 * np.sin(np.pi * s)], 1))
    below = _path(np.stack([-4 + 8 * s, -1.5 * np.sin(np.pi * s)], 1))
    ha, hb = oracle.h_signature(p, above, len(above), obs), oracle.h_signature(p, below, len(below), obs)
    assert ha[0] < -p.h_signature_threshold and hb[0] > p.h_signature_threshold     # passing above: obstacle on the right
    assert np.all(np.abs(ha) < 1) and np.all(np.abs(hb) < 1)
    assert abs(ha[1]) < 0.02 and abs(hb[1]) < 0.02                                  # obstacle 40 m away
    assert abs(ha[0] + hb[0]) < 1e-12                                               # mirror paths, static obstacle on the axis
    # without time information the transition times are |dz| / max_vel_x (:309-310) = the dt of this synthetic path
    hn = oracle.h_signature(p, above, len(above), obs, use_timediffs=False)
    assert np.allclose(hn, ha, rtol=0, atol=1e-12)


def test_exploration_restatement_pieces():
    """oracle/hcp_explore.py: mt19937 known answers (first draw, the standard's 10000th value), the path initialisation
    against a hand evaluation (timed_elastic_band.hpp:46-185), segment intersection corner cases"""
    from oracle import hcp_explore as X
    m = X.MT19937()
    xs = [m() for _ in range(10000)]
    assert xs[0] == 3499211612 and xs[-1] == 4123659995
    rec = X.init_from_path([(0, 0), (1, 0), (1, 2)], max_vel_x=0.4, acc_lim_x=0.5, start_orient=0.3, goal_orient=-0.2,
                           min_samples=3, guess_backwards=False)
    assert rec.shape == (3, 4)
    assert np.allclose(rec[:, :3], [[0, 0, 0.3], [1, 0, 0.0], [1, 2, -0.2]])
    assert np.allclose(rec[:2, 3], [max(1 / 0.4, np.sqrt(2 * 1 / 0.5)), max(2 / 0.4, np.sqrt(2 * 2 / 0.5))])
    short = X.init_from_path([(0, 0), (1, 0)], 0.4, 0.5, 0.0, 0.0, min_samples=4, guess_backwards=False)
    assert len(short) == 4 and np.allclose(short[:, 0], [0, 0.5, 0.75, 1.0]) and np.allclose(short[:3, 3], [1.25, 0.625, 0.625])
    a = np.array
    assert X._segments_intersect(a([0., 0]), a([1., 1]), a([0., 1]), a([1., 0]))
    assert not X._segments_intersect(a([0., 0]), a([1., 0]), a([0., 1]), a([1., 1]))       # parallel
    assert not X._segments_intersect(a([0., 0]), a([1., 0]), a([2., -1]), a([2., 1]))      # beyond the end


def test_planner_cycles_restatement_invariants(oracle):
    """oracle/hcp_explore.py Planner: three HomotopyClassPlanner::plan cycles with a moving start pose keep at most
    max_number_classes bands in pairwise different classes, keep the best band first in the class list, and the best
    band only changes when another band beats its cost times the hysteresis factor"""
    from oracle import hcp_explore as X
    p = abi.default_params()
    p.include_dynamic_obstacles = 0
    rows = np.zeros(3, abi.OBST_DTYPE)
    rows["x"], rows["y"], rows["radius"] = [-1.5, 0.5, 2.0], [0.3, -0.4, 0.5], [0, 0.3, 0]
    rows["type"] = [abi.TEB_OBST_POINT, abi.TEB_OBST_CIRCULAR, abi.TEB_OBST_POINT]
    obstacles = [X.Obst("point", (-1.5, 0.3)), X.Obst("circle", (0.5, -0.4), 0.3), X.Obst("point", (2.0, 0.5))]
    pl = X.Planner(p, {"max_number_classes": 4, "obstacle_heading_threshold": 0.45}, oracle, rows, obstacles)
    args = abi.make_args(5, 4, True, p.selection_obst_cost_scale, p.selection_viapoint_cost_scale, False)
    prev_best_cost = None
    for cycle, sx in enumerate((-4.0, -3.9, -3.8)):
        best = pl.plan([sx, 0.0, 0.1], [4, 0.2, -0.2], args, abi)
        assert best is not None and 2 <= len(pl.tebs) <= 4
        sigs = [pl._signature(r) for r in pl.tebs]
        if cycle == 0:      # freshly explored bands are pairwise in different classes (the optimiser may merge them later)
            init_classes = list(pl.classes)
            for i in range(len(init_classes)):
                for j in range(i + 1, len(init_classes)):
                    assert not pl._is_equal(init_classes[i], init_classes[j])
        assert all(np.isfinite(c) for c in pl.costs)
        assert all(np.array_equal(r[0, :3], [sx, 0.0, 0.1]) and np.array_equal(r[-1, :3], [4, 0.2, -0.2]) for r in pl.tebs)
        assert pl.costs[best] * (p.selection_cost_hysteresis if prev_best_cost is not None else 1.0) <= min(pl.costs) + 1e-12 or \
            pl.costs[best] == min(pl.costs)
        prev_best_cost = pl.costs[best]
        assert len(sigs) == len(pl.tebs)


# --------------------------------------------------------------------------- reference-generated golden vectors
def test_oracle_matches_reference_generated_golden(oracle):
    """tests/golden/golden_ref_v1.npz was generated by the REFERENCE'S OWN code (oracle/_ref: src/optimal_planner.cpp,
    src/timed_elastic_band.cpp, src/obstacles.cpp and their headers compiled against stand-ins for Eigen / boost / ROS /
    the g2o optimizer; generator: tests/golden/make_golden_ref.py). The restatement must reproduce every stored number
    bit for bit: chi2 and right-hand side of the normal equations at the initial state, final band, n, cost and the
    number of LM trials of whole optimizeTEB calls, for all 24 feature scenarios of tests/scenarios.py."""
    import os
    from tests import scenarios
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref_v1.npz"), allow_pickle=False)
    for name in scenarios.ALL:
        p, hb = scenarios.scenario(name)
        args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                             p.selection_viapoint_cost_scale, False)
        for b in range(hb.B):
            kw = scenarios.band_kwargs(hb, b)
            n = int(hb.n[b])
            H, rhs, c2 = oracle.build_system(p, hb.poses[b], n, jac_mode=oracle.JAC_G2O, **kw)
            assert c2 == z[name + "/chi2_0"][b], name
            assert np.array_equal(rhs, z[name + "/b_0"][b][:len(rhs)]), name
            rec, cost, st = oracle.optimize_band(p, hb.poses[b], n, args=args, jac_mode=oracle.JAC_G2O, n_cap=hb.n_cap, **kw)
            assert len(rec) == z[name + "/n"][b], name
            assert np.array_equal(rec, z[name + "/poses"][b][:len(rec)]), name
            assert cost == z[name + "/cost"][b] and st.lm_trials == z[name + "/trials"][b], name
            assert bool(st.status & abi.TEB_STATUS_TERMINATED) == bool(z[name + "/terminated"][b]), name
