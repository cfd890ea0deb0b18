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
