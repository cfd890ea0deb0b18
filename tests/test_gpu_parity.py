"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C-ABI (ctypes) and is
checked against the CPU oracle on the same seeded inputs, against the committed golden fixtures, and - at the
BASELINE.json sizes - through size-independent properties.

Tolerances (fp64 everywhere):
  * kernel A (H, b, chi2) vs oracle closed-form system: 1e-12 relative (same formulas, different summation order)
  * optimizeTEB vs oracle in closed-form-Jacobian mode: 1e-6 absolute per pose component (observed ~1e-10)
  * optimizeTEB vs oracle in g2o mode (numeric Jacobians, delta = 1e-9, i.e. what the reference runs):
    north-star tolerance 1e-4 per pose component. The numeric Jacobian carries ~1e-7 relative noise, which flips a
    discrete LM accept/reject or hinge decision on a small fraction of bands; the oracle shows the same spread
    between its own two Jacobian modes (worst on the car-like + autoResize config). The test therefore requires, per
    band, that a band outside 1e-4 is one where the oracle's two modes disagree with each other as well, the median
    band within 1e-5, and writes the measured fraction per configuration to gpurun_out/parity_report.json.
"""
import ctypes as C
import os

import numpy as np
import pytest

import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz")


def _gpu(hb, p):
    g = T.TebGpu(hb.B, hb.n_cap, hb.S, max(hb.M_cap, 1), hb.V_cap, max_obst_vertices=hb.PV_cap)
    g.set_params(p)
    return g


def _padded_from_dense(Hd, bd, n):
    N = 4 * n - 7
    Hb = np.zeros((4 * n, 12))
    Hb[:, 0] = 1.0
    for r in range(N):
        for k in range(min(r, 10) + 1):
            Hb[r + 3, k] = Hd[r, r - k]
        Hb[r + 3, 11] = bd[r]
    return Hb


def _report(key, entry):
    """measured parity numbers -> gpurun_out/parity_report.json (merged back by gpurun)"""
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        rep = json.load(open(path)) if os.path.exists(path) else {}
        rep[key] = entry
        json.dump(rep, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _pose_diff(ha, hb_):
    out = []
    for b in range(ha.B):
        if ha.n[b] != hb_.n[b]:
            out.append(np.inf)
        else:
            out.append(np.abs(ha.poses[b, :ha.n[b]] - hb_.poses[b, :ha.n[b]]).max())
    return np.array(out)


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
def test_kernel_a_system_matches_oracle(oracle, cfg):
    p, hb = scenes.make_config_batch(cfg, candidates=5, seed=21)
    g = _gpu(hb, p)
    for outer_index in (0, 3):
        Hb, chi2 = g.build_system(hb, outer_index)
        for b in range(hb.B):
            n = hb.n[b]
            Hd, bd, c2 = oracle.build_system(p, hb.poses[b], n, hb.obstacles[0][:hb.obst_count[0]],
                                             via=hb.via[b] if hb.V_cap else None,
                                             weight_multiplier=2.0 ** outer_index, jac_mode=oracle.JAC_ANALYTIC)
            ref = _padded_from_dense(Hd, bd, n)
            got = Hb[b, :4 * n]
            assert np.abs(got[:, :11] - ref[:, :11]).max() <= 1e-12 * np.abs(ref[:, :11]).max()
            assert np.abs(got[:, 11] - ref[:, 11]).max() <= 1e-12 * max(np.abs(ref[:, 11]).max(), 1.0)
            assert abs(chi2[b] - c2) <= 1e-12 * max(c2, 1.0)
    g.close()


def test_kernel_a_thread_mappings_agree():
    """thread-per-pose (default, k_linearize2) and thread-per-band-row (first generation) kernel A: two independent
    mappings of the same formulas; band and chi2 equal to round-off (different summation order)"""
    p, hb = scenes.make_config_batch("C4", candidates=6, seed=2)
    g = _gpu(hb, p)
    g.set_linearize_variant(1)
    H1, c1 = g.build_system(hb, 2)
    g.set_linearize_variant(0)
    H0, c0 = g.build_system(hb, 2)
    g.close()
    assert np.abs(H1 - H0).max() <= 1e-12 * np.abs(H0).max()
    assert np.allclose(c1, c0, rtol=1e-12)


def test_kernel_a_all_edge_families(oracle):
    """two-circles footprint, cost exponent, exact arc length, shortest path, prefer-rotdir, car-like, free goal vel"""
    p, hb = scenes.make_config_batch("C3", candidates=4, seed=5)
    p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
    p.footprint_front_offset, p.footprint_front_radius = 0.3, 0.2
    p.footprint_rear_offset, p.footprint_rear_radius = 0.2, 0.25
    p.obstacle_cost_exponent = 2.0
    p.exact_arc_length = 1
    p.weight_shortest_path = 0.5
    hb.prefer_rotdir[:] = [abi.TEB_ROTDIR_LEFT, abi.TEB_ROTDIR_RIGHT, 0, abi.TEB_ROTDIR_LEFT]
    hb.vel_start[:, 0], hb.vel_start[:, 2] = 0.3, -0.1
    hb.vel_goal[1, 3] = 0.0        # free_goal_vel for band 1
    hb.vel_goal[2, 0] = 0.2
    g = _gpu(hb, p)
    Hb, chi2 = g.build_system(hb, 1)
    for b in range(hb.B):
        n = hb.n[b]
        Hd, bd, c2 = oracle.build_system(p, hb.poses[b], n, hb.obstacles[0][:hb.obst_count[0]], vel_start=hb.vel_start[b],
                                         vel_goal=hb.vel_goal[b], rotdir=int(hb.prefer_rotdir[b]), weight_multiplier=2.0,
                                         jac_mode=oracle.JAC_ANALYTIC)
        ref = _padded_from_dense(Hd, bd, n)
        got = Hb[b, :4 * n]
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
        assert abs(chi2[b] - c2) <= 1e-12 * max(c2, 1.0)
    g.close()


@pytest.mark.parametrize("cfg,autosize", [("C1", True), ("C2", False), ("C2", True), ("C3", False), ("C3", True),
                                          ("C4", False), ("C4", True)])
def test_optimize_matches_oracle(oracle, cfg, autosize):
    p, hb0 = scenes.make_config_batch(cfg, candidates=16, seed=31, autosize=autosize)
    args = abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                         p.selection_viapoint_cost_scale, False)
    g = _gpu(hb0, p)
    hg = hb0.copy()
    g.optimize(hg, args)
    g.close()
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=8)
    hn = hb0.copy()
    oracle.optimize_batch(p, hn, args, jac_mode=oracle.JAC_G2O, threads=8)
    # closed-form mode: same algorithm, same decisions
    assert np.array_equal(hg.n, ha.n)
    da = _pose_diff(hg, ha)
    assert da.max() < 1e-6, da
    assert np.allclose(hg.cost, ha.cost, rtol=1e-6)
    assert np.array_equal(hg.lm_iters, ha.lm_iters)
    assert np.array_equal(hg.status, ha.status)
    # g2o (numeric-Jacobian) mode: the reference's own path, north-star tolerance 1e-4
    dn = _pose_diff(hg, hn)
    within = dn <= 1e-4
    oracle_spread = _pose_diff(ha, hn)
    # per band, not statistically: a band outside 1e-4 must be one where the CPU implementation's OWN two Jacobian modes
    # part ways (delta = 1e-9 differences flip a discrete accept / reject or hinge decision); the measured fraction goes
    # to gpurun_out/parity_report.json next to the per-scenario numbers of test_gpu_reference.py
    assert np.all(within | (oracle_spread > 1e-5)), (dn, oracle_spread)
    assert np.median(dn) <= 1e-5, dn
    _report(f"{cfg}{'_autosize' if autosize else ''}/16_candidates_vs_g2o_mode_oracle",
            {"bands": int(len(dn)), "fraction_within_1e-4_of_reference": float(within.mean()),
             "median_abs_pose_diff": float(np.median(dn)), "max_abs_pose_diff": float(dn.max()),
             "bands_outside_explained_by_cpu_mode_spread": int((~within).sum())})
    # fixed start / goal, finite outputs
    for b in range(hg.B):
        assert np.array_equal(hg.poses[b, 0, :3], hb0.poses[b, 0, :3])
        assert np.array_equal(hg.poses[b, hg.n[b] - 1, :3], hb0.poses[b, hb0.n[b] - 1, :3])
    assert np.all(np.isfinite(hg.cost)) and np.all(hg.status & abi.TEB_STATUS_OPTIMIZED)


def test_golden_fixtures(oracle):
    from tests.golden import make_golden
    gold = np.load(GOLDEN, allow_pickle=False)
    for name, hb_in, p, args in make_golden.cases():
        hb = hb_in.copy()
        g = _gpu(hb, p)
        g.optimize(hb, args)
        g.close()
        assert np.array_equal(hb.n, gold[f"{name}_n"])
        for b in range(hb.B):
            assert np.abs(hb.poses[b, :hb.n[b]] - gold[f"{name}_poses"][b, :hb.n[b]]).max() < 1e-6
        assert np.allclose(hb.cost, gold[f"{name}_cost"], rtol=1e-6)
        assert np.array_equal(hb.status, gold[f"{name}_status"])


def test_golden_fixtures_v2():
    """golden_v2.npz without the oracle library: holonomic robot (exact), candidate classification (both signature kinds)"""
    from tests.golden import make_golden
    gold = np.load(os.path.join(os.path.dirname(GOLDEN), "golden_v2.npz"), allow_pickle=False)
    for name, hb_in, p, args, gpu_exact in make_golden.cases_v2():
        if not gpu_exact:
            continue
        hb = hb_in.copy()
        g = _gpu(hb, p)
        g.optimize(hb, args)
        g.close()
        assert np.array_equal(hb.n, gold[f"{name}_n"])
        for b in range(hb.B):
            assert np.abs(hb.poses[b, :hb.n[b]] - gold[f"{name}_poses"][b, :hb.n[b]]).max() < 1e-6
        assert np.allclose(hb.cost, gold[f"{name}_cost"], rtol=1e-6)
    for name, p, hb in make_golden.hsig_cases():
        g = _gpu(hb, p)
        got = g.h_signature(hb)
        g.close()
        want = gold[name]
        if p.include_dynamic_obstacles:
            assert np.abs(got[:, :want.shape[1]] - want).max() < 1e-12
        else:
            want = want[:, 0] + 1j * want[:, 1]
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()


def test_ragged_batch_multi_scene_and_edge_cases(oracle):
    """different n per band, several scenes with different obstacle counts (incl. none), minimal n = 3,
    too-few-poses guard, alternative time cost"""
    p = abi.default_params()
    p.teb_autosize = 0
    n_list = [3, 4, 7, 33, 64, 65, 66, 129, 2]
    B, n_cap = len(n_list), 160
    poses = np.zeros((B, n_cap, 4))
    for b, n in enumerate(n_list):
        poses[b, :n] = scenes.make_band(max(n, 2), 0.3 * (b - 4))[:n]
    rng = np.random.default_rng(4)
    obst = np.zeros((3, 12), abi.OBST_DTYPE)
    obst[0, :12] = scenes.make_obstacles(rng, 12, 6.0)
    obst[1, :5] = scenes.make_obstacles(rng, 5, 3.0, inflated=True)
    hb0 = abi.HostBatch(poses, np.array(n_list, np.int32), obst, np.array([12, 5, 0], np.int32),
                        scene_id=np.array([0, 1, 2, 0, 1, 2, 0, 1, 0], np.int32))
    args = abi.make_args(5, 4, True, 100.0, 1.0, True)
    g = _gpu(hb0, p)
    hg = hb0.copy()
    g.optimize(hg, args)
    g.close()
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC)
    assert np.array_equal(hg.status, ha.status)
    assert hg.status[-1] & abi.TEB_STATUS_TOO_FEW_POSES and not (hg.status[-1] & abi.TEB_STATUS_OPTIMIZED)
    assert np.isinf(hg.cost[-1])
    d = _pose_diff(hg, ha)
    assert d.max() < 1e-6, d
    assert np.allclose(hg.cost[:-1], ha.cost[:-1], rtol=1e-6)


def test_full_size_properties_c3():
    """BASELINE C3 at full size (B = 128, n = 200, M = 64 inflated, car-like): size-independent properties"""
    p, hb0 = scenes.make_config_batch("C3", requests=1, seed=0)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    g = _gpu(hb0, p)
    h1 = hb0.copy()
    g.optimize(h1, args)
    h2 = hb0.copy()
    g.optimize(h2, args)
    # determinism: bitwise identical on repeat
    assert np.array_equal(h1.poses, h2.poses) and np.array_equal(h1.cost, h2.cost)
    # permutation invariance of the batch axis (bands are independent)
    perm = np.random.default_rng(0).permutation(hb0.B)
    hp = abi.HostBatch(hb0.poses[perm], hb0.n[perm], hb0.obstacles, hb0.obst_count, hb0.scene_id[perm])
    g.optimize(hp, args)
    assert np.array_equal(hp.poses, h1.poses[perm]) and np.array_equal(hp.cost, h1.cost[perm])
    # chi2 of the final state is not larger than the chi2 of the initial state of the LAST outer iteration... at
    # least: finite, optimized, endpoints fixed, dt positive
    assert np.all(h1.status & abi.TEB_STATUS_OPTIMIZED) and np.all(np.isfinite(h1.cost)) and np.all(np.isfinite(h1.poses))
    assert np.array_equal(h1.poses[:, 0, :3], hb0.poses[:, 0, :3]) and np.array_equal(h1.poses[:, 199, :3], hb0.poses[:, 199, :3])
    assert np.all(h1.lm_iters == 20)
    # accepted LM steps never increase chi2 under the frozen graph: chi2 after one inner iteration (reported through
    # batch.chi2 = currentChi) <= chi2 at the linearisation point (kernel A's evaluation)
    _, chi_before = g.build_system(hb0, 0)
    hs = hb0.copy()
    g.optimize(hs, abi.make_args(1, 1, False))
    assert np.all(hs.chi2 <= chi_before * (1 + 1e-12))
    assert np.any(hs.chi2 < chi_before)
    g.close()


@pytest.mark.parametrize("B,n,M,moving,autosize", [(64, 50, 8, False, False), (96, 400, 256, False, False), (4096, 50, 8, True, False),
                                                   (64, 400, 256, True, True), (512, 150, 32, True, False)])
def test_c5_sweep_corners(oracle, B, n, M, moving, autosize):
    """corners of the BASELINE C5 sweep (B 64-4096 x n 50-400 x M 8-256) and C4 at full size (B = 512): capacity paths
    (13 tiles per band, 4-word association masks, speculation width fallback), two bands against the oracle, the rest
    through size-independent properties"""
    p = scenes.config_params("C4" if moving else "C3")
    p.teb_autosize = int(autosize)
    n_cap = min(512, n + 112) if autosize else n
    hb0 = scenes.make_batch(n, M, candidates=min(B, 32), requests=max(1, B // 32), seed=17, inflated=not moving, moving=moving,
                            via_points=4 if moving else 0, n_cap=n_cap)
    assert hb0.B == B
    args = abi.make_args(3, 2, True, 100.0, 1.0, False)
    g = _gpu(hb0, p)
    hg = hb0.copy()
    g.optimize(hg, args)
    h2 = hb0.copy()
    g.optimize(h2, args)
    g.close()
    assert np.array_equal(hg.poses, h2.poses) and np.array_equal(hg.cost, h2.cost)          # deterministic
    assert np.all(hg.status & abi.TEB_STATUS_OPTIMIZED) and np.all(np.isfinite(hg.cost))
    for b in range(B):
        assert np.array_equal(hg.poses[b, 0, :3], hb0.poses[b, 0, :3])
        assert np.array_equal(hg.poses[b, hg.n[b] - 1, :3], hb0.poses[b, hb0.n[b] - 1, :3])
    assert np.all(np.isfinite(hg.poses))
    for b in (0, B - 1):
        s = hb0.scene_id[b]
        ref, cost, st = oracle.optimize_band(p, hb0.poses[b], hb0.n[b], hb0.obstacles[s][:hb0.obst_count[s]],
                                             via=hb0.via[b] if hb0.V_cap else None, args=args, jac_mode=oracle.JAC_ANALYTIC,
                                             n_cap=n_cap)
        assert len(ref) == hg.n[b]
        assert np.abs(ref - hg.poses[b, :hg.n[b]]).max() < 1e-6
        assert abs(cost - hg.cost[b]) <= 1e-6 * max(abs(cost), 1.0)


@pytest.mark.parametrize("legacy,vor", [(1, 0.0), (0, 2.0), (1, 2.0)])
def test_legacy_association_and_velocity_obstacle_ratio(oracle, legacy, vor):
    """AddEdgesObstaclesLegacy (optimal_planner.cpp:551-643, incl. the triple edge on the centre pose) and
    EdgeVelocityObstacleRatio (edge_velocity_obstacle_ratio.h:82-122): kernel-A system and full optimisation vs oracle"""
    p, hb0 = scenes.make_config_batch("C3", candidates=6, seed=4)
    p.legacy_obstacle_association = legacy
    p.obstacle_poses_affected = 10
    p.weight_velocity_obstacle_ratio = vor
    p.obstacle_proximity_lower_bound, p.obstacle_proximity_upper_bound = 0.2, 1.0
    g = _gpu(hb0, p)
    Hb, chi2 = g.build_system(hb0, 1)
    for b in range(hb0.B):
        n = hb0.n[b]
        Hd, bd, c2 = oracle.build_system(p, hb0.poses[b], n, hb0.obstacles[0][:hb0.obst_count[0]], weight_multiplier=2.0,
                                         jac_mode=oracle.JAC_ANALYTIC)
        ref = _padded_from_dense(Hd, bd, n)
        got = Hb[b, :4 * n]
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
        assert abs(chi2[b] - c2) <= 1e-12 * max(c2, 1.0)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    hg = hb0.copy()
    g.optimize(hg, args)
    g.close()
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=8)
    assert _pose_diff(hg, ha).max() < 1e-6
    assert np.allclose(hg.cost, ha.cost, rtol=1e-6) and np.array_equal(hg.lm_iters, ha.lm_iters)


@pytest.mark.parametrize("acc_lim_y,max_vel_trans", [(0.5, 0.45), (0.0, 0.0), (0.4, 0.0)])
def test_holonomic_edges(oracle, acc_lim_y, max_vel_trans):
    """EdgeVelocityHolonomic (edge_velocity.h:236-273, bounds coupled through max_vel_trans) and
    EdgeAccelerationHolonomic / Start / Goal (edge_acceleration.h:487-712); with acc_lim_y == 0 the reference keeps the
    non-holonomic acceleration edges next to the holonomic velocity edge (optimal_planner.cpp:778)"""
    p, hb0 = scenes.make_config_batch("C3", candidates=6, seed=8)
    p.max_vel_y, p.acc_lim_y, p.max_vel_trans = 0.3, acc_lim_y, max_vel_trans
    p.weight_kinematics_nh = 1.0
    p.weight_max_vel_y, p.weight_acc_lim_y = 2.0, 1.5
    rng = np.random.default_rng(3)
    for b in range(hb0.B):                                   # sideways motion and limits in play, away from the kinks
        n = hb0.n[b]
        hb0.poses[b, 1:n - 1, 2] += rng.normal(0, 0.35, n - 2)
        hb0.poses[b, :n - 1, 3] *= 0.8
    hb0.vel_start[:, 0], hb0.vel_start[:, 1], hb0.vel_start[:, 2] = 0.25, 0.1, -0.1
    hb0.vel_goal[1, 3] = 0.0
    hb0.vel_goal[2, 1] = 0.15
    g = _gpu(hb0, p)
    for variant in (0, 1):
        g.set_linearize_variant(variant)
        Hb, chi2 = g.build_system(hb0, 1)
        for b in range(hb0.B):
            n = hb0.n[b]
            Hd, bd, c2 = oracle.build_system(p, hb0.poses[b], n, hb0.obstacles[0][:hb0.obst_count[0]],
                                             vel_start=hb0.vel_start[b], vel_goal=hb0.vel_goal[b], weight_multiplier=2.0,
                                             jac_mode=oracle.JAC_ANALYTIC)
            ref = _padded_from_dense(Hd, bd, n)
            got = Hb[b, :4 * n]
            assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
            assert abs(chi2[b] - c2) <= 1e-12 * max(c2, 1.0)
    g.set_linearize_variant(0)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    hg = hb0.copy()
    g.optimize(hg, args)
    g.close()
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=8)
    assert np.array_equal(hg.n, ha.n)
    assert _pose_diff(hg, ha).max() < 1e-6
    assert np.allclose(hg.cost, ha.cost, rtol=1e-6) and np.array_equal(hg.lm_iters, ha.lm_iters)
    hn = hb0.copy()
    oracle.optimize_batch(p, hn, args, jac_mode=oracle.JAC_G2O, threads=8)
    dn = _pose_diff(hg, hn)
    assert np.median(dn) <= 1e-4, dn


@pytest.mark.parametrize("footprint,legacy,vor", [("point", 0, 0.0), ("line", 0, 0.0), ("polygon", 0, 0.0), ("two_circles", 0, 0.0),
                                                  ("circular", 1, 0.0), ("polygon", 1, 0.0), ("line", 0, 2.0)])
def test_shape_obstacles_and_footprints(oracle, footprint, legacy, vor):
    """Line / Pill / Polygon obstacles (obstacles.h:597-1045; static and moving) with every footprint model incl.
    Line / Polygon (robot_footprint_model.h:439-760): kernel-A system and the full optimisation against the oracle"""
    p, hb0 = scenes.make_config_batch("C4", candidates=6, seed=6)
    hb0 = scenes.add_shape_obstacles(hb0, seed=1)
    hb0.obstacles["dynamic"][0, ::3] = 0           # a mix of static and moving shapes
    p.legacy_obstacle_association, p.obstacle_poses_affected = legacy, 8
    p.weight_velocity_obstacle_ratio = vor
    p.obstacle_proximity_lower_bound, p.obstacle_proximity_upper_bound = 0.2, 1.0
    if footprint == "line":
        scenes.set_line_footprint(p)
    elif footprint == "polygon":
        scenes.set_polygon_footprint(p)
    elif footprint == "circular":
        p.footprint_type, p.footprint_radius = abi.TEB_FOOTPRINT_CIRCULAR, 0.2
    elif footprint == "two_circles":
        p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
        p.footprint_front_offset, p.footprint_front_radius, p.footprint_rear_offset, p.footprint_rear_radius = 0.3, 0.15, 0.2, 0.2
    g = _gpu(hb0, p)
    obst = hb0.obstacles[0][:hb0.obst_count[0]]
    for variant in (0, 1):
        g.set_linearize_variant(variant)
        Hb, chi2 = g.build_system(hb0, 1)
        for b in range(hb0.B):
            n = hb0.n[b]
            Hd, bd, c2 = oracle.build_system(p, hb0.poses[b], n, obst, via=hb0.via[b], weight_multiplier=2.0,
                                             jac_mode=oracle.JAC_ANALYTIC, obst_vertices=hb0.obst_vertices[0])
            ref = _padded_from_dense(Hd, bd, n)
            got = Hb[b, :4 * n]
            assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
            assert abs(chi2[b] - c2) <= 1e-11 * max(c2, 1.0)
    g.set_linearize_variant(0)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    outs = {}
    for solver in (2, 0):
        g.set_solver(solver)
        h = hb0.copy()
        g.optimize(h, args)
        outs[solver] = h
    g.close()
    hg = outs[2]
    assert _pose_diff(outs[0], hg).max() < 1e-7
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=8)
    assert np.array_equal(hg.n, ha.n)
    d = _pose_diff(hg, ha)
    print("shape parity", footprint, legacy, vor, "max", d.max(), "bands > 1e-6:", np.where(d > 1e-6)[0].tolist())
    assert d.max() < 1e-6, d
    assert np.all(hg.status & abi.TEB_STATUS_OPTIMIZED)


def test_bad_obstacle_rows_are_reported():
    p, hb = scenes.make_config_batch("C1", candidates=2)
    hb = scenes.add_shape_obstacles(hb, seed=2)
    g = _gpu(hb, p)
    bad = hb.copy()
    bad.obstacles["vertex_count"][0, 0] = 99                 # vertex range outside the pool
    with pytest.raises(T.TebGpuError, match="rc=-1"):
        g.optimize(bad, abi.make_args())
    with pytest.raises(T.TebGpuError, match="rc=-3"):
        q = abi.default_params()
        q.footprint_type = 9
        g.set_params(q)
    g.close()


@pytest.mark.parametrize("cfg,use_dt", [("C1", True), ("C2", True), ("C3", True), ("C4", True), ("C4", False)])
def test_h_signatures_match_oracle(oracle, cfg, use_dt):
    """calculateEquivalenceClass for a batch of candidates: 2-D signature (static scenes) against the long double
    oracle to 1e-9 relative, x-y-t signature (include_dynamic_obstacles) to 1e-12 absolute; equal candidates map to
    equal classes with the reference's isEqual rules (h_signature.h:191-207, :366-388)"""
    p, hb = scenes.make_config_batch(cfg, candidates=12, seed=13)
    if cfg != "C4":
        p.include_dynamic_obstacles = 0
    g = _gpu(hb, p)
    got = g.h_signature(hb, use_timediffs=use_dt)
    g.close()
    obst = hb.obstacles[0][:hb.obst_count[0]]
    ref = [oracle.h_signature(p, hb.poses[b], hb.n[b], obst, use_timediffs=use_dt) for b in range(hb.B)]
    if p.include_dynamic_obstacles:
        ref = np.array(ref)
        assert np.abs(got[:, :len(obst)] - ref).max() < 1e-12
        def equal(x, y):
            keep = (np.abs(x) >= p.h_signature_threshold) & (np.abs(y) >= p.h_signature_threshold)
            return bool(np.all(np.sign(x[keep]) == np.sign(y[keep])))
        classes_g = [[equal(got[i, :len(obst)], got[j, :len(obst)]) for j in range(hb.B)] for i in range(hb.B)]
        classes_r = [[equal(ref[i], ref[j]) for j in range(hb.B)] for i in range(hb.B)]
    else:
        ref = np.array(ref)
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
        eq = lambda x, y: abs(x.real - y.real) <= p.h_signature_threshold and abs(x.imag - y.imag) <= p.h_signature_threshold
        classes_g = [[eq(got[i], got[j]) for j in range(hb.B)] for i in range(hb.B)]
        classes_r = [[eq(ref[i], ref[j]) for j in range(hb.B)] for i in range(hb.B)]
    assert classes_g == classes_r
    # with 20+ obstacles the 2-D signature shrinks below h_signature_threshold (the product in A_l, h_signature.h:149-163)
    # and every candidate falls into one class - in the reference as well; the small scene and the x-y-t variant separate
    if cfg in ("C1", "C4"):
        assert not all(all(r) for r in classes_r)


def test_solvers_and_speculation_widths_agree():
    """the three linear solvers give the same bands up to round-off; the speculation width K does not change a bit"""
    p, hb0 = scenes.make_config_batch("C4", candidates=12, seed=9)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    g = _gpu(hb0, p)
    outs = {}
    for name, solver, k in (("spec4", 2, 4), ("spec2", 2, 2), ("spec6", 2, 6), ("spec8", 2, 8), ("bcr", 1, 0), ("seq", 0, 0)):
        g.set_solver(solver)
        g.set_speculation(k)
        h = hb0.copy()
        g.optimize(h, args)
        outs[name] = h
    g.close()
    for name in ("spec2", "spec6", "spec8"):
        assert np.array_equal(outs[name].poses, outs["spec4"].poses) and np.array_equal(outs[name].cost, outs["spec4"].cost)
        assert np.array_equal(outs[name].lm_iters, outs["spec4"].lm_iters)
    for name in ("bcr", "seq"):
        assert _pose_diff(outs[name], outs["spec4"]).max() < 1e-7
        assert np.allclose(outs[name].cost, outs["spec4"].cost, rtol=1e-7)
        assert np.array_equal(outs[name].status, outs["spec4"].status)


def test_errors_are_loud():
    p, hb = scenes.make_config_batch("C1", candidates=2)
    g = T.TebGpu(2, 50, 1, 8, 0)
    q = abi.default_params()
    q.footprint_type = 7
    with pytest.raises(T.TebGpuError, match="rc=-3"):
        g.set_params(q)                                   # unknown footprint model: refused, not approximated
    p2, big = scenes.make_config_batch("C2", candidates=4)
    with pytest.raises(T.TebGpuError, match="rc=-4"):
        g.optimize(big, abi.make_args())                  # exceeds the context limits
    g.close()
    with pytest.raises(T.TebGpuError):
        T.TebGpu(1, 1024, 1, 8, 0)                         # max_poses > 512


def test_device_pointer_entry_point_matches_host_entry_point():
    import torch
    p, hb0 = scenes.make_config_batch("C2", candidates=8, seed=3)
    args = abi.make_args(5, 4, True, 100.0, 1.0, False)
    g = _gpu(hb0, p)
    ref = hb0.copy()
    g.optimize(ref, args)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    poses, n, sid = t(hb0.poses), t(hb0.n), t(hb0.scene_id)
    ob, oc = t(hb0.obstacles), t(hb0.obst_count)
    vs, vg = t(hb0.vel_start), t(hb0.vel_goal)
    cost = torch.zeros(hb0.B, dtype=torch.float64, device=dev)
    chi2 = torch.zeros_like(cost)
    status = torch.zeros(hb0.B, dtype=torch.int32, device=dev)
    iters = torch.zeros_like(status)
    bs = abi.TebBatch()
    bs.B, bs.n_cap, bs.S, bs.M_cap, bs.V_cap = hb0.B, hb0.n_cap, hb0.S, hb0.M_cap, 0
    bs.poses, bs.n, bs.scene_id = poses.data_ptr(), n.data_ptr(), sid.data_ptr()
    bs.obstacles, bs.obst_count = ob.data_ptr(), oc.data_ptr()
    bs.vel_start, bs.vel_goal = vs.data_ptr(), vg.data_ptr()
    bs.cost, bs.chi2, bs.status, bs.lm_iters = cost.data_ptr(), chi2.data_ptr(), status.data_ptr(), iters.data_ptr()
    g.optimize_device(bs, args)
    g.synchronize()
    got = poses.cpu().numpy().view(np.float64).reshape(hb0.poses.shape)
    assert np.array_equal(got, ref.poses)
    assert np.array_equal(cost.cpu().numpy(), ref.cost)
    # begin + outer*(build + inner*(A + 2 rounds x (solve, eval) [K=6]) + (inner-1) side-stream A of the retry list)
    # + 3 outer boundaries x (side-stream build + A of the retry list) + finalize
    assert g.launch_count() == 1 + 4 * (1 + 5 * (1 + 2 * 2) + 4) + 3 * 2 + 1
    g.close()
