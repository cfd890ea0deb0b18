"""GPU parity against the REFERENCE's results and the branches round 1 left untested (run on the B200 box: pytest -m gpu).

Two checkers:
  * tests/golden/golden_ref_v1.npz - final bands / costs of whole optimizeTEB calls produced by the reference's own code
    (oracle/_ref, generator tests/golden/make_golden_ref.py) for the 24 feature scenarios of tests/scenarios.py. The
    reference linearises numerically (central differences, delta 1e-9), the kernels in closed form, so the comparison
    uses the north-star tolerance 1e-4 per pose component; the measured fractions are written to
    gpurun_out/parity_report.json and checked against the closed-form oracle run band by band.
  * the oracle in closed-form mode (same algorithm as the kernels): 1e-6 per pose component, identical n / LM iteration
    counts / status - for the branches switched on by the scenarios (divergence detection, ordered via-points, ...)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import teb_local_planner_b200 as T
from teb_local_planner_b200 import abi, scenes
from tests import scenarios

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_REF = os.path.join(HERE, "golden", "golden_ref_v1.npz")
REPORT = os.path.join(os.path.dirname(HERE), "gpurun_out", "parity_report.json")


def _gpu(hb, p):
    g = T.TebGpu(hb.B, hb.n_cap, hb.S, max(hb.M_cap, 1), hb.V_cap, max_obst_vertices=hb.PV_cap)
    g.set_params(p)
    return g


def _args(p):
    return abi.make_args(p.no_inner_iterations, p.no_outer_iterations, True, p.selection_obst_cost_scale,
                         p.selection_viapoint_cost_scale, False)


def _diff(poses_a, n_a, poses_b, n_b):
    return np.array([np.abs(poses_a[b, :n_a[b]] - poses_b[b, :n_b[b]]).max() if n_a[b] == n_b[b] else np.inf for b in range(len(n_a))])


@pytest.mark.parametrize("name", scenarios.ALL)
def test_scenarios_match_closed_form_oracle_and_reference_golden(oracle, name):
    p, hb0 = scenarios.scenario(name)
    args = _args(p)
    g = _gpu(hb0, p)
    hg = hb0.copy()
    g.optimize(hg, args)
    g.close()
    # (1) same algorithm: the oracle with closed-form Jacobians
    ha = hb0.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=4)
    assert np.array_equal(hg.n, ha.n)
    da = _diff(hg.poses, hg.n, ha.poses, ha.n)
    assert da.max() < 1e-6, (name, da)
    assert np.allclose(hg.cost, ha.cost, rtol=1e-6) and np.array_equal(hg.lm_iters, ha.lm_iters)
    assert np.array_equal(hg.status, ha.status)
    # (2) the reference's own results (numeric Jacobians): north-star tolerance, measured and reported
    z = np.load(GOLDEN_REF, allow_pickle=False)
    n_ref, poses_ref = z[name + "/n"], z[name + "/poses"]
    dn = _diff(hg.poses, hg.n, poses_ref, n_ref)
    spread = _diff(ha.poses, ha.n, poses_ref, n_ref)   # closed-form vs numeric Jacobians inside the CPU implementations
    within = dn <= 1e-4
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        rep = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        rep[name] = {"bands": int(len(dn)), "fraction_within_1e-4_of_reference": float(within.mean()),
                     "median_abs_pose_diff": float(np.median(dn[np.isfinite(dn)])) if np.isfinite(dn).any() else None,
                     "max_abs_pose_diff": float(dn.max()), "n_equal": bool(np.array_equal(hg.n, n_ref)),
                     "cost_rel_diff_median": float(np.median(np.abs(hg.cost - z[name + "/cost"]) / np.maximum(1.0, np.abs(z[name + "/cost"]))))}
        json.dump(rep, open(REPORT, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
    # every band outside the tolerance must be one where the CPU implementation's own two Jacobian modes part ways
    assert np.all(within | (spread > 1e-5)), (name, dn, spread)
    assert np.median(dn) <= 1e-4, (name, dn)


def test_divergence_detection_changes_the_cost_source_and_reports_chi2(oracle):
    """recovery.divergence_detection_enable (optimal_planner.cpp:331, :1023-1039): batch statistics recompute the errors
    at the accepted state after every iteration, so computeCurrentCost sees those instead of the last trial's; chi2 of
    the final state is what hasDiverged compares"""
    p, hb0 = scenarios.scenario("divergence", candidates=16)
    args = _args(p)
    outs = {}
    for flag in (0, 1):
        p.divergence_detection_enable = flag
        g = _gpu(hb0, p)
        h = hb0.copy()
        g.optimize(h, args)
        g.close()
        ha = hb0.copy()
        oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=4)
        assert _diff(h.poses, h.n, ha.poses, ha.n).max() < 1e-6
        assert np.allclose(h.cost, ha.cost, rtol=1e-6) and np.allclose(h.chi2, ha.chi2, rtol=1e-6)
        outs[flag] = h
    assert np.array_equal(outs[0].poses, outs[1].poses)          # the flag never changes the optimisation itself
    # ... and it can only change the cost of a band whose last LM iteration ended on a REJECTED trial (all ten rejected:
    # Terminate), because an accepted trial is always the last one evaluated
    ended_on_accept = (outs[0].status & abi.TEB_STATUS_TERMINATED) == 0
    assert np.array_equal(outs[0].cost[ended_on_accept], outs[1].cost[ended_on_accept])
    thr = float(np.median(outs[1].chi2))
    assert (outs[1].chi2 > thr).any() and (outs[1].chi2 <= thr).any()   # hasDiverged would be true for some, false for others


def test_optimization_deactivated_and_capacity_overflow():
    p, hb0 = scenes.make_config_batch("C2", candidates=4, seed=3)
    args = _args(p)
    p.optimization_activate = 0            # optimizeTEB returns false at once (optimal_planner.cpp:185)
    g = _gpu(hb0, p)
    h = hb0.copy()
    g.optimize(h, args)
    assert np.array_equal(h.poses, hb0.poses) and np.all(np.isinf(h.cost))
    assert np.all(h.status & abi.TEB_STATUS_DISABLED) and not np.any(h.status & abi.TEB_STATUS_OPTIMIZED)
    g.close()
    # autoResize would need more records than the batch provides: reported per band, others unaffected
    p2, hb1 = scenes.make_config_batch("C2", candidates=4, seed=3, autosize=True)
    tight = abi.HostBatch(hb1.poses[:, :hb1.n.max() + 12].copy(), hb1.n, hb1.obstacles, hb1.obst_count, hb1.scene_id)
    tight.poses[0, :tight.n[0] - 1, 3] *= 6.0     # band 0: time differences far above dt_ref -> far more than 12 insertions
    g = _gpu(tight, p2)
    h = tight.copy()
    g.optimize(h, args)
    g.close()
    assert h.status[0] & abi.TEB_STATUS_CAPACITY and not (h.status[0] & abi.TEB_STATUS_OPTIMIZED)
    assert np.all(h.status[1:] & abi.TEB_STATUS_OPTIMIZED)


def test_bad_band_inputs_are_rejected_or_reported():
    p, hb = scenes.make_config_batch("C1", candidates=3, seed=1)
    g = _gpu(hb, p)
    args = _args(p)
    for field, value in (("n", hb.n_cap + 1), ("scene_id", 5), ("n", -1)):
        h = hb.copy()
        getattr(h, field)[1] = value
        with pytest.raises(T.TebGpuError):
            g.optimize(h, args)
    h = hb.copy()
    h.n[1] = 0                             # an empty band: the reference's optimizeGraph returns false (optimal_planner.cpp:377)
    g.optimize(h, args)
    assert h.status[1] & abi.TEB_STATUS_BAD_INPUT and not (h.status[1] & abi.TEB_STATUS_OPTIMIZED)
    assert (h.status[0] & abi.TEB_STATUS_OPTIMIZED) and (h.status[2] & abi.TEB_STATUS_OPTIMIZED)
    g.close()


def test_gather_costs_single_rank_is_identity():
    """tebgpu_gather_costs without a communicator (one rank): both pointer modes return the local vector"""
    import torch
    p, hb = scenes.make_config_batch("C1", candidates=5, seed=2)
    g = _gpu(hb, p)
    h = hb.copy()
    allc = g.optimize_gather(h, _args(p))
    assert np.array_equal(allc, h.cost) and g.info(3) == 1 and g.info(4) == 0
    assert np.array_equal(g.gather_costs(h.cost), h.cost)
    d_in = torch.from_numpy(h.cost).cuda()
    d_out = torch.zeros_like(d_in)
    g.gather_costs_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr())
    g.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), h.cost)
    g.close()


def test_compute_cost_outside_optimize_is_the_scaled_chi2_at_the_state(oracle):
    """tebgpu_compute_cost (the reference's value on that path is undefined, tests/test_reference_pin.py): scaled chi2
    by family at the current state"""
    from tests.test_reference_pin import _cost_from_oracle
    for name in ("C2", "C4", "all_edge_families"):
        p, hb = scenarios.scenario(name)
        args = abi.make_args(5, 4, True, 50.0, 2.5, False)
        g = _gpu(hb, p)
        h = hb.copy()
        g.compute_cost(h, args)
        g.close()
        for b in range(hb.B):
            want = _cost_from_oracle(oracle, p, hb, b, args)
            assert abs(h.cost[b] - want) <= 1e-9 * max(1.0, abs(want)), (name, b, h.cost[b], want)


def test_cuda_graph_replay_gives_identical_results():
    """tebgpu_set_graph: the captured launch sequence (retry side stream included) replays to the same bits as direct
    launches; one graph per distinct key, reused across calls"""
    p, hb0 = scenes.make_config_batch("C3", candidates=32, seed=12, autosize=True)
    args = _args(p)
    g = _gpu(hb0, p)
    outs = {}
    for mode in (0, 1):
        g.set_graph(mode)
        runs = []
        for _ in range(3):
            h = hb0.copy()
            g.optimize(h, args)
            runs.append(h)
        assert all(np.array_equal(runs[0].poses, r.poses) and np.array_equal(runs[0].cost, r.cost) for r in runs[1:])
        outs[mode] = runs[0]
    assert g.info(6) == 1 and g.info(5) == 1       # one captured graph, replayed three times
    assert np.array_equal(outs[0].poses, outs[1].poses) and np.array_equal(outs[0].n, outs[1].n)
    assert np.array_equal(outs[0].cost, outs[1].cost) and np.array_equal(outs[0].lm_iters, outs[1].lm_iters)
    p.weight_optimaltime = 2.0                      # new parameters -> new key -> second graph, different result
    g.set_params(p)
    h = hb0.copy()
    g.optimize(h, args)
    assert g.info(6) == 2 and not np.array_equal(h.cost, outs[1].cost)
    g.close()


@pytest.mark.parametrize("cfg,autosize", [("C1", True), ("C3", False), ("C4", True)])
def test_warp_solver_is_bit_identical_to_thread_solver(cfg, autosize):
    """k_solve_warp (one warp per system) and k_solve_tpb (one thread per system) run the same factorisation in the same
    operation order: final bands, costs and LM iteration counts must agree bit for bit"""
    p, hb0 = scenes.make_config_batch(cfg, candidates=32, seed=19, autosize=autosize)
    args = _args(p)
    g = _gpu(hb0, p)
    g.set_graph(0)
    outs = {}
    for mode in (0, 1):
        g.set_warp_solver(mode)
        h = hb0.copy()
        g.optimize(h, args)
        outs[mode] = h
    g.close()
    assert np.array_equal(outs[0].n, outs[1].n) and np.array_equal(outs[0].lm_iters, outs[1].lm_iters)
    assert np.array_equal(outs[0].poses, outs[1].poses) and np.array_equal(outs[0].cost, outs[1].cost)
    assert np.array_equal(outs[0].status, outs[1].status)


@pytest.mark.parametrize("cfg,autosize,cands", [("C1", True, 32), ("C2", False, 32), ("C3", False, 32), ("C3", True, 5), ("C4", True, 48)])
def test_latency_solver_matches_thread_solver(cfg, autosize, cands):
    """k_solve_lat (twisted factorisation by one warp, system resident in shared memory) eliminates in another order than
    k_solve_tpb: same LM decisions, iteration counts and status, trajectories and costs equal to rounding"""
    p, hb0 = scenes.make_config_batch(cfg, candidates=cands, seed=23, autosize=autosize)
    args = _args(p)
    g = _gpu(hb0, p)
    outs = {}
    for mode in (0, 3):
        g.set_warp_solver(mode)
        h = hb0.copy()
        g.optimize(h, args)
        outs[mode] = h
    g.close()
    a, b = outs[0], outs[3]
    assert np.array_equal(a.n, b.n) and np.array_equal(a.lm_iters, b.lm_iters) and np.array_equal(a.status, b.status)
    d = _diff(a.poses, a.n, b.poses, b.n)
    assert d.max() < 1e-8, d
    assert np.allclose(a.cost, b.cost, rtol=1e-8)


@pytest.mark.parametrize("mode", [0, 3])
def test_short_bands_with_both_solvers(oracle, mode):
    """3 .. 9 poses: for k_solve_lat the middle block is most of the system, the bottom sweep has zero to a few pivots"""
    p = abi.default_params()
    p.teb_autosize = 0
    for n in (3, 4, 5, 6, 7, 9):
        hb = scenes.make_batch(n, 4, 6, seed=40 + n, n_cap=n + 4, amp=0.5)
        args = _args(p)
        g = _gpu(hb, p)
        g.set_warp_solver(mode)
        hg = hb.copy()
        g.optimize(hg, args)
        g.close()
        ha = hb.copy()
        oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=2)
        d = _diff(hg.poses, hg.n, ha.poses, ha.n)
        print("short bands: mode", mode, "n", n, "max diff", d.max(), "lm_iters", hg.lm_iters.tolist(), ha.lm_iters.tolist())
        assert np.array_equal(hg.n, ha.n) and d.max() < 1e-6, n
        # a converged 3-pose band can stop an inner loop one step earlier or later: g2o terminates on rho == 0 (trial
        # chi2 EXACTLY equal to the current one), a tie that the last bit decides; the trajectories then agree to 1e-8
        same = hg.lm_iters == ha.lm_iters
        assert np.all(same | (d < 1e-8)), (n, hg.lm_iters, ha.lm_iters, d)


@pytest.mark.parametrize("cfg,autosize,cands", [("C1", True, 8), ("C3", False, 32), ("C4", True, 24), ("shapes", False, 0)])
def test_per_trial_evaluation_kernel_is_bit_identical(cfg, autosize, cands, monkeypatch):
    """k_trial_eval3 (a CTA per (band, trial), last CTA of a band decides) against k_trial_eval2 (a CTA per band): same tile
    tasks, same fold order, same accept / reject replay - every output must agree bit for bit, retry rounds included
    (speculation width 2 makes them frequent)"""
    if cfg == "shapes":
        p, hb0 = scenarios.scenario("shapes_polygon")
    else:
        p, hb0 = scenes.make_config_batch(cfg, candidates=cands, seed=31, autosize=autosize)
    args = _args(p)
    outs = {}
    for spec in (0, 2):
        for mode in ("0", "1"):
            monkeypatch.setenv("TEBGPU_EVAL3", mode)
            g = _gpu(hb0, p)
            g.set_speculation(spec)
            h = hb0.copy()
            g.optimize(h, args)
            g.close()
            outs[(spec, mode)] = h
    ref = outs[(0, "0")]
    for key, h in outs.items():
        assert np.array_equal(ref.n, h.n) and np.array_equal(ref.lm_iters, h.lm_iters) and np.array_equal(ref.status, h.status), key
        assert np.array_equal(ref.poses, h.poses) and np.array_equal(ref.cost, h.cost), key


@pytest.mark.parametrize("cfg,autosize", [("C1", True), ("C2", False)])
def test_split_round0_schedule_is_bit_identical(cfg, autosize, monkeypatch):
    """throughput regime (> 2368 bands), TEBGPU_OVERLAP=1: the retry rounds run on the side stream; with TEBGPU_SPLIT the side
    stream also solves / evaluates round 0 for the bands it linearised while the main stream handles the others. Only the
    schedule differs: every output must agree bit for bit with the one-stream schedule and with a repeat of itself"""
    p, hb0 = scenes.make_config_batch(cfg, requests=80, candidates=32, seed=41, autosize=autosize)
    assert hb0.B * 8 > 148 * 4 * 32
    args = _args(p)
    outs = []
    for overlap, mode in (("0", "0"), ("1", "0"), ("1", "1"), ("1", "1")):
        monkeypatch.setenv("TEBGPU_OVERLAP", overlap)
        monkeypatch.setenv("TEBGPU_SPLIT", mode)
        g = _gpu(hb0, p)
        h = hb0.copy()
        g.optimize(h, args)
        g.close()
        outs.append(h)
    for h in outs[1:]:
        assert np.array_equal(outs[0].n, h.n) and np.array_equal(outs[0].lm_iters, h.lm_iters) and np.array_equal(outs[0].status, h.status)
        assert np.array_equal(outs[0].poses, h.poses) and np.array_equal(outs[0].cost, h.cost)


def test_many_obstacles_per_scene(oracle):
    """600 point obstacles in one scene (10 association words per pose; round 1 stopped at 256): same association, same
    trajectories as the closed-form oracle"""
    p = abi.default_params()
    p.teb_autosize = 0
    hb = scenes.make_batch(40, 600, 6, seed=77, amp=1.0)
    args = _args(p)
    g = _gpu(hb, p)
    hg = hb.copy()
    g.optimize(hg, args)
    g.close()
    ha = hb.copy()
    oracle.optimize_batch(p, ha, args, jac_mode=oracle.JAC_ANALYTIC, threads=4)
    assert np.array_equal(hg.n, ha.n) and np.array_equal(hg.lm_iters, ha.lm_iters)
    assert _diff(hg.poses, hg.n, ha.poses, ha.n).max() < 1e-6
    assert np.allclose(hg.cost, ha.cost, rtol=1e-6)
