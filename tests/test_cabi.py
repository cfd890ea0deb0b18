"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, struct sizes match,
host-side pure functions behave like the reference. No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from teb_local_planner_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(teblib):
    hdr = open(os.path.join(ROOT, "include", "teb_b200.h")).read()
    names = sorted(set(re.findall(r"\b(tebgpu_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 12
    for nme in names:
        assert hasattr(teblib, nme), f"{nme} declared in include/teb_b200.h but not exported"


def test_struct_sizes_and_defaults(teblib):
    for which, st in enumerate((abi.TebParams, abi.TebObstacle, abi.TebBatch, abi.TebOptimizeArgs, abi.TebGpuLimits)):
        assert teblib.tebgpu_sizeof(which) == C.sizeof(st)
    p = abi.TebParams()
    teblib.tebgpu_default_params(C.byref(p))
    q = abi.default_params()
    assert bytes(p) == bytes(q)
    # TebConfig() ctor defaults teb_config.h:245-390 (spot checks incl. the ctor-vs-cfg differences SURVEY App. C)
    assert p.include_dynamic_obstacles == 1 and p.selection_obst_cost_scale == 100.0
    assert p.no_inner_iterations == 5 and p.no_outer_iterations == 4 and p.weight_kinematics_nh == 1000
    assert abs(p.force_reinit_new_goal_angular - 0.5 * np.pi) < 1e-15


def test_no_gpu_fails_loudly(teblib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lim = abi.TebGpuLimits(4, 50, 1, 8, 0, 0)
    ctx = C.c_void_p()
    rc = teblib.tebgpu_create(C.byref(lim), 0, C.byref(ctx))
    assert rc == abi.TEBGPU_ERR_NO_DEVICE and not ctx.value


def test_select_best_matches_reference_semantics(teblib):
    """homotopy_class_planner.cpp:564-616: hysteresis on the last best, preference on the initial-plan band,
    strict '<' so the first minimum wins."""
    from teb_local_planner_b200 import distributed as D
    cost = np.array([5.0, 4.0, 4.0, 4.1])
    f = lambda *a: teblib.tebgpu_select_best(cost.ctypes.data, len(cost), *a)
    assert f(-1, -1, 1.0, 0.95) == 1
    assert f(3, -1, 0.9, 0.95) == 3            # 4.1 * 0.9 = 3.69 wins through hysteresis
    assert f(-1, 0, 1.0, 0.5) == 0             # 5.0 * 0.5 = 2.5 initial plan preferred
    assert f(2, 2, 0.99, 0.1) == 2             # last_best takes precedence over initial_plan (:599-602)
    assert teblib.tebgpu_select_best(cost.ctypes.data, 0, -1, -1, 1.0, 1.0) == -1
    for args in ((-1, -1, 1.0, 0.95), (3, -1, 0.9, 0.95), (-1, 0, 1.0, 0.5)):
        assert D.select_best(cost, *args) == f(*args)


@pytest.mark.parametrize("case", ["large_at_end", "small_at_end", "middle_and_end"])
def test_host_autoresize_reference_gtests(teblib, oracle, case):
    """test/teb_basics.cpp:5-68 on the product's host/device autoResize routine + bitwise equality with the oracle"""
    dt, hyst = 0.1, 0.1 / 3.0
    dts = [dt] * 9
    if case == "large_at_end":
        dts.append(dt + 2 * hyst)
    elif case == "small_at_end":
        dts.append(dt - 2 * hyst)
    else:
        dts[5] = dt + 2 * hyst
        dts.append(dt - 2 * hyst)
    n = len(dts) + 1
    rec = np.zeros((64, 4))
    rec[:n, 0] = np.arange(n)
    rec[:n - 1, 3] = dts
    ref = oracle.auto_resize(rec, n, dt, hyst, 3, 100, False, n_cap=64)
    nn = teblib.tebgpu_auto_resize_host(rec.ctypes.data, n, 64, dt, hyst, 3, 100, 0)
    assert nn == len(ref)
    d = rec[:nn - 1, 3]
    assert np.all(d <= dt + hyst + 1e-3) and np.all(dt - hyst - 1e-3 <= d)
    assert np.array_equal(rec[:nn], ref)


def test_host_autoresize_random_matches_oracle(teblib, oracle):
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(3, 40))
        rec = np.zeros((256, 4))
        rec[:n, :2] = np.cumsum(rng.uniform(0, 0.5, (n, 2)), axis=0)
        rec[:n, 2] = rng.uniform(-3, 3, n)
        rec[:n - 1, 3] = rng.uniform(0.01, 1.5, n - 1)
        fast = int(rng.integers(0, 2))
        ref = oracle.auto_resize(rec, n, 0.3, 0.1, 3, 500, fast, n_cap=256)
        nn = teblib.tebgpu_auto_resize_host(rec.ctypes.data, n, 256, 0.3, 0.1, 3, 500, fast)
        assert nn == len(ref)
        assert np.allclose(rec[:nn], ref, rtol=0, atol=1e-15)
    # capacity overflow is reported, not silently truncated
    rec = np.array([[0, 0, 0, 100.0], [1, 0, 0, 0]] + [[0, 0, 0, 0]] * 2, float)
    assert teblib.tebgpu_auto_resize_host(rec.ctypes.data, 2, 4, 0.3, 0.1, 3, 500, 0) == abi.TEBGPU_ERR_CAPACITY


def test_shard_range_partitions():
    from teb_local_planner_b200 import distributed as D
    for B in (1, 7, 32, 512, 4097):
        for w in (1, 2, 4, 8):
            rs = [D.shard_range(B, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_select_best_random_against_python_restatement(teblib):
    """tebgpu_select_best vs the Python restatement on random cost vectors, incl. ties, infinities and every combination of
    last-best / initial-plan indices (homotopy_class_planner.cpp:564-616)"""
    from hypothesis import given, settings, strategies as st
    from teb_local_planner_b200 import distributed as D

    costs = st.lists(st.one_of(st.floats(min_value=0.0, max_value=100.0), st.sampled_from([1.0, 2.0, float("inf")])),
                     min_size=1, max_size=12)

    @settings(max_examples=300, deadline=None)
    @given(costs, st.integers(-1, 11), st.integers(-1, 11), st.floats(0.5, 1.0), st.floats(0.1, 1.0))
    def check(c, last, init, hyst, pref):
        c = np.array(c, dtype=np.float64)
        last = last if last < len(c) else -1
        init = init if init < len(c) else -1
        got = teblib.tebgpu_select_best(c.ctypes.data, len(c), last, init, hyst, pref)
        assert got == D.select_best(c, last, init, hyst, pref)

    check()
