"""tools/twisted_model.py executes the index arithmetic of k_solve_lat (twisted banded LDL^T by one warp: ownership of the
window columns, merge of the two sweeps in the middle block, ring back substitution) lane by lane in numpy. It is the CPU
check of that mapping; the kernel itself is compared with the thread-per-system solver and the oracle in the GPU tests."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("twisted_model", os.path.join(ROOT, "tools", "twisted_model.py"))
tm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tm)


@pytest.mark.parametrize("N", [12, 16, 20, 24, 28, 36, 48, 100, 400])
def test_twisted_factorisation_solves_the_banded_system(N):
    for seed in range(2):
        A, b, Hs = tm.make_system(N, seed)
        x = tm.solve_lat(Hs, N)
        ref = np.linalg.solve(A, b)
        assert np.abs(x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
