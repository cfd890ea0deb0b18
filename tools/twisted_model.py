"""Lane-level model of k_solve_lat (teb_solve_lat.cuh): twisted banded LDL^T of one system by one warp.

The top half-warp eliminates unknowns 0 .. m-1 downwards, the bottom half-warp N-1 .. m+11 upwards, the 11 unknowns in
between are eliminated last by the top half after the two Schur contributions were added. This script executes the same
index arithmetic lane by lane in numpy and compares the solution with a dense solve; it exists to check the mapping
(ownership, merge, the two back-substitution phases) on the CPU, where there is no GPU to run the kernel on."""
import numpy as np

HB = 10


def make_system(N, seed):
    rng = np.random.default_rng(seed)
    A = np.zeros((N, N))
    for r in range(N):
        for k in range(1, HB + 1):
            if r - k >= 0:
                A[r, r - k] = A[r - k, r] = rng.normal() * 0.3
    A += np.eye(N) * (np.abs(A).sum(1).max() + 1.0)
    b = rng.normal(size=N)
    Hs = np.zeros((N, 12))
    for r in range(N):
        for k in range(0, HB + 1):
            if r - k >= 0:
                Hs[r, k] = A[r, r - k]
        Hs[r, 11] = b[r]
    return A, b, Hs


def solve_lat(Hs, N):
    Hs = Hs.copy()
    m = (N - 11 + 1) // 2
    T_bot = N - 11 - m
    assert 0 <= T_bot <= m
    R = np.zeros((32, 11))
    Ry = np.zeros(32)
    cb = np.zeros((2, 2, 12))
    mid = np.zeros((11, 12))
    x = np.full(N, np.nan)

    def slot(h, q):
        return q if h == 0 else N - 1 - q

    def gather(h, q):
        col = np.zeros(11)
        y = 0.0
        if h == 0:
            if q < m + 11:
                for k in range(11):
                    if q + k < m + 11:
                        col[k] = Hs[q + k, k]
                y = Hs[q, 11]
            else:
                col[0] = 1.0
        else:
            c = N - 1 - q
            if c >= m + 11:
                for k in range(11):
                    if c - k >= m:
                        col[k] = Hs[c, k]
                y = Hs[c, 11]
        return col, y

    # prologue: columns 0..10
    for lane in range(32):
        h, mm = lane >> 4, lane & 15
        if mm <= 10:
            R[lane], Ry[lane] = gather(h, mm)

    def step(t, act):
        s = t & 15
        # publish
        for lane in range(32):
            h, mm = lane >> 4, lane & 15
            if mm == s and act[h]:
                cb[h, t & 1, :11] = R[lane]
                cb[h, t & 1, 11] = Ry[lane]
        for lane in range(32):
            h, mm = lane >> 4, lane & 15
            uC = (mm - s) & 15
            if uC == 11:
                R[lane], Ry[lane] = gather(h, t + 11)
        fac = {}
        for lane in range(32):
            h, mm = lane >> 4, lane & 15
            if not act[h]:
                continue
            uC = (mm - s) & 15
            c = cb[h, t & 1]
            d = c[0]
            assert d > 0
            inv = 1.0 / d
            yj = c[11]
            if uC == 0:
                fac[(slot(h, t), 0)] = yj * inv
            if 1 <= uC <= 10:
                lq = c[uC] * inv
                for k in range(0, 11 - uC):
                    R[lane, k] -= c[uC + k] * lq
                Ry[lane] -= yj * lq
                fac[(slot(h, t), uC)] = lq
        for (r, u), v in fac.items():
            Hs[r, u] = v

    for t in range(m):
        step(t, (True, t < T_bot))
    # merge
    mid[:] = 0
    for lane in range(16, 32):
        mm = lane & 15
        # the column of the bottom window owned by this lane: q' = T_bot + i, i = 0..10
        i = (mm - T_bot) & 15
        if i <= 10:
            a = (m + 10) - i
            for k in range(11):
                if a - k >= m:
                    mid[a - m, k] = R[lane, k]
            mid[a - m, 11] = Ry[lane]
    for lane in range(16):
        mm = lane & 15
        i = (mm - m) & 15
        if i <= 10:
            bcol = m + i
            for k in range(11):
                if bcol + k <= m + 10:
                    R[lane, k] += mid[bcol + k - m, k]
            Ry[lane] += mid[bcol - m, 11]
    for t in range(m, m + 11):
        step(t, (True, False))

    # back substitution, dot-product form with a ring of the ten most recent solutions (every lane of a half computes
    # the same values; the model keeps one copy per half)
    def row_solve(row, W, P):
        f = Hs[row]
        a = f[0] - f[10] * W[(P + 10) % 10]
        b2 = -f[9] * W[(P + 9) % 10]
        c2 = -f[8] * W[(P + 8) % 10]
        d2 = -f[7] * W[(P + 7) % 10]
        a -= f[6] * W[(P + 6) % 10]
        b2 -= f[5] * W[(P + 5) % 10]
        c2 -= f[4] * W[(P + 4) % 10]
        d2 -= f[3] * W[(P + 3) % 10]
        a -= f[2] * W[(P + 2) % 10]
        return ((a + b2) + (c2 + d2)) - f[1] * W[(P + 1) % 10]

    W = [0.0] * 10
    xm10 = 0.0
    for i in range(11):
        P = (19 - i) % 10
        j = m + 10 - i
        v = row_solve(j, W, P)
        if i == 0:
            xm10 = v
        W[P] = v
        x[j] = v
    Wb = [0.0] * 10
    Wb[9] = xm10
    for k in range(9):
        Wb[k] = W[8 - k]
    Wh = [list(W), Wb]
    for h in range(2):
        cnt = T_bot if h else m
        row = (m + 11) if h else (m - 1)
        rinc = 1 if h else -1
        i0 = 0
        while i0 < m:
            for ii in range(10):
                P = (18 - ii) % 10
                valid = i0 + ii < cnt
                v = row_solve(row if valid else 0, Wh[h], P)
                Wh[h][P] = v
                if valid:
                    x[row] = v
                row += rinc
            i0 += 10
    return x


if __name__ == "__main__":
    worst = 0.0
    for N in (12, 13, 16, 20, 23, 24, 32, 33, 44, 48, 100, 101, 400, 800):
        for seed in range(3):
            A, b, Hs = make_system(N, seed)
            x = solve_lat(Hs, N)
            ref = np.linalg.solve(A, b)
            err = np.abs(x - ref).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err < 1e-11, (N, seed, err)
    print("twisted model OK, worst relative error", worst)
