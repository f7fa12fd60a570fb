"""CPU restatement of the iterative tier's window iteration (band_solve_windows, chrono_photo_b200/csrc/chb_kernels.cuh): the
placement of the five-value windows, the separate resolution of the two ranks of a median pair and the bracket bookkeeping,
checked for exactness and termination on random and degenerate series. The CUDA code is covered bit for bit by the GPU parity
tests; this test pins the LOGIC (a stall of the first version -- one bracket for two ranks whose values lie apart -- was found
with exactly this restatement) and documents the window counts quoted in DESIGN.md."""
import numpy as np
import pytest


def rint(x):
    return int(np.rint(x))


def solve(x, cap, max_windows=64):
    """Returns (mlo, mhi, windows used). x: the window's samples (uint8 values); cap: bytes the lane holds (zeros pad the rest)."""
    n = len(x)
    pad = cap - n
    kp1, kp2 = (n - 1) // 2 + pad, n // 2 + pad
    xs = np.concatenate([np.asarray(x, dtype=np.int64), np.zeros(pad, dtype=np.int64)])
    F = lambda c: int(np.abs(xs - c).sum())
    cnt = lambda c: int((xs <= c).sum())
    bsum, inv = int(np.sum(x)), np.float32(1.0) / np.float32(n)
    g = rint(np.float32(bsum) * inv)
    kt, xl, cl, xh, ch = kp1, -1, 0, 255, cap
    have1 = have2 = False
    mlo = mhi = 0
    for it in range(max_windows):
        p = min(max(g - 2, 0), 251)
        cn = [cnt(p + k) for k in range(4)]
        lo_end, hi_end = p == 0, p == 251
        if not have1 and (cn[0] <= kp1 or lo_end) and (kp1 < cn[3] or hi_end):
            mlo, have1 = p + sum(c <= kp1 for c in cn), True
            if kp2 != kp1 and not have2:
                kt, xh, ch = kp2, 255, cap
        if not have2 and (cn[0] <= kp2 or lo_end) and (kp2 < cn[3] or hi_end):
            mhi, have2 = p + sum(c <= kp2 for c in cn), True
        if have1 and have2:
            return mlo, mhi, it + 1
        for k in range(4):
            if cn[k] <= kt and p + k > xl:
                xl, cl = p + k, cn[k]
        for k in range(3, -1, -1):
            if cn[k] > kt and p + k < xh:
                xh, ch = p + k, cn[k]
        if it == 0:
            gc, nj, fm = p + 2, cn[2], F(p + 2)
            up = kp1 >= nj
            if up:
                sum_up = (fm + 2 * nj - cap + bsum - cap * (gc + 1)) >> 1
                side, far, fside = gc + 1 + rint(sum_up / max(cap - nj, 1)), kp1 + 1 - nj, (fm + 2 * nj - cap) - pad * (gc + 1)
            else:
                sum_dn = ((fm - bsum + cap * gc) >> 1) - pad * gc
                side, far, fside = gc - rint(sum_dn / max(nj - pad, 1)), nj - kp1, fm - pad * gc
            step = rint(far * 3.5 * fside * float(inv) * float(inv))
            dside = side - gc if up else gc - side
            g2 = (gc + step if up else gc - step) if 4 * step < dside else side
        elif it >= 3 and (it & 1):
            g2 = (xl + xh + 1) >> 1
        else:
            g2 = xl + rint((kt + 1 - cl) * (xh - xl) / max(ch - cl, 1))
        g = min(max(g2, xl + 2), max(xl + 2, xh - 1))
    raise AssertionError("window iteration did not terminate")


def check(x, cap):
    s = np.sort(np.asarray(x))
    n = len(x)
    mlo, mhi, w = solve(x, cap)
    assert (mlo, mhi) == (s[(n - 1) // 2], s[n // 2]), (list(x)[:16], mlo, mhi)
    return w


@pytest.mark.parametrize("n,cap", [(200, 208), (199, 208), (25, 32), (7, 16), (1, 16), (2, 16), (256, 256)])
def test_window_iteration_is_exact_on_iid_bytes(n, cap):
    rng = np.random.default_rng(n)
    for _ in range(300):
        check(rng.integers(0, 256, n), cap)


def test_window_iteration_on_two_clusters_and_degenerate_series():
    rng = np.random.default_rng(5)
    for _ in range(300):  # an object resting on the pixel for 10 - 40 % of the series
        n = 200
        bg = int(rng.integers(40, 200))
        x = bg + rng.integers(-5, 6, n)
        idx = rng.choice(n, int(rng.integers(n // 10, n * 4 // 10)), replace=False)
        x[idx] = np.clip(bg + rng.choice([-1, 1]) * rng.integers(30, 100) + rng.integers(-5, 6, len(idx)), 0, 255)
        assert check(x, 208) <= 8
    for x in (np.zeros(200, int), np.full(200, 255), np.array([0, 255] * 100), np.array([3] * 100 + [200] * 100),
              np.array([0] * 101 + [255] * 99), np.array([255] * 101 + [0] * 99), np.arange(200), np.arange(200)[::-1] + 56,
              np.array([7] * 100 + [8] * 100), np.array([0] * 100 + [1] * 100), np.array([254] * 100 + [255] * 100)):
        check(x, 208)  # the two ranks may lie far apart; every series terminates (bisection bounds the walk)


def test_window_counts_quoted_in_design():
    """iid bytes: about four windows for the slowest of 32 lanes (the kernel's cap is eight, then band_solve takes over)."""
    rng = np.random.default_rng(1)
    w = np.array([check(rng.integers(0, 256, 200), 208) for _ in range(3200)]).reshape(100, 32).max(axis=1)
    assert 3.5 < w.mean() < 5.5 and (w <= 8).mean() > 0.95
