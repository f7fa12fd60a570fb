"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py): the oracle must still produce
them (CPU), and the CUDA path must reproduce them bit for bit through the C ABI (GPU)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as orc
from test_oracle import BG, OM

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES  # noqa: E402


def load(name):
    return np.load(os.path.join(HERE, "golden", name))


def test_fixtures_are_committed():
    assert len(glob.glob(os.path.join(HERE, "golden", "outlier_*.npz"))) == len(CASES)
    assert os.path.exists(os.path.join(HERE, "golden", "simple.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_golden(case):
    name, n, h, w, c, thr, bg, om, wts, fade, idx, spos = case
    g = load(f"outlier_{name}.npz")
    f = orc.fade(*fade) if fade else None
    img, msk, warn, dbg = orc.outlier(g["stack"], orc.threshold(*thr), BG[bg], OM[om], wts, f, idx, spos, seed=77, want_debug=True)
    assert np.array_equal(img, g["image"]) and np.array_equal(msk, g["mask"]) and warn == int(g["warnings"])
    assert np.array_equal(dbg["median"], g["median"]) and np.array_equal(dbg["n_outliers"], g["n_outliers"])


def test_oracle_reproduces_golden_simple():
    g = load("simple.npz")
    st = g["stack"]
    assert np.array_equal(orc.simple(st, True), g["darker"]) and np.array_equal(orc.simple(st, False), g["lighter"])
    assert np.array_equal(orc.simple(st, True, weights=(1, 0.5, 0.25, 0)), g["darker_weighted"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_reproduces_golden(case):
    import chrono_photo_b200 as cp
    name, n, h, w, c, thr, bg, om, wts, fade, idx, spos = case
    g = load(f"outlier_{name}.npz")
    st = g["stack"]
    ctx = cp.Context()
    fs = cp.FrameStack(ctx, w, h, c, n)
    fs.upload_all(st)
    sample = None
    seed = 77
    if spos is not None:
        # the fixture fixes the subset explicitly; the library draws its own, so find a seed-independent route:
        # run with the library's subset for the same size and compare against the oracle on that subset instead
        sample = len(spos)
    proc = cp.OutlierProcessor(cp.Threshold(*thr), BG[bg], OM[om], wts, cp.Fade(*fade) if fade else None, None, sample, seed=seed)
    img, msk, dbg = proc.process(fs, idx, debug=True)
    if spos is None:
        assert np.array_equal(img, g["image"]) and np.array_equal(msk, g["mask"]) and proc.warnings == int(g["warnings"])
        assert np.array_equal(dbg["median"], g["median"]) and np.array_equal(dbg["n_outliers"], g["n_outliers"])
        if not thr[0]:
            assert np.array_equal(dbg["q1"], g["q1"]) and np.array_equal(dbg["q3"], g["q3"])
    else:
        nwin = len(idx) if idx is not None else n
        lib_pos = cp.sample_positions(seed, nwin, sample)
        oimg, omsk, owarn = orc.outlier(st, orc.threshold(*thr), BG[bg], OM[om], wts, orc.fade(*fade) if fade else None, idx, lib_pos, seed=seed)
        assert np.array_equal(img, oimg) and np.array_equal(msk, omsk)
    fs.close()
    ctx.close()


@pytest.mark.gpu
def test_cuda_reproduces_golden_simple():
    import chrono_photo_b200 as cp
    g = load("simple.npz")
    st = g["stack"]
    n, h, w, c = st.shape
    ctx = cp.Context()
    fs = cp.FrameStack(ctx, w, h, c, n)
    fs.upload_all(st)
    assert np.array_equal(cp.SimpleProcessor(darker=True).process(fs), g["darker"])
    assert np.array_equal(cp.SimpleProcessor(darker=False).process(fs), g["lighter"])
    assert np.array_equal(cp.SimpleProcessor((1, 0.5, 0.25, 0), None, True).process(fs), g["darker_weighted"])
    fade = cp.Fade(0, False, [(0, 1.0), (8, 0.0)])
    assert np.array_equal(cp.SimpleProcessor((1, 1, 1, 1), fade, False).process(fs, list(range(2, 20, 2))), g["lighter_fade_window"])
    fs.close()
    ctx.close()
