"""CPU tests (no GPU): the oracle against the reference's own known-answer vectors, against an
independent numpy restatement, and against properties the algorithm must have."""
import numpy as np
import pytest

import np_restatement as npr
import oracle_lib as orc

BG = {"first": 0, "random": 1, "average": 2, "median": 3}
OM = {"first": 0, "last": 1, "extreme": 2, "average": 3, "forward": 4, "backward": 5}


def make_stack(rng, n, h, w, c, noise=5, n_obj=3):
    base = rng.integers(40, 200, size=(1, h, w, c))
    st = base + rng.integers(-noise, noise + 1, size=(n, h, w, c))
    for _ in range(n_obj):
        f = rng.integers(0, n, size=max(1, n // 6))
        y, x = rng.integers(0, h), rng.integers(0, w)
        st[f, y:y + 2, x:x + 3, :] = rng.integers(0, 256, size=(1, 1, 1, c))
    return np.clip(st, 0, 255).astype(np.uint8)


# ---- the reference's own vectors -------------------------------------------------------------------------------
def test_quartiles_known_answer():
    # src/chrono.rs:598-604
    assert orc.quartiles([0, 1, 2, 3, 4, 5, 6]) == (1.0, 3.0, 5.0)


def test_blend_known_answers():
    # src/color.rs:55-63 (commented-out blend_test documents the intent)
    assert list(orc.blend_into_u8([0] * 4, [255] * 4, 0.0)) == [0] * 4
    assert list(orc.blend_into_u8([0] * 4, [255] * 4, 0.5)) == [128] * 4
    assert list(orc.blend_into_u8([0] * 4, [255] * 4, 1.0)) == [255] * 4


def test_threshold_defaults():
    # abs/0.05/0.2 in f32
    t = orc.threshold(True, 0.05, 0.2)
    assert t.min == np.float32(0.05) * np.float32(255) and t.max == np.float32(0.2) * np.float32(255)
    assert abs(t.min - 12.75) < 1e-5 and abs(t.max - 51.0) < 1e-5
    assert np.float32(t.scale) == np.float32(1.0) / ((np.float32(0.2) - np.float32(0.05)) * np.float32(255))
    single = orc.threshold(True, 0.1, 0.1)  # single-value form: max == min -> scale = inf, never used
    assert np.isinf(single.scale)


def test_fade_test_vector():
    # src/options.rs:350-356 "clamp/abs/0,0/10,1"
    f, keep = orc.fade(0, True, [(0, 0.0), (10, 1.0)])
    lib = orc.lib()
    assert f.n_values == 11 and f.offset == 0
    assert lib.orc_fade_get(f, -5) == 0.0 and lib.orc_fade_get(f, 50) == 1.0
    assert abs(lib.orc_fade_get(f, 5) - 0.5) < 1e-6
    f2, keep2 = orc.fade(1, True, [(0, 0.0), (4, 1.0)])  # repeat
    assert lib.orc_fade_get(f2, 5) == lib.orc_fade_get(f2, 0) and lib.orc_fade_get(f2, -1) == lib.orc_fade_get(f2, 4)


# ---- order statistics -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 25, 200, 255, 256, 257, 1000])
def test_median_and_quartiles_match_numpy_restatement(n):
    rng = np.random.default_rng(n)
    for _ in range(20):
        d = np.sort(rng.integers(0, 256, size=n).astype(np.uint8))
        q1, m, q3 = orc.quartiles(d)
        assert m == npr.median_sorted(list(d)) and q1 == npr.quantile_sorted(list(d), 0.25) and q3 == npr.quantile_sorted(list(d), 0.75)
        # results are exact multiples of .5 / .25
        assert (m * 2) % 1 == 0 and (q1 * 4) % 1 == 0 and (q3 * 4) % 1 == 0


def test_quantile_positions():
    # N=200 -> Q1 = .75 d[49] + .25 d[50]; N=25 -> .5 (d[5] + d[6])
    d = np.arange(200, dtype=np.uint8)
    assert orc.quartiles(d)[0] == 0.75 * 49 + 0.25 * 50 and orc.quartiles(d)[2] == 0.25 * 149 + 0.75 * 150
    d = np.arange(25, dtype=np.uint8)
    assert orc.quartiles(d)[0] == 0.5 * (5 + 6) and orc.quartiles(d)[2] == 0.5 * (18 + 19)


# ---- full pixel path against the independent restatement ----------------------------------------------------------
@pytest.mark.parametrize("bg", ["first", "random", "average", "median"])
@pytest.mark.parametrize("om", ["first", "last", "extreme", "average", "forward", "backward"])
def test_outlier_abs_matches_numpy_restatement(bg, om):
    rng = np.random.default_rng(hash((bg, om)) % 2**32)
    st = make_stack(rng, 12, 5, 6, 3)
    thr = orc.threshold(True, 0.05, 0.2)
    img, msk, warn = orc.outlier(st, thr, BG[bg], OM[om], seed=7)
    img2, msk2, warn2 = npr.outlier(st, True, thr.min, thr.max, thr.scale, BG[bg], OM[om], seed=7)
    assert np.array_equal(img, img2) and np.array_equal(msk, msk2) and warn == warn2


@pytest.mark.parametrize("om", ["extreme", "average", "forward", "backward"])
@pytest.mark.parametrize("n,c", [(9, 3), (10, 4)])
def test_outlier_rel_matches_numpy_restatement(om, n, c):
    rng = np.random.default_rng(n * 10 + c)
    st = make_stack(rng, n, 4, 5, c)
    thr = orc.threshold(False, 3.0, 5.0)
    for bg in ("first", "average"):
        img, msk, warn = orc.outlier(st, thr, BG[bg], OM[om])
        img2, msk2, warn2 = npr.outlier(st, False, thr.min, thr.max, thr.scale, BG[bg], OM[om])
        assert np.array_equal(img, img2) and np.array_equal(msk, msk2) and warn == warn2


def test_outlier_window_fade_weights_sample_matches_numpy_restatement():
    rng = np.random.default_rng(5)
    st = make_stack(rng, 20, 4, 4, 3)
    thr = orc.threshold(True, 0.04, 0.15)
    idx = list(range(3, 17, 2))
    spos = [0, 2, 3, 5, 6]
    for absolute in (True, False):
        fo = orc.fade(1, absolute, [(0, 0.0), (4, 1.0), (6, 0.25)])
        fn = npr.Fade.build(1, absolute, [(0, 0.0), (4, 1.0), (6, 0.25)])
        for w in ((1, 1, 1, 0), (0, 0, 1, 0), (1, 0.5, 0.5, 0)):  # docs/options.md:157-158
            img, msk, warn = orc.outlier(st, thr, BG["first"], OM["forward"], weights=w, fade_=fo, indices=idx, sample_pos=spos)
            img2, msk2, warn2 = npr.outlier(st, True, thr.min, thr.max, thr.scale, BG["first"], OM["forward"], weights=w, fade=fn,
                                            indices=idx, sample_pos=spos)
            assert np.array_equal(img, img2) and np.array_equal(msk, msk2) and warn == warn2


@pytest.mark.parametrize("seed", range(96))
def test_random_option_fuzz_oracle_against_numpy_restatement(seed):
    # The C oracle is the pin of GPU parity; its own pin beyond the reference's few known answers is this independent
    # restatement written from the Rust source. Seeded walk through the option space on small stacks (the restatement is a
    # pure-Python loop): frame counts incl. 1-2, RGB / RGBA, four data regimes, both threshold kinds, every policy pair,
    # zero / fractional / negative weights, fades of both kinds and modes, contiguous and stepped windows, --sample.
    rng = np.random.default_rng(50_000 + seed)
    n = int(rng.choice([1, 2, 3, 4, 5, 8, 13, 16, 17, 24]))
    c = int(rng.choice([3, 4]))
    h, w = int(rng.integers(2, 5)), int(rng.integers(2, 6))
    regime = int(rng.integers(0, 4))
    if regime == 0:
        st = make_stack(rng, n, h, w, c, noise=int(rng.integers(0, 12)), n_obj=int(rng.integers(0, 6)))
    elif regime == 1:
        st = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    elif regime == 2:
        st = rng.choice(np.array([0, 255, 3, 252], np.uint8), size=(n, h, w, c))
    else:
        base = rng.integers(0, 256, size=(1, h, w, c))
        st = np.clip(base + np.rint(rng.normal(0, rng.uniform(0.5, 9), size=(n, h, w, c))), 0, 255).astype(np.uint8)
    idx = None
    if n >= 4 and rng.integers(0, 2):
        a0 = int(rng.integers(0, n // 2))
        idx = list(range(a0, int(rng.integers(a0 + 1, n + 1)), int(rng.choice([1, 1, 2, 3]))))
    nwin = len(idx) if idx is not None else n
    absolute = bool(rng.integers(0, 2)) or nwin < 3  # quantile() needs three samples (src/chrono.rs:569-570)
    if absolute:
        mn = float(rng.choice([0.0, 0.01, 0.05, 0.1, 0.3]))
        thr = orc.threshold(True, mn, mn + float(rng.choice([0.0, 0.05, 0.15, 0.5])))
    else:
        mn = float(rng.choice([0.5, 1.0, 3.0, 6.0]))
        thr = orc.threshold(False, mn, mn + float(rng.choice([0.0, 1.0, 2.0])))
    weights = tuple(float(x) for x in rng.choice([1.0, 1.0, 1.0, 0.0, 0.5, 2.0, -1.0], size=4))
    if not any(weights[:c]):
        weights = (1.0,) + weights[1:]
    fo = fn = None
    if rng.integers(0, 3) == 0:
        mode, fabs = int(rng.integers(0, 2)), bool(rng.integers(0, 2))
        f0 = int(rng.integers(-3, 4))
        pts = [(f0, float(rng.uniform(-0.5, 1.5))), (f0 + int(rng.integers(1, 6)), float(rng.uniform(-0.5, 1.5)))]
        if rng.integers(0, 2):
            pts.append((pts[-1][0] + int(rng.integers(1, 5)), float(rng.uniform(0.0, 1.0))))
        fo, fn = orc.fade(mode, fabs, pts), npr.Fade.build(mode, fabs, pts)
    spos = None
    if nwin >= 4 and rng.integers(0, 4) == 0:
        k = int(rng.integers(3 if not absolute else 1, nwin))
        spos = sorted(int(v) for v in rng.choice(nwin, size=k, replace=False))
    bg, om = int(rng.integers(0, 4)), int(rng.integers(0, 6))
    seed_rng = int(rng.integers(0, 1000))
    img, msk, warn = orc.outlier(st, thr, bg, om, weights=weights, fade_=fo, indices=idx, sample_pos=spos, seed=seed_rng, pixel_offset=11)
    img2, msk2, warn2 = npr.outlier(st, absolute, thr.min, thr.max, thr.scale, bg, om, weights=weights, fade=fn, indices=idx,
                                    sample_pos=spos, seed=seed_rng, pixel_offset=11)
    tag = f"n={n} c={c} regime={regime} abs={absolute} bg={bg} om={om} w={weights} idx={idx} spos={spos}"
    assert np.array_equal(msk, msk2), "mask " + tag
    assert np.array_equal(img, img2), "composite " + tag
    assert warn == warn2, "warnings " + tag


def test_all_outlier_warning_path():
    # threshold 0: every frame is an outlier (dist_sq >= 0), first_excluded returns (0, warning) src/chrono.rs:510-516
    rng = np.random.default_rng(11)
    st = make_stack(rng, 6, 3, 3, 3)
    thr = orc.threshold(True, 0.0, 0.2)
    img, msk, warn = orc.outlier(st, thr, BG["first"], OM["extreme"])
    assert warn == 9
    img2, msk2, warn2 = npr.outlier(st, True, thr.min, thr.max, thr.scale, BG["first"], OM["extreme"])
    assert np.array_equal(img, img2) and np.array_equal(msk, msk2) and warn2 == 9


def test_constant_frames_no_outliers_and_iqr_zero_branch():
    st = np.full((8, 3, 3, 3), 77, np.uint8)  # IQR == 0 -> iqr_inv = 1 (src/chrono.rs:249-251)
    for absolute, mn, mx in ((True, 0.05, 0.2), (False, 3.0, 5.0)):
        img, msk, warn = orc.outlier(st, orc.threshold(absolute, mn, mx), BG["median"], OM["extreme"])
        assert (img == 77).all() and (msk == 0).all() and warn == 0


def test_mask_alpha_band_is_255():
    rng = np.random.default_rng(3)
    st = make_stack(rng, 9, 3, 3, 4)
    _, msk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG["first"], OM["extreme"])
    assert (msk[..., 3] == 255).all()  # src/chrono.rs:186-190


def test_sample_excluded_quirk_is_reproduced():
    # n=10, outliers {3, 9}: after the position swaps the candidate range holds frame 9 and lacks frame 8
    st = np.full((10, 1, 1, 3), 100, np.uint8)
    st[3] = 0
    st[9] = 255
    thr = orc.threshold(True, 0.05, 0.2)
    seen = set()
    for seed in range(400):
        _, _, _, d = orc.outlier(st, thr, BG["random"], OM["first"], seed=seed, want_debug=True)
        seen.add(int(d["bg_index"][0]))
    assert seen == {0, 1, 2, 9, 4, 5, 6, 7}


def test_simple_matches_numpy_restatement():
    rng = np.random.default_rng(2)
    st = make_stack(rng, 11, 4, 5, 3)
    for darker in (True, False):
        assert np.array_equal(orc.simple(st, darker), npr.simple(st, darker))
        fo = orc.fade(0, False, [(0, 1.0), (5, 0.0)])
        fn = npr.Fade.build(0, False, [(0, 1.0), (5, 0.0)])
        assert np.array_equal(orc.simple(st, darker, weights=(1, 0.5, 0.25, 0), fade_=fo, indices=[1, 2, 5, 6, 9]),
                              npr.simple(st, darker, weights=(1, 0.5, 0.25, 0), fade=fn, indices=[1, 2, 5, 6, 9]))


@pytest.mark.parametrize("seed", range(32))
def test_simple_fuzz_oracle_against_numpy_restatement(seed):
    # darker / lighter: ties (few distinct values), RGB / RGBA incl. the alpha band in the weighted sum, fractional / zero /
    # negative weights, fades (running order-dependent blend), windows
    rng = np.random.default_rng(70_000 + seed)
    n, c = int(rng.choice([1, 2, 3, 7, 16, 17, 30])), int(rng.choice([3, 4]))
    h, w = int(rng.integers(1, 5)), int(rng.integers(1, 7))
    if rng.integers(0, 2):
        st = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    else:
        st = rng.choice(np.array([0, 1, 2, 128, 254, 255], np.uint8), size=(n, h, w, c))
    weights = tuple(float(x) for x in rng.choice([1.0, 1.0, 0.0, 0.5, 0.25, 2.0, -1.0], size=4))
    idx = None
    if n >= 3 and rng.integers(0, 2):
        a0 = int(rng.integers(0, n - 1))
        idx = list(range(a0, int(rng.integers(a0 + 1, n + 1)), int(rng.choice([1, 2, 3]))))
    fo = fn = None
    if rng.integers(0, 2):
        mode, fabs = int(rng.integers(0, 2)), bool(rng.integers(0, 2))
        f0 = int(rng.integers(-2, 3))
        pts = [(f0, float(rng.uniform(-0.5, 1.5))), (f0 + int(rng.integers(1, 8)), float(rng.uniform(-0.5, 1.5)))]
        fo, fn = orc.fade(mode, fabs, pts), npr.Fade.build(mode, fabs, pts)
    for darker in (True, False):
        got = orc.simple(st, darker, weights=weights, fade_=fo, indices=idx)
        want = npr.simple(st, darker, weights=weights, fade=fn, indices=idx)
        assert np.array_equal(got, want), f"darker={darker} n={n} c={c} w={weights} idx={idx} fade={fo is not None}"


def test_simple_first_frame_wins_ties():
    st = np.zeros((3, 1, 2, 3), np.uint8)
    st[0, 0, 0] = (10, 20, 30)
    st[1, 0, 0] = (30, 20, 10)  # same sum: strict compare keeps frame 0 (src/simple.rs:108-118)
    st[2, 0, 0] = (20, 20, 20)
    assert tuple(orc.simple(st, True)[0, 0]) == (10, 20, 30) and tuple(orc.simple(st, False)[0, 0]) == (10, 20, 30)


def test_multithreaded_oracle_is_identical():
    rng = np.random.default_rng(8)
    st = make_stack(rng, 15, 16, 9, 3)
    thr = orc.threshold(True, 0.05, 0.2)
    a = orc.outlier(st, thr, BG["random"], OM["extreme"], seed=3, n_threads=1)
    b = orc.outlier(st, thr, BG["random"], OM["extreme"], seed=3, n_threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert np.array_equal(orc.simple(st, True, n_threads=1), orc.simple(st, True, n_threads=3))


# ---- host-side callers --------------------------------------------------------------------------------------------
def test_crop_create():
    # src/shake.rs:136-176
    assert orc.crop_create([(0, 0), (0, 0)], 100, 80) is None
    xy, w, h = orc.crop_create([(0, 0), (3, -2), (-1, 4)], 100, 80)
    assert (w, h) == (100 - 1 - 3, 80 - 2 - 4)
    assert xy.tolist() == [[1, 2], [4, 0], [0, 6]]


def test_video_windows_quirks():
    # --video-in 0/25/1 on 1800 images: v_lower = -24, 1824 output frames, last image never enters
    n, ws, we, num = orc.video_windows(1800, (0, 25, 1), (None, None, 1))
    assert n == 1824 and num[0] == 0 and (ws[0], we[0]) == (0, 1) and (ws[24], we[24]) == (0, 25)
    assert we.max() == 1799  # end capped at image_count - step (src/main.rs:272-276)
    # open start: growing trail from 0 (src/main.rs:270)
    n, ws, we, num = orc.video_windows(10, (None, 3, 1), (None, None, 1))
    assert (ws == 0).all() and we.tolist() == [min(9, f + 3) for f in range(10)]
    # step 2 windows start on the step grid
    n, ws, we, num = orc.video_windows(20, (-4, 2, 2), (0, 20, 1))
    for f in range(20):
        st = f - 4
        while st < 0:
            st += 2
        assert ws[f] == max(st % 2, f - 4)


# ---- shake analysis (src/shake.rs:190-386) -----------------------------------------------------------------------
def shaky_clip(rng, n, H, W, C, max_shift, margin):
    """A textured scene seen through a camera that shifts by a known integer offset per frame (frame 0 unshifted)."""
    scene = rng.integers(0, 256, size=(H + 2 * margin, W + 2 * margin, C), dtype=np.uint8)
    offs = rng.integers(-max_shift, max_shift + 1, size=(n, 2))
    offs[0] = 0
    frames = np.stack([scene[margin - oy:margin - oy + H, margin - ox:margin - ox + W] for ox, oy in offs])
    return frames, [tuple(int(v) for v in o) for o in offs]


def np_shake_offsets(frames, anchors, r, s):
    """Independent numpy restatement of calc_diffs + first minimum."""
    out, tables = [(0, 0)], []
    f0 = frames[0].astype(np.int64)
    for fr in frames[1:]:
        fr = fr.astype(np.int64)
        diffs = np.zeros((2 * s + 1, 2 * s + 1), np.int64)
        for cx, cy in anchors:
            win = f0[cy - r:cy + r + 1, cx - r:cx + r + 1]
            for oy in range(-s, s + 1):
                for ox in range(-s, s + 1):
                    pat = fr[cy + oy - r:cy + oy + r + 1, cx + ox - r:cx + ox + r + 1]
                    diffs[oy + s, ox + s] += ((win - pat) ** 2).sum()
        d32 = ((diffs + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)  # i32 wrap-around
        idx = int(np.argmin(d32.ravel()))  # argmin returns the first minimum
        out.append((idx % (2 * s + 1) - s, idx // (2 * s + 1) - s))
        tables.append(d32.ravel())
    return out, tables


def test_shake_offsets_recover_known_shifts_and_match_numpy():
    rng = np.random.default_rng(321)
    frames, true = shaky_clip(rng, 7, 60, 80, 3, 4, 6)
    anchors = [(30, 25), (55, 40)]
    got, tables = orc.shake_analyze(frames, anchors, 6, 5, want_diffs=True)
    ref, ref_tables = np_shake_offsets(frames, anchors, 6, 5)
    assert got == ref and all(np.array_equal(a, b) for a, b in zip(tables, ref_tables))
    # a frame shifted by (ox, oy) shows the anchor's content at anchor + (ox, oy): the analysis returns exactly that
    assert got == true
    # Crop::create then realigns the frames (src/shake.rs:136-176)
    xy, w, h = orc.crop_create(got, 80, 60)
    aligned = np.stack([frames[i, xy[i][1]:xy[i][1] + h, xy[i][0]:xy[i][0] + w] for i in range(len(frames))])
    assert all(np.array_equal(aligned[0], a) for a in aligned[1:])


def test_shake_first_minimum_wins_and_range_panics():
    flat = np.full((3, 40, 40, 3), 77, np.uint8)  # every offset has the same (zero) difference: the first one wins
    assert orc.shake_analyze(flat, [(20, 20)], 3, 2) == [(0, 0), (-2, -2), (-2, -2)]
    with pytest.raises(IndexError):
        orc.shake_analyze(flat, [(2, 20)], 3, 2)        # window leaves the image: the reference panics
    with pytest.raises(IndexError):
        orc.shake_analyze(flat, [(4, 20)], 3, 2)        # search square leaves the image


def test_shake_diffs_wrap_like_i32():
    # 3 anchors x 129^2 x 3 bytes of maximal difference overflow i32; the release build of the reference wraps
    a = np.zeros((2, 140, 140, 3), np.uint8)
    a[1] = 255
    anchors = [(69, 69)] * 3
    got, tables = orc.shake_analyze(a, anchors, 64, 0, want_diffs=True)
    exact = 3 * 129 * 129 * 3 * 255 * 255
    assert exact > 2 ** 31 and int(tables[0][0]) == ((exact + 2 ** 31) % 2 ** 32) - 2 ** 31 and got == [(0, 0), (0, 0)]
