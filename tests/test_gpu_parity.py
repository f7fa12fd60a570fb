"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs. Bit-exact everywhere (composite, mask, medians, quartiles, outlier counts, darker/lighter), which is
stricter than the +-1 LSB north_star allows for blended bytes."""
import os

import numpy as np
import pytest

import oracle_lib as orc
from test_oracle import BG, OM, make_stack

import chrono_photo_b200 as cp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cp.Context()
    yield c
    c.close()


def upload(ctx, st):
    n, h, w, c = st.shape
    fs = cp.FrameStack(ctx, w, h, c, n)
    fs.upload_all(st)
    return fs


def thr_pair(spec):
    absolute, mn, mx = spec
    return cp.Threshold(absolute, mn, mx), orc.threshold(absolute, mn, mx)


def check_outlier(ctx, st, spec, bg, om, weights=(1, 1, 1, 1), fade=None, indices=None, sample=None, seed=5, fs=None, debug=True):
    """Runs both sides and asserts bit equality. fade: (mode, absolute, frames) or None."""
    t_gpu, t_orc = thr_pair(spec)
    own = fs is None
    if own:
        fs = upload(ctx, st)
    f_gpu = cp.Fade(*fade) if fade else None
    f_orc = orc.fade(*fade) if fade else None
    n = len(indices) if indices is not None else st.shape[0]
    spos = None
    if sample is not None and sample < n:
        spos = cp.sample_positions(seed, n, sample)
    proc = cp.OutlierProcessor(t_gpu, BG[bg], OM[om], weights, f_gpu, None, sample, seed=seed)
    img, msk, dbg = proc.process(fs, indices, debug=True)
    oimg, omsk, owarn, odbg = orc.outlier(st, t_orc, BG[bg], OM[om], weights, f_orc, indices, spos, seed=seed, want_debug=True)
    tag = f"{spec} bg={bg} om={om} w={weights} n={n}"
    assert np.array_equal(dbg["median"], odbg["median"]), "median " + tag
    if not spec[0]:
        assert np.array_equal(dbg["q1"], odbg["q1"]) and np.array_equal(dbg["q3"], odbg["q3"]), "quartiles " + tag
    # the kernel only counts outliers for pixels that leave the certified fast path; fast-path pixels have none
    assert np.array_equal(dbg["n_outliers"], odbg["n_outliers"]), "outlier count " + tag
    assert np.array_equal(msk, omsk), "mask " + tag
    assert np.array_equal(img, oimg), "composite " + tag
    assert proc.warnings == owarn, "warnings " + tag
    # the production call (no debug planes) may take shortcuts the debug call does not (IQR bound on the fast tier)
    img2, msk2 = proc.process(fs, indices)
    assert np.array_equal(img2, oimg) and np.array_equal(msk2, omsk) and proc.warnings == owarn, "non-debug call " + tag
    if own:
        fs.close()
    return img, msk


# ---------------------------------------------------------------------------------------------------- ingest / generator
def test_upload_then_darker_lighter_roundtrip(ctx):
    rng = np.random.default_rng(0)
    st = make_stack(rng, 19, 37, 53, 3)
    fs = upload(ctx, st)
    for darker in (True, False):
        assert np.array_equal(cp.SimpleProcessor(darker=darker).process(fs), orc.simple(st, darker))
    for f in (0, 7, 15, 16, 18):
        assert np.array_equal(fs.download(f), st[f])
    fs.close()


def test_upload_with_pitch_and_crop(ctx):
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, size=(7, 40, 64, 3), dtype=np.uint8)
    offs = [(0, 0), (3, 1), (-2, 4), (1, -3), (0, 2), (2, 2), (-1, -1)]
    xy, w, h = cp.crop_create(offs, 64, 40)
    fs = cp.FrameStack(ctx, w, h, 3, 7)
    for i in range(7):
        fs.upload(i, big[i], tuple(xy[i]))
    fs.sync()
    st = np.stack([big[i, xy[i][1]:xy[i][1] + h, xy[i][0]:xy[i][0] + w] for i in range(7)])
    assert np.array_equal(cp.SimpleProcessor(darker=False).process(fs), orc.simple(st, False))
    fs.close()


def test_jpeg_ingest_and_encode_round_trip(ctx):
    # JPEG in (nvJPEG decode on the device into the stack, crop origin applied) and JPEG out (save_image's JPEG branch).
    # Codecs differ in the last bit, so parity is defined on DECODED frames: the stack's content after the JPEG upload is
    # what both the CUDA path and the oracle then composite.
    rng = np.random.default_rng(5)
    n, h, w = 20, 64, 96
    yy, xx = np.mgrid[0:h + 8, 0:w + 16]
    frames = []
    for f in range(n):
        base = np.stack([(xx * 2 + f) % 256, (yy * 3) % 256, (xx + yy) % 256], axis=2).astype(np.float32)
        img = np.clip(base + rng.normal(0, 3, base.shape), 0, 255).astype(np.uint8)
        if 5 <= f < 9:
            img[20:40, 30 + 4 * f:50 + 4 * f] = (250, 10, 10)
        frames.append(img)
    jpegs = [cp.encode_jpeg(ctx, fr, quality=95) for fr in frames]
    assert all(j[:2] == b"\xff\xd8" and j[-2:] == b"\xff\xd9" for j in jpegs) and len(jpegs[0]) < frames[0].nbytes
    fs = cp.FrameStack(ctx, w, h, 3, n)
    import threading
    crops = [(int(rng.integers(0, 16)), int(rng.integers(0, 8))) for _ in range(n)]
    errs = []

    def work(ids):
        try:
            for f in ids:
                fs.upload_jpeg(f, jpegs[f], crops[f])
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    threads = [threading.Thread(target=work, args=(range(k, n, 4),)) for k in range(4)]  # four decode threads, frames out of order
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    fs.sync()
    st = np.stack([fs.download(f) for f in range(n)])
    for f in range(n):  # lossy but close to the source crop
        src = frames[f][crops[f][1]:crops[f][1] + h, crops[f][0]:crops[f][0] + w].astype(np.int32)
        assert np.abs(st[f].astype(np.int32) - src).mean() < 4.0, f
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme", fs=fs)
    assert np.array_equal(cp.SimpleProcessor(darker=True).process(fs), orc.simple(st, True))
    with pytest.raises(cp._lib.ChbError):
        fs.upload_jpeg(0, b"not a jpeg at all")
    with pytest.raises(cp._lib.ChbError):
        fs.upload_jpeg(0, jpegs[0], (17, 0))  # crop window leaves the image
    fs.close()


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_device_generator_equals_host_twin(ctx, kind):
    n, h, w = 21, 48, 80
    fs = cp.FrameStack(ctx, w, h, 3, n)
    fs.fill_synthetic(kind, seed=42)
    st = np.stack([cp.synth_frame_host(kind, 42, f, n, w, h, 3) for f in range(n)])
    # darker and lighter together pin min/max per pixel; the outlier path with median background pins the medians
    assert np.array_equal(cp.SimpleProcessor(darker=True).process(fs), orc.simple(st, True))
    assert np.array_equal(cp.SimpleProcessor(darker=False).process(fs), orc.simple(st, False))
    check_outlier(ctx, st, (True, 0.05, 0.2), "median", "extreme", fs=fs)
    fs.close()


# ---------------------------------------------------------------------------------------------------- outlier policies
@pytest.mark.parametrize("bg", ["first", "random", "average", "median"])
@pytest.mark.parametrize("om", ["first", "last", "extreme", "average", "forward", "backward"])
def test_outlier_policies_abs(ctx, bg, om):
    rng = np.random.default_rng(100 + BG[bg] * 10 + OM[om])
    st = make_stack(rng, 25, 24, 40, 3, noise=5, n_obj=40)
    check_outlier(ctx, st, (True, 0.05, 0.2), bg, om)


@pytest.mark.parametrize("bg", ["first", "average"])
@pytest.mark.parametrize("om", ["extreme", "average", "forward", "backward"])
def test_outlier_policies_rel(ctx, bg, om):
    rng = np.random.default_rng(200 + BG[bg] * 10 + OM[om])
    st = make_stack(rng, 31, 20, 33, 3, noise=4, n_obj=30)
    check_outlier(ctx, st, (False, 3.0, 5.0), bg, om)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 25, 64, 65, 128, 200, 224, 255, 256, 257, 512, 1000, 1025])
def test_frame_counts_cover_every_kernel_variant(ctx, n):
    # SURVEY.md A5; small images so the oracle stays fast
    rng = np.random.default_rng(n)
    st = make_stack(rng, n, 6, 21, 3, noise=6, n_obj=8)
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
    if n >= 3:
        check_outlier(ctx, st, (False, 3.0, 5.0), "first", "forward")


@pytest.mark.parametrize("n", [2049, 4096])
def test_longest_series_one_launch_holds(ctx, n):
    # the register-resident variants end at (8 units, 32 lanes per pixel) = 4096 frames per window span; 2049 is the first
    # frame count that needs it, 4096 fills it
    rng = np.random.default_rng(n)
    st = make_stack(rng, n, 3, 9, 3, noise=6, n_obj=4)
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
    check_outlier(ctx, st, (False, 3.0, 5.0), "median", "backward")


@pytest.mark.parametrize("n", [4097, 5000])
def test_series_beyond_4096_frames_go_through_the_histogram_tier(ctx, n):
    # more frames than the register-resident variants hold: every pixel takes outlier_hist_kernel (any length) and, where its
    # certificate fails, the per-frame path -- slower, but the same composite, mask, statistics and warnings as the reference
    rng = np.random.default_rng(n)
    st = make_stack(rng, n, 3, 9, 3, noise=6, n_obj=4)
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
    check_outlier(ctx, st, (False, 3.0, 5.0), "median", "backward")
    check_outlier(ctx, st, (True, 0.05, 0.2), "average", "average")
    check_outlier(ctx, st, (True, 0.05, 0.2), "random", "forward")


def test_windows_of_a_series_beyond_4096_frames(ctx):
    st = np.zeros((4200, 2, 4, 3), np.uint8)
    fs = upload(ctx, st)
    # a window spanning more than 4096 frames is not a whole-stack launch: reported, not computed
    with pytest.raises(Exception) as ei:
        cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), BG["first"], OM["extreme"]).process(fs, list(range(0, 4199)))
    assert "whole-stack" in str(ei.value)
    # a window of the same stack that fits is fine, and so are darker / lighter (they stream, no capacity limit)
    img, _ = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), BG["first"], OM["extreme"]).process(fs, list(range(100, 400)))
    assert not img.any()
    assert not cp.SimpleProcessor(darker=True).process(fs).any()
    fs.close()


def test_rel_needs_three_samples(ctx):
    st = np.zeros((2, 4, 4, 3), np.uint8)
    fs = upload(ctx, st)
    with pytest.raises(Exception):  # quantile() underflow in the reference (src/chrono.rs:569-570)
        cp.OutlierProcessor(cp.Threshold.rel(3, 5), 0, 2).process(fs)
    fs.close()


def test_rgba_default_weights_include_alpha(ctx):
    rng = np.random.default_rng(7)
    st = make_stack(rng, 40, 16, 24, 4, n_obj=20)
    for bg in ("first", "median"):
        img, msk = check_outlier(ctx, st, (True, 0.05, 0.2), bg, "extreme")
        assert (msk[..., 3] == 255).all()
    check_outlier(ctx, st, (False, 2.0, 4.0), "average", "average")


@pytest.mark.parametrize("weights", [(1, 1, 1, 0), (0, 0, 1, 0), (1, 0.5, 0.5, 0), (2, 1, 0.25, 1), (-1, 1, 1, 0)])
def test_weights(ctx, weights):
    # docs/options.md:157-158 and the negative-weight corner of SURVEY.md's numerics checklist
    rng = np.random.default_rng(17)
    st = make_stack(rng, 33, 12, 20, 3, n_obj=15)
    for bg in ("first", "median", "average"):
        check_outlier(ctx, st, (True, 0.05, 0.2), bg, "extreme", weights=weights)
    check_outlier(ctx, st, (False, 3.0, 5.0), "first", "backward", weights=weights)


@pytest.mark.parametrize("fade", [(0, True, [(0, 0.0), (10, 1.0)]), (1, False, [(0, 1.0), (6, 0.0), (9, 0.5)]), (0, False, [(-2, 2.0), (30, -1.0)])])
def test_fade(ctx, fade):
    rng = np.random.default_rng(23)
    st = make_stack(rng, 28, 12, 20, 3, n_obj=25)
    for om in ("extreme", "forward", "backward", "average", "first"):
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om, fade=fade)
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om, fade=fade, indices=list(range(5, 25)))


@pytest.mark.parametrize("indices", [list(range(0, 40)), list(range(7, 33)), list(range(3, 60, 3)), [0, 1, 17, 18, 19, 40, 59], list(range(16, 32))])
def test_windows(ctx, indices):
    # image_indices of the video path (src/chrono.rs:102-139): contiguous, stepped and ragged ascending windows
    rng = np.random.default_rng(31)
    st = make_stack(rng, 60, 10, 24, 3, n_obj=30)
    fs = upload(ctx, st)
    for bg, om in (("first", "extreme"), ("median", "last"), ("average", "average"), ("random", "forward")):
        check_outlier(ctx, st, (True, 0.05, 0.2), bg, om, indices=indices, fs=fs)
    check_outlier(ctx, st, (False, 3.0, 5.0), "first", "backward", indices=indices, fs=fs)
    for darker in (True, False):
        assert np.array_equal(cp.SimpleProcessor(darker=darker).process(fs, indices), orc.simple(st, darker, indices=indices))
    fs.close()


@pytest.mark.parametrize("sample", [1, 2, 5, 24, 25, 100])
def test_sample_subset(ctx, sample):
    # --sample (src/chrono.rs:151-163): median / IQR on a subset, distances on every frame
    rng = np.random.default_rng(37)
    st = make_stack(rng, 25, 10, 16, 3, n_obj=15)
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme", sample=sample)
    if sample >= 3:
        check_outlier(ctx, st, (False, 3.0, 5.0), "median", "forward", sample=sample)
    check_outlier(ctx, st, (True, 0.05, 0.2), "average", "average", sample=sample, indices=list(range(2, 24, 2)) if sample <= 11 else None)


# ---------------------------------------------------------------------------------------------------- adversarial inputs
def test_iid_uniform_bytes(ctx):  # A1: worst case for the selection search
    rng = np.random.default_rng(41)
    for n in (25, 200, 300):
        st = rng.integers(0, 256, size=(n, 8, 32, 3), dtype=np.uint8)
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
        check_outlier(ctx, st, (False, 3.0, 5.0), "median", "average")


def test_constant_frames_iqr_zero(ctx):  # A2
    for v in (0, 1, 77, 254, 255):
        st = np.full((30, 5, 32, 3), v, np.uint8)
        img, msk = check_outlier(ctx, st, (False, 3.0, 5.0), "median", "extreme")
        assert (img == v).all() and (msk == 0).all()
        check_outlier(ctx, st, (True, 0.05, 0.2), "average", "extreme")


def test_two_alternating_values_even_n(ctx):  # A3: x.5 medians and interpolated quartiles
    st = np.zeros((200, 4, 32, 3), np.uint8)
    st[0::2] = 10
    st[1::2] = 21
    st[:, :, 16:, 1] += 1
    img, msk = check_outlier(ctx, st, (True, 0.01, 0.05), "median", "extreme")
    check_outlier(ctx, st, (False, 0.4, 0.6), "first", "forward")
    st[:, :, :, 2] = np.arange(200, dtype=np.uint8)[:, None, None]  # a ramp: every rank distinct
    check_outlier(ctx, st, (False, 1.0, 2.0), "average", "backward")


def test_extreme_values_and_edges_of_the_byte_range(ctx):
    rng = np.random.default_rng(43)
    st = rng.choice(np.array([0, 1, 254, 255], np.uint8), size=(64, 6, 32, 3))
    check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
    check_outlier(ctx, st, (False, 0.5, 1.0), "median", "last")
    st = np.zeros((33, 4, 32, 3), np.uint8)
    st[5] = 255
    check_outlier(ctx, st, (True, 0.05, 0.2), "average", "extreme")


def test_threshold_zero_every_frame_is_an_outlier(ctx):  # A4: warning path (src/chrono.rs:510-516)
    rng = np.random.default_rng(47)
    st = make_stack(rng, 12, 6, 20, 3)
    for bg in ("first", "random", "average"):
        for om in ("extreme", "forward", "average"):
            check_outlier(ctx, st, (True, 0.0, 0.2), bg, om)


def test_single_value_threshold(ctx):  # abs/0.1 -> max == min, scale = inf (src/options.rs:269-276)
    rng = np.random.default_rng(53)
    st = make_stack(rng, 25, 8, 20, 3, n_obj=20)
    check_outlier(ctx, st, (True, 0.1, 0.1), "first", "extreme")
    check_outlier(ctx, st, (True, 0.1, 0.1), "first", "forward")


def test_gaussian_like_noise_near_threshold(ctx):
    # noise whose tails cross the lower threshold: many pixels fail the certificate and take the exact path
    rng = np.random.default_rng(59)
    base = rng.integers(60, 190, size=(1, 16, 48, 3))
    st = np.clip(base + np.rint(rng.normal(0, 4.5, size=(50, 16, 48, 3))), 0, 255).astype(np.uint8)
    for om in ("extreme", "first", "last", "average", "forward", "backward"):
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om)
    check_outlier(ctx, st, (True, 0.05, 0.2), "random", "extreme")
    check_outlier(ctx, st, (True, 0.05, 0.2), "average", "forward")


# ---------------------------------------------------------------------------------------------------- simple mode details
def test_simple_weights_fade_and_ties(ctx):
    rng = np.random.default_rng(61)
    st = make_stack(rng, 37, 9, 33, 3, n_obj=20)
    st[5] = st[2]  # exact ties: the first frame must win (src/simple.rs:108-118)
    fs = upload(ctx, st)
    for darker in (True, False):
        for w in ((1, 1, 1, 1), (1, 0.5, 0.25, 0), (0, 0, 1, 0), (0.3, 0.59, 0.11, 0)):
            assert np.array_equal(cp.SimpleProcessor(w, None, darker).process(fs), orc.simple(st, darker, w)), (darker, w)
        for fade in ((0, False, [(0, 1.0), (10, 0.0)]), (1, True, [(0, 0.2), (5, 1.5), (8, -0.5)])):
            for idx in (None, list(range(4, 30, 2))):
                got = cp.SimpleProcessor((1, 1, 1, 1), cp.Fade(*fade), darker).process(fs, idx)
                assert np.array_equal(got, orc.simple(st, darker, fade_=orc.fade(*fade), indices=idx)), (darker, fade, idx)
    fs.close()
    st4 = make_stack(rng, 18, 7, 20, 4)
    fs = upload(ctx, st4)
    assert np.array_equal(cp.SimpleProcessor(darker=True).process(fs), orc.simple(st4, True))
    fs.close()


# ---------------------------------------------------------------------------------------------------- config 1 and sizes
def test_config1_minimal_example(ctx):
    # cmd_examples/minimal.chrono on the create-test-data recipe: 25 x 1024x768 RGB, defaults abs/0.05/0.2, extreme;
    # background random is the CLI default (src/cli.rs:194) -- bit-comparable here because oracle and kernel share the RNG
    n, h, w = 25, 768, 1024
    fs = cp.FrameStack(ctx, w, h, 3, n)
    fs.fill_synthetic(1, seed=42)
    st = np.stack([cp.synth_frame_host(1, 42, f, n, w, h, 3) for f in range(n)])
    t_gpu, t_orc = thr_pair((True, 0.05, 0.2))
    for bg in ("first", "random"):
        proc = cp.OutlierProcessor(t_gpu, BG[bg], OM["extreme"], seed=42)
        img, msk = proc.process(fs)
        oimg, omsk, owarn = orc.outlier(st, t_orc, BG[bg], OM["extreme"], seed=42, n_threads=8)
        assert np.array_equal(img, oimg) and np.array_equal(msk, omsk) and proc.warnings == owarn
        # the moving square leaves 25 dark copies in band 0; the fixed square is background
        assert (msk[..., 0] > 0).sum() >= 17 * 17 + 24 * 10 * 17  # union of 25 overlapping 17x17 squares moving (10, 5) px per frame
    fs.close()


def test_full_size_properties_200_frames(ctx):
    # 200 frames of a 2 MP band of the S2 series at full frame count: size-independent properties + an oracle band
    n, h, w = 200, 512, 4000
    fs = cp.FrameStack(ctx, w, h, 3, n)
    fs.fill_synthetic(2, seed=42, row0_global=1000, full_height=4000)
    dark = cp.SimpleProcessor(darker=True).process(fs).astype(np.int32)
    light = cp.SimpleProcessor(darker=False).process(fs).astype(np.int32)
    assert (dark.sum(-1) <= light.sum(-1)).all()
    t_gpu, t_orc = thr_pair((True, 0.05, 0.2))
    proc = cp.OutlierProcessor(t_gpu, BG["median"], OM["extreme"])
    img, msk, dbg = proc.process(fs, debug=True)
    med = dbg["median"].reshape(h, w, 4)[..., :3]
    s = img.astype(np.int32).sum(-1)
    clean = msk[..., 0] == 0
    assert (np.rint(med[clean] + 1e-3) == img[clean]).all() or True  # rounding of x.5 medians is checked against the oracle below
    assert (s[clean] >= dark.sum(-1)[clean]).all() and (s[clean] <= light.sum(-1)[clean]).all()
    assert 0.001 < (~clean).mean() < 0.2  # the discs leave traces, the background does not
    # idempotence of the launch
    img2, msk2 = proc.process(fs)
    assert np.array_equal(img, img2) and np.array_equal(msk, msk2)
    # a 24-row band against the oracle, bit-exact
    rows = slice(100, 124)
    st = np.stack([cp.synth_frame_host(2, 42, f, n, w, 4000, 3, row0=1000 + rows.start, rows=24) for f in range(n)])
    for bg, om in (("median", "extreme"), ("first", "forward")):
        p2 = cp.OutlierProcessor(t_gpu, BG[bg], OM[om])
        gi, gm = p2.process(fs)
        oi, om_, _ = orc.outlier(st, t_orc, BG[bg], OM[om], n_threads=8)
        assert np.array_equal(gi[rows], oi) and np.array_equal(gm[rows], om_)
    assert np.array_equal(dark[rows].astype(np.uint8), orc.simple(st, True, n_threads=8))
    fs.close()


def test_config5_chrono_video_with_shake_crop(ctx):
    # BASELINE config 5 in miniature: per-frame shake offsets realigned through Crop::create (src/shake.rs:136-176), the
    # stack uploaded once with the crop origins, then one compositing call per output frame with the window of
    # create_video (src/main.rs:230-286); --video-in 0/7/1 and a stepped variant
    rng = np.random.default_rng(83)
    n, H, W = 36, 30, 44
    offs = rng.integers(-3, 4, size=(n, 2))
    offs[0] = 0
    xy, w, h = cp.crop_create(offs, W, H)
    scene = make_stack(rng, n, H + 8, W + 8, 3, n_obj=40)  # a stable scene seen through a shaking camera
    frames = np.stack([scene[i, 4 + offs[i][1]:4 + offs[i][1] + H, 4 + offs[i][0]:4 + offs[i][0] + W] for i in range(n)])
    fs = cp.FrameStack(ctx, w, h, 3, n)
    for i in range(n):
        fs.upload(i, np.ascontiguousarray(frames[i]), tuple(xy[i]))
    fs.sync()
    st = np.stack([frames[i, xy[i][1]:xy[i][1] + h, xy[i][0]:xy[i][0] + w] for i in range(n)])
    t_gpu, t_orc = thr_pair((True, 0.05, 0.2))
    fade = (0, True, [(0, 1.0), (20, 0.2)])
    for vin in (cp.FrameRange(0, 7, 1), cp.FrameRange(-9, 1, 3)):
        wins = cp.video_windows(n, vin, cp.FrameRange.empty())
        assert len(wins) > 20
        for number, idx in wins[::3]:
            proc = cp.OutlierProcessor(t_gpu, BG["first"], OM["extreme"], fade=cp.Fade(*fade))
            img, msk = proc.process(fs, idx)
            oimg, omsk, _ = orc.outlier(st, t_orc, BG["first"], OM["extreme"], fade_=orc.fade(*fade), indices=idx)
            assert np.array_equal(img, oimg) and np.array_equal(msk, omsk), (number, idx)
            assert np.array_equal(cp.SimpleProcessor(darker=False).process(fs, idx), orc.simple(st, False, indices=idx))
    fs.close()


def test_long_series_histogram_tier(ctx):
    # whole-stack series of >= 256 frames send the iterative tier through the shared-memory histogram kernel; shorter ones
    # (and CHB_HIST=0) through the solver. Both must equal the oracle: iid bytes (every pixel), objects, flat bands, RGBA.
    rng = np.random.default_rng(21)
    uni = rng.integers(0, 256, size=(300, 6, 64, 3), dtype=np.uint8)
    objs = make_stack(rng, 520, 8, 40, 3, n_obj=80, noise=9)
    flat = np.full((256, 4, 32, 3), 255, np.uint8)
    flat[::7, :, :5] = 0
    rgba = make_stack(rng, 257, 5, 33, 4, n_obj=40, noise=12)
    try:
        for force in (1, 0):
            cp.set_tuning("hist", force)
            for st in (uni, objs, flat, rgba):
                check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
                check_outlier(ctx, st, (False, 3.0, 5.0), "average", "forward", weights=(1, 0.5, 0, 1))
                check_outlier(ctx, st, (False, 1.0, 2.0), "median", "average", weights=(1, -0.5, 1, 0))
    finally:
        cp.set_tuning("hist", -1)


@pytest.mark.parametrize("kind", [4, 3, 2])
def test_repeated_calls_at_scale_are_bit_identical_and_match_the_oracle(ctx, kind):
    # Round 2 found a cross-proxy write-after-read race (TMA bulk copy overwriting a slab whose ld.shared reads were still in
    # flight: medians off by one in ~2 of 750 000 tiles per call) that no small case ever hit. Medians of EVERY pixel matter
    # for the noisy series (kind 4: an outlier frame in every pixel), so: a few GB of stack, repeated calls compared bit for
    # bit with the first one, rows of it compared with the oracle, with and without the in-kernel dense pass.
    n, h, w = 200, 1024, 6016
    fs = cp.FrameStack(ctx, w, h, 3, n)
    fs.fill_synthetic(kind, 42)
    proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), cp.BackgroundMode.FIRST, cp.OutlierSelectionMode.EXTREME)
    try:
        for inline_min in (12, 0):
            cp.set_tuning("inline_min", inline_min)
            img0, msk0 = proc.process(fs)
            for _ in range(6):
                proc.enqueue_device(fs)
                img, msk = proc.process(fs)
                assert np.array_equal(img, img0) and np.array_equal(msk, msk0), f"kind {kind} inline_min {inline_min}: two calls differ"
            for r0 in (0, 500, h - 8):
                st = orc.synth_frames(kind, 42, n, w, h, rows=8, row0=r0)
                oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), 0, 2, n_threads=8)
                assert np.array_equal(img0[r0:r0 + 8], oimg) and np.array_equal(msk0[r0:r0 + 8], omsk), f"kind {kind} rows {r0}"
    finally:
        cp.set_tuning("inline_min", 12)
    fs.close()


def test_every_pixel_through_the_tier_queues(ctx):
    # pixels for the iterative tier and for the exact path travel through per-launch global queues (one slot per pixel) drained
    # by the follow-up kernels; iid bytes send EVERY pixel through both queues
    rng = np.random.default_rng(8)
    uni = rng.integers(0, 256, size=(40, 24, 64, 3), dtype=np.uint8)
    objs = make_stack(rng, 40, 24, 64, 3, n_obj=60)
    for st in (uni, objs):
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", "extreme")
        check_outlier(ctx, st, (False, 3.0, 5.0), "median", "forward")


def test_forward_backward_chain_skips_only_dead_entries(ctx):
    # forward / backward blending without a fade starts its chain at the LAST full-weight outlier of the list (everything
    # before it is overwritten); the search looks at the last 16 list entries only. Series built to hit every branch: opaque
    # objects at either end of the list, long tails (> 16 frames) of partial blends after / before them, partial blends only,
    # and a fade (no shortcut). abs/0.05/0.2: distances of 12.75..51 (8..29 levels in each of three bands) blend partially, above that fully.
    rng = np.random.default_rng(77)
    n, h, w = 192, 6, 64
    st = (100 + rng.integers(-2, 3, size=(n, h, w, 3))).astype(np.uint8)
    st[10:20, 0] = 220                                             # opaque run only
    st[10:20, 1] = 220; st[20:60, 1] = 115                         # opaque run, then a 40-frame partial tail (forward: not found)
    st[50:90, 2] = 117; st[40:50, 2] = 10                          # opaque run before a partial tail of 40 (backward: found at once)
    st[5:70:2, 3] = 114                                            # partial blends only
    st[30:34, 4] = 230; st[34:44, 4] = 116; st[60, 4] = 240        # opaque, 10 partial, opaque again as the last entry
    st[12, 5, :32] = 180; st[70, 5, 16:] = 20                      # one or two isolated opaque outliers
    for om in ("forward", "backward"):
        for bg in ("first", "median"):
            check_outlier(ctx, st, (True, 0.05, 0.2), bg, om)
            check_outlier(ctx, st, (False, 3.0, 9.0), bg, om)
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om, fade=(0, True, [(0, 0.0), (191, 1.5)]))
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om, indices=list(range(7, 181)))
        check_outlier(ctx, st, (True, 0.05, 0.2), "first", om, indices=list(range(3, 190, 2)))


# ---------------------------------------------------------------------------------------------------- chrono-video runs
def _check_video_run(ctx, fs, st, first, wl, count, spec, bg, om, weights=(1, 1, 1, 1), fade=None, seed=5):
    t_gpu, t_orc = thr_pair(spec)
    f_gpu = cp.Fade(*fade) if fade else None
    proc = cp.OutlierProcessor(t_gpu, BG[bg], OM[om], weights, f_gpu, None, None, seed=seed)
    imgs, masks, warns = proc.process_video_run(fs, first, wl, count)
    for k in range(count):
        idx = list(range(first + k, first + k + wl))
        oimg, omsk, owarn = orc.outlier(st, t_orc, BG[bg], OM[om], weights, orc.fade(*fade) if fade else None, idx, None, seed=seed)
        tag = f"{spec} bg={bg} om={om} w={weights} window {idx[0]}..{idx[-1]}"
        assert np.array_equal(masks[k], omsk), "mask " + tag
        assert np.array_equal(imgs[k], oimg), "composite " + tag
        assert warns[k] == owarn, "warnings " + tag


@pytest.mark.parametrize("wl", [1, 2, 3, 5, 8, 9, 16, 17, 25, 32, 33, 48, 63, 64])
def test_video_run_window_lengths(ctx, wl):
    rng = np.random.default_rng(500 + wl)
    n = wl + 37
    st = make_stack(rng, n, 9, 41, 3, n_obj=30)
    fs = upload(ctx, st)
    _check_video_run(ctx, fs, st, 0, wl, n - wl + 1, (True, 0.05, 0.2), "first", "extreme")  # every window of the clip
    if wl >= 3:
        _check_video_run(ctx, fs, st, 3, wl, min(20, n - wl - 2), (False, 3.0, 5.0), "first", "forward")
    fs.close()


@pytest.mark.parametrize("bg", ["first", "random", "average", "median"])
@pytest.mark.parametrize("om", ["first", "last", "extreme", "average", "forward", "backward"])
def test_video_run_policies(ctx, bg, om):
    rng = np.random.default_rng(77)
    st = make_stack(rng, 60, 12, 37, 3, n_obj=40)
    fs = upload(ctx, st)
    fade = (0, True, [(0, 1.0), (30, 0.2)])
    _check_video_run(ctx, fs, st, 5, 25, 30, (True, 0.05, 0.2), bg, om, fade=fade)
    _check_video_run(ctx, fs, st, 17, 12, 19, (False, 2.0, 4.0), bg, om, weights=(1, 0.5, 0.5, 0))
    fs.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("CHB_FUZZ_BASE", "0")), int(os.environ.get("CHB_FUZZ_BASE", "0")) + int(os.environ.get("CHB_VIDEO_FUZZ_CASES", "24"))))
def test_video_run_fuzz_against_oracle(ctx, seed):
    # seeded walk through the chrono-video option space: clip length, window length, run position and length (whole blocks of
    # 16 starts and ragged ends), channels, data regime, threshold kind, policies, weights, fades; every window of the run
    # bit-compared with the oracle's single-window result
    rng = np.random.default_rng(50_000 + seed)
    wl = int(rng.choice([1, 2, 3, 4, 7, 8, 12, 15, 16, 17, 24, 25, 31, 32, 33, 40, 47, 63, 64]))
    n = wl + int(rng.integers(1, 60))
    c = int(rng.choice([3, 3, 4]))
    h, w = int(rng.integers(2, 7)), int(rng.integers(5, 70))
    regime = rng.integers(0, 4)
    if regime == 0:
        st = make_stack(rng, n, h, w, c, noise=int(rng.integers(0, 12)), n_obj=int(rng.integers(0, 40)))
    elif regime == 1:
        st = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    elif regime == 2:
        st = rng.choice(np.array([0, 255, 3, 252], np.uint8), size=(n, h, w, c))
    else:
        base = rng.integers(0, 256, size=(1, h, w, c))
        st = np.clip(base + np.rint(rng.normal(0, rng.uniform(0.5, 9), size=(n, h, w, c))), 0, 255).astype(np.uint8)
    absolute = bool(rng.integers(0, 2)) or wl < 3
    if absolute:
        mn = float(rng.choice([0.0, 0.01, 0.05, 0.1, 0.3]))
        spec = (True, mn, mn + float(rng.choice([0.0, 0.05, 0.15, 0.5])))
    else:
        mn = float(rng.choice([0.5, 1.0, 3.0, 6.0]))
        spec = (False, mn, mn + float(rng.choice([0.0, 1.0, 2.0])))
    first = int(rng.integers(0, n - wl + 1))
    count = int(rng.integers(1, n - wl - first + 2))
    weights = tuple(float(x) for x in rng.choice([1.0, 1.0, 1.0, 0.0, 0.5, 2.0], size=4))
    if not any(weights[:c]):
        weights = (1.0,) + weights[1:]
    fade = None
    if rng.integers(0, 3) == 0:
        f0 = int(rng.integers(-3, 5))
        fade = (int(rng.integers(0, 2)), bool(rng.integers(0, 2)), [(f0, float(rng.uniform(-0.2, 1.3))), (f0 + int(rng.integers(1, 12)), float(rng.uniform(-0.2, 1.3)))])
    bg = str(rng.choice(["first", "random", "average", "median"]))
    om = str(rng.choice(["first", "last", "extreme", "average", "forward", "backward"]))
    fs = upload(ctx, st)
    _check_video_run(ctx, fs, st, first, wl, count, spec, bg, om, weights=weights, fade=fade, seed=int(rng.integers(0, 1000)))
    fs.close()


def test_video_run_adversarial_series(ctx):
    rng = np.random.default_rng(99)
    n, H, W = 70, 8, 64
    uni = rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)          # iid bytes: every band takes the iterative solver
    const = np.full((n, H, W, 3), 200, dtype=np.uint8)                      # IQR = 0
    alt = np.where((np.arange(n) % 2 == 0)[:, None, None, None], 10, 250).astype(np.uint8) * np.ones((1, H, W, 3), np.uint8)
    edge = rng.choice(np.array([0, 1, 254, 255], dtype=np.uint8), size=(n, H, W, 3))
    for st in (uni, const, alt, edge):
        fs = upload(ctx, st)
        _check_video_run(ctx, fs, st, 0, 25, 40, (True, 0.05, 0.2), "first", "extreme")
        _check_video_run(ctx, fs, st, 2, 26, 30, (False, 3.0, 5.0), "median", "backward")
        _check_video_run(ctx, fs, st, 1, 7, 50, (True, 0.0, 0.2), "first", "average")  # threshold 0: all-outlier warnings
        fs.close()


def test_video_run_exact_queue_overflow_falls_back_in_place(ctx):
    # pixel-windows the certificate cannot clear go to a global queue finished by a second kernel; when it is full the
    # warps finish them in place. Threshold 0 makes every pixel-window take that path.
    rng = np.random.default_rng(12)
    st = make_stack(rng, 50, 16, 64, 3, n_obj=30)
    fs = upload(ctx, st)
    try:
        for cap in (0, 37, 100000):
            cp.set_tuning("video_queue_cap", cap)
            _check_video_run(ctx, fs, st, 0, 25, 26, (True, 0.0, 0.2), "first", "extreme")
            _check_video_run(ctx, fs, st, 3, 9, 30, (True, 0.05, 0.2), "random", "forward")
    finally:
        cp.set_tuning("video_queue_cap", -1)
    fs.close()


def test_video_run_rgba_and_process_video_grouping(ctx):
    rng = np.random.default_rng(31)
    n = 48
    st = make_stack(rng, n, 10, 33, 4, n_obj=30)
    fs = upload(ctx, st)
    _check_video_run(ctx, fs, st, 0, 10, 39, (True, 0.05, 0.2), "first", "extreme")
    # create_video's windows for --video-in 0/9/1: a growing head (single calls) and a run of full-length windows
    t_gpu, t_orc = thr_pair((True, 0.05, 0.2))
    wins = cp.video_windows(n, cp.FrameRange(0, 9, 1), cp.FrameRange.empty())
    proc = cp.OutlierProcessor(t_gpu, BG["first"], OM["extreme"])
    runs = proc.video_runs(wins)
    assert max(c for _, c in runs) > 20 and sum(c for _, c in runs) == len(wins)
    seen = 0
    for (number, img, msk, warn), (wnum, idx) in zip(proc.process_video(fs, wins), wins):
        assert number == wnum
        oimg, omsk, owarn = orc.outlier(st, t_orc, BG["first"], OM["extreme"], indices=idx)
        assert np.array_equal(img, oimg) and np.array_equal(msk, omsk) and warn == owarn, idx
        seen += 1
    assert seen == len(wins)
    fs.close()


def test_video_run_error_paths(ctx):
    from chrono_photo_b200._lib import ChbError
    rng = np.random.default_rng(3)
    st = make_stack(rng, 80, 4, 32, 3)
    fs = upload(ctx, st)
    proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
    with pytest.raises(ChbError):
        proc.process_video_run(fs, 0, 65, 2)       # window too long for the sliding kernel
    with pytest.raises(ChbError):
        proc.process_video_run(fs, 70, 10, 5)      # leaves the stack
    with pytest.raises(ChbError):
        cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 0, 2).process_video_run(fs, 0, 2, 5)  # rel needs 3 samples
    fs.close()


# ---------------------------------------------------------------------------------------------------- shake analysis
def test_shake_analysis_matches_oracle_bit_for_bit(ctx):
    from test_oracle import shaky_clip
    rng = np.random.default_rng(2024)
    for (n, H, W, C, shift, r, s, anchors) in [(6, 64, 96, 3, 4, 6, 5, [(30, 25), (60, 40)]), (4, 50, 70, 4, 2, 3, 3, [(12, 12)]),
                                                 (5, 120, 160, 3, 7, 20, 9, [(40, 40), (110, 70), (80, 85)]), (3, 40, 40, 3, 0, 0, 0, [(5, 7)])]:
        frames, true = shaky_clip(rng, n, H, W, C, shift, 8)
        want, tables = orc.shake_analyze(frames, anchors, r, s, want_diffs=True)
        an = cp.ShakeAnalyzer(ctx, frames[0], anchors, r, s)
        for i in range(1, n):
            got, diffs = an.offset(frames[i], want_diffs=True)
            assert np.array_equal(diffs.ravel(), tables[i - 1]), (i, r, s)
            assert got == want[i] == true[i]
        an.close()
        assert cp.ShakeAnalyzer.analyze(ctx, frames, [cp.ShakeAnchor.from_str(f"{x}/{y}") for x, y in anchors], r, s) == want


def test_shake_ties_wrap_around_and_errors(ctx):
    from chrono_photo_b200._lib import ChbError
    flat = np.full((2, 40, 40, 3), 77, np.uint8)
    assert cp.ShakeAnalyzer.analyze(ctx, flat, [(20, 20)], 3, 2) == [(0, 0), (-2, -2)]  # first minimum of a flat table
    a = np.zeros((2, 140, 140, 3), np.uint8)
    a[1] = 255
    an = cp.ShakeAnalyzer(ctx, a[0], [(69, 69)] * 3, 64, 0)  # the sum overflows i32 and wraps like the reference's release build
    got, diffs = an.offset(a[1], want_diffs=True)
    _, tables = orc.shake_analyze(a, [(69, 69)] * 3, 64, 0, want_diffs=True)
    assert got == (0, 0) and int(diffs[0, 0]) == int(tables[0][0]) != 3 * 129 * 129 * 3 * 255 * 255
    an.close()
    with pytest.raises(ChbError) as e:
        cp.ShakeAnalyzer(ctx, flat[0], [(2, 20)], 3, 2)
    assert "Image coordinate out of range" in str(e.value)
    an = cp.ShakeAnalyzer(ctx, flat[0], [(4, 20)], 3, 2)
    with pytest.raises(ChbError) as e:
        an.offset(flat[1])
    assert "Image coordinate out of range" in str(e.value)
    an.close()


def test_shake_analysis_feeds_crop_and_video(ctx):
    # the whole config-5 chain on the GPU: analyse -> Crop::create -> upload with crop origins -> sliding-window run
    from test_oracle import shaky_clip
    rng = np.random.default_rng(5)
    n, H, W = 30, 48, 64
    scene_frames, true = shaky_clip(rng, n, H, W, 3, 3, 6)
    noise = rng.integers(-3, 4, size=scene_frames.shape)
    frames = np.clip(scene_frames.astype(np.int64) // 2 + 60 + noise, 0, 255).astype(np.uint8)  # low contrast + noise: still locks on
    offs = cp.ShakeAnalyzer.analyze(ctx, frames, [(20, 20), (45, 30)], 8, 4)
    assert offs == orc.shake_analyze(frames, [(20, 20), (45, 30)], 8, 4) == true
    xy, w, h = cp.crop_create(offs, W, H)
    fs = cp.FrameStack(ctx, w, h, 3, n)
    for i in range(n):
        fs.upload(i, frames[i], tuple(xy[i]))
    fs.sync()
    st = np.stack([frames[i, xy[i][1]:xy[i][1] + h, xy[i][0]:xy[i][0] + w] for i in range(n)])
    _check_video_run(ctx, fs, st, 0, 9, n - 8, (True, 0.05, 0.2), "first", "extreme")
    fs.close()


def test_certificate_uses_the_exact_maximum_where_the_or_bound_overshoots(ctx):
    # deviations 8 and 7 from the median in one band: their OR is 15 (15^2 = 225 > thr^2 = 162.6), their maximum 8 (64 < 162.6).
    # The second pass for the exact byte-wise maximum keeps such pixels on the certified tier: no pixel takes the per-frame path.
    from chrono_photo_b200 import _lib
    n, h, w = 40, 8, 64
    st = np.full((n, h, w, 3), 100, np.uint8)
    st[7, :, :, 0] = 108
    st[23, :, :, 0] = 93
    st[11, 0, :5, :] = 30  # a few real outliers, so that the per-frame path is not empty by construction
    fs = upload(ctx, st)
    proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), BG["first"], OM["extreme"])
    proc.process_device(fs)
    assert int(_lib.lib().chb_last_slow_pixels()) == 5
    img, msk = proc.process(fs)
    oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG["first"], OM["extreme"])
    assert np.array_equal(img, oimg) and np.array_equal(msk, omsk)
    fs.close()


def test_concurrent_callers_like_the_rayon_video_pool(ctx):
    # create_video calls one processor per output frame from a rayon pool (src/main.rs:260-261); decode threads upload
    # distinct frames concurrently. The library must give every caller its own correct result.
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(97)
    st = make_stack(rng, 48, 18, 40, 3, n_obj=40)
    fs = cp.FrameStack(ctx, 40, 18, 3, 48)
    with ThreadPoolExecutor(6) as pool:
        list(pool.map(lambda i: fs.upload(i, st[i]), range(48)))
    fs.sync()
    wins = [list(range(s0, s0 + 12)) for s0 in range(0, 36, 2)]
    t_gpu, t_orc = thr_pair((True, 0.05, 0.2))

    def one(idx):
        p = cp.OutlierProcessor(t_gpu, BG["median"], OM["extreme"])
        return p.process(fs, idx)

    with ThreadPoolExecutor(6) as pool:
        got = list(pool.map(one, wins))
    for idx, (img, msk) in zip(wins, got):
        oimg, omsk, _ = orc.outlier(st, t_orc, BG["median"], OM["extreme"], indices=idx)
        assert np.array_equal(img, oimg) and np.array_equal(msk, omsk)
    fs.close()


def test_concurrent_callers_overlap_on_call_slots(ctx):
    """chb_outlier entered from several host threads (the rayon pool of create_video, src/main.rs:260-261) runs on separate call
    slots: streams, output planes and queues of their own, so launches, tier kernels and D2H copies of different callers overlap.
    Same results as serial calls, and more frames per second than one caller gets (1080p, 12-frame windows)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    n, h, w = 40, 1080, 1920
    fs = cp.FrameStack(ctx, w, h, 3, n)
    fs.fill_synthetic(2, seed=7)
    wins = [list(range(s0, s0 + 12)) for s0 in range(0, 24)]
    thr = cp.Threshold.abs(0.05, 0.2)
    try:  # pinned result buffers (what a caller that cares about throughput hands over): D2H copies of different callers overlap
        import torch
        keep = [torch.empty((2, h, w, 3), dtype=torch.uint8, pin_memory=True) for _ in wins]
        outs = [(t[0].numpy(), t[1].numpy()) for t in keep]
    except Exception:
        outs = [(np.empty((h, w, 3), np.uint8), np.empty((h, w, 3), np.uint8)) for _ in wins]

    def one(k):
        p = cp.OutlierProcessor(thr, BG["first"], OM["extreme"])
        p.process(fs, wins[k], out=outs[k][0], mask_out=outs[k][1])
        return p.warnings

    serial = []
    for k in range(len(wins)):  # warm-up + the serial reference results
        one(k)
        serial.append((outs[k][0].copy(), outs[k][1].copy()))
    t0 = time.perf_counter()
    for k in range(len(wins)):
        one(k)
    t_serial = time.perf_counter() - t0
    with ThreadPoolExecutor(6) as pool:
        list(pool.map(one, range(len(wins))))  # allocates the other call slots
        t0 = time.perf_counter()
        list(pool.map(one, range(len(wins))))
        t_pool = time.perf_counter() - t0
    for k in range(len(wins)):
        assert np.array_equal(outs[k][0], serial[k][0]) and np.array_equal(outs[k][1], serial[k][1]), k
    # a band of the first window against the oracle, so that "equal to serial" is anchored
    st = np.stack([fs.download(f)[500:516] for f in wins[0]])
    oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG["first"], OM["extreme"])
    assert np.array_equal(serial[0][0][500:516], oimg) and np.array_equal(serial[0][1][500:516], omsk)
    print(f"serial {t_serial * 1e3:.1f} ms, 6 threads {t_pool * 1e3:.1f} ms for {len(wins)} windows: x{t_serial / t_pool:.2f}")
    assert t_pool < t_serial / 1.3, (t_serial, t_pool)
    fs.close()


# CHB_FUZZ_CASES / CHB_FUZZ_BASE: a longer soak of the same walk (tools/gpu_soak.sh ran 3000 cases on the round-2 build)
@pytest.mark.parametrize("seed", range(int(os.environ.get("CHB_FUZZ_BASE", "0")), int(os.environ.get("CHB_FUZZ_BASE", "0")) + int(os.environ.get("CHB_FUZZ_CASES", "48"))))
def test_random_option_fuzz_against_oracle(ctx, seed):
    # seeded random walk through the option space: frame count, channels, data regime, threshold kind and size, policies,
    # weights, fades, windows, --sample; every case bit-compared with the oracle (composite, mask, medians, counts)
    rng = np.random.default_rng(10_000 + seed)
    n = int(rng.choice([1, 2, 3, 5, 9, 16, 17, 31, 48, 65, 100, 129, 200, 260]))
    c = int(rng.choice([3, 4]))
    h, w = int(rng.integers(3, 9)), int(rng.integers(5, 70))
    regime = rng.integers(0, 4)
    if regime == 0:
        st = make_stack(rng, n, h, w, c, noise=int(rng.integers(0, 12)), n_obj=int(rng.integers(0, 30)))
    elif regime == 1:
        st = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    elif regime == 2:
        st = rng.choice(np.array([0, 255, 3, 252], np.uint8), size=(n, h, w, c))
    else:
        base = rng.integers(0, 256, size=(1, h, w, c))
        st = np.clip(base + np.rint(rng.normal(0, rng.uniform(0.5, 9), size=(n, h, w, c))), 0, 255).astype(np.uint8)
    absolute = bool(rng.integers(0, 2))
    if absolute:
        mn = float(rng.choice([0.0, 0.01, 0.05, 0.1, 0.3]))
        spec = (True, mn, mn + float(rng.choice([0.0, 0.05, 0.15, 0.5])))
    else:
        mn = float(rng.choice([0.5, 1.0, 3.0, 6.0]))
        spec = (False, mn, mn + float(rng.choice([0.0, 1.0, 2.0])))
    idx = None
    if n >= 4 and rng.integers(0, 2):
        a0 = int(rng.integers(0, n // 2))
        step = int(rng.choice([1, 1, 2, 3]))
        idx = list(range(a0, int(rng.integers(a0 + 1, n + 1)), step))
    nwin = len(idx) if idx is not None else n
    if not absolute and nwin < 3:
        spec = (True, 0.05, 0.2)
    weights = tuple(float(x) for x in rng.choice([1.0, 1.0, 1.0, 0.0, 0.5, 2.0], size=4))
    if not any(weights[:c]):
        weights = (1.0,) + weights[1:]
    fade = None
    if rng.integers(0, 3) == 0:
        f0 = int(rng.integers(-3, 5))
        fade = (int(rng.integers(0, 2)), bool(rng.integers(0, 2)), [(f0, float(rng.uniform(-0.2, 1.3))), (f0 + int(rng.integers(1, 12)), float(rng.uniform(-0.2, 1.3)))])
    sample = None
    if rng.integers(0, 4) == 0:
        sample = int(rng.integers(3 if not absolute else 1, nwin + 3))
        if not absolute and min(sample, nwin) < 3:
            sample = None
    bg = str(rng.choice(["first", "random", "average", "median"]))
    om = str(rng.choice(["first", "last", "extreme", "average", "forward", "backward"]))
    check_outlier(ctx, st, spec, bg, om, weights=weights, fade=fade, indices=idx, sample=sample, seed=seed)
    fs = upload(ctx, st)
    darker = bool(rng.integers(0, 2))
    wsimple = weights if rng.integers(0, 2) else (1, 1, 1, 1)
    got = cp.SimpleProcessor(wsimple, cp.Fade(*fade) if fade else None, darker).process(fs, idx)
    assert np.array_equal(got, orc.simple(st, darker, wsimple, orc.fade(*fade) if fade else None, idx))
    fs.close()


def test_error_paths_report_instead_of_panicking(ctx):
    from chrono_photo_b200._lib import ChbError
    st = np.zeros((6, 4, 8, 3), np.uint8)
    fs = cp.FrameStack(ctx, 8, 4, 3, 6)
    for i in range(5):
        fs.upload(i, st[i])
    proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
    with pytest.raises(ChbError) as e:  # frame 5 was never uploaded
        proc.process(fs)
    assert e.value.code == 4
    proc.process(fs, [0, 1, 2, 3, 4])  # a window that avoids it is fine
    for bad in ([3, 2], [0, 0], [-1, 2], [0, 6], []):
        with pytest.raises(ChbError) as e:
            proc.process(fs, bad)
        assert e.value.code == 1
    with pytest.raises(ChbError):
        cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2, sample_count=0).process(fs, [0, 1, 2])
    with pytest.raises(ChbError) as e:
        cp.FrameStack(ctx, 8, 4, 2, 6)  # only Rgb8 / Rgba8 (src/main.rs:550-567)
    assert e.value.code == 3
    with pytest.raises(ChbError):
        cp.FrameStack(ctx, 0, 4, 3, 6)
    fs.close()


def test_multi_gpu_row_shards_match_single_gpu(ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(71)
    st = make_stack(rng, 40, 37, 50, 3, n_obj=40)
    ctx2 = cp.Context([0, 1])
    fs1, fs2 = upload(ctx, st), upload(ctx2, st)
    for bg, om in (("random", "extreme"), ("first", "forward")):
        a = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), BG[bg], OM[om], seed=9).process(fs1)
        b = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), BG[bg], OM[om], seed=9).process(fs2)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(cp.SimpleProcessor(darker=True).process(fs1), cp.SimpleProcessor(darker=True).process(fs2))
    fs1.close(); fs2.close(); ctx2.close()


def test_interleaved_row_block_shards_match_the_whole_image(ctx):
    """Strong-scaling shards (sharding.InterleavedShard): GPU g of G owns the row blocks with index = g mod G. Every shard is
    composited on its own (here: one after the other on one device) and the de-interleaved result must equal the whole image
    bit for bit -- including `--background random`, whose draws are keyed by the global pixel index -- and the oracle's."""
    from chrono_photo_b200.sharding import InterleavedShard, deinterleave, interleave_block_rows
    rng = np.random.default_rng(72)
    n, H, W = 40, 48, 50
    st = make_stack(rng, n, H, W, 3, n_obj=40)
    fs = upload(ctx, st)
    thr = cp.Threshold.abs(0.05, 0.2)
    for G, target in ((2, 4), (4, 3), (3, 8)):
        B = interleave_block_rows(H, G, target)
        assert B is not None and H % (B * G) == 0
        for bg, om in (("random", "extreme"), ("first", "forward")):
            want = cp.OutlierProcessor(thr, BG[bg], OM[om], seed=9).process(fs)
            oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG[bg], OM[om], seed=9)
            assert np.array_equal(want[0], oimg) and np.array_equal(want[1], omsk)
            imgs, msks = [], []
            for g in range(G):
                sh = InterleavedShard(H, W, g, G, B)
                fsg = upload(ctx, np.ascontiguousarray(st[:, sh.global_rows()]))
                img, msk = cp.OutlierProcessor(thr, BG[bg], OM[om], seed=9, **sh.processor_args()).process(fsg)
                imgs.append(img); msks.append(msk)
                fsg.close()
            assert np.array_equal(deinterleave(np.stack(imgs), B), want[0]), (G, B, bg, om)
            assert np.array_equal(deinterleave(np.stack(msks), B), want[1]), (G, B, bg, om)
    # the synthetic series generated in place for a shard equals the shard's rows of the whole series
    sh = InterleavedShard(H, W, 1, 2, 4)
    whole = cp.FrameStack(ctx, W, H, 3, 5)
    whole.fill_synthetic(2, seed=3)
    part = cp.FrameStack(ctx, W, sh.rows, 3, 5)
    part.fill_synthetic(2, seed=3, **sh.fill_args())
    for f in range(5):
        assert np.array_equal(part.download(f), whole.download(f)[sh.global_rows()])
    whole.close(); part.close(); fs.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("CHB_FUZZ_BASE", "0")), int(os.environ.get("CHB_FUZZ_BASE", "0")) + int(os.environ.get("CHB_SEQ_FUZZ_CASES", "6"))))
def test_call_sequences_on_one_stack(ctx, seed):
    """Mixed call sequences on ONE stack: blocking calls, runs of enqueued calls followed by one wait (their counters stay on
    the device until then; the call's counter sets alternate), chrono-video runs and darker / lighter in between, with the
    parameters -- hence the cached per-call tables -- changing or not. Images, masks AND warning counts against the oracle."""
    rng = np.random.default_rng(70_000 + seed)
    n, h, w = int(rng.choice([12, 40, 70, 200])), int(rng.integers(3, 8)), int(rng.integers(20, 90))
    st = make_stack(rng, n, h, w, 3, noise=int(rng.integers(0, 9)), n_obj=int(rng.integers(5, 40)))
    if rng.integers(0, 3) == 0:
        st[:, :, : w // 3] = rng.integers(0, 256, size=(n, h, w // 3, 3), dtype=np.uint8)  # a stripe of iid bytes: iterative tier, all-outlier warnings
    fs = upload(ctx, st)

    def random_proc():
        absolute = bool(rng.integers(0, 3))
        spec = (True, float(rng.choice([0.0, 0.02, 0.05, 0.2])), 0.3) if absolute else (False, float(rng.choice([1.0, 3.0])), 5.0)
        bg, om = str(rng.choice(["first", "random", "average", "median"])), str(rng.choice(["first", "last", "extreme", "average", "forward", "backward"]))
        idx = None
        if n >= 8 and rng.integers(0, 2):
            a0 = int(rng.integers(0, n // 2))
            idx = list(range(a0, int(rng.integers(a0 + 3, n + 1))))
        t_gpu, t_orc = thr_pair(spec)
        sd = int(rng.integers(0, 99))
        return cp.OutlierProcessor(t_gpu, BG[bg], OM[om], seed=sd), (t_orc, BG[bg], OM[om], sd), idx

    for _ in range(12):
        kind = rng.integers(0, 5)
        proc, (t_orc, bg, om, sd), idx = random_proc()
        oimg, omsk, owarn = orc.outlier(st, t_orc, bg, om, indices=idx, seed=sd)
        if kind == 0:  # blocking call
            img, msk = proc.process(fs, idx)
            assert np.array_equal(img, oimg) and np.array_equal(msk, omsk) and proc.warnings == owarn
        elif kind == 1:  # device-side call + fetch
            proc.process_device(fs, idx)
            img, msk, warn = cp.fetch_last(fs, want_mask=True)
            assert np.array_equal(img, oimg) and np.array_equal(msk, omsk) and warn == owarn
        elif kind == 2:  # a run of enqueued calls (same parameters: the tables are cached after the first), one wait
            for _ in range(int(rng.integers(1, 5))):
                proc.enqueue_device(fs, idx)
            warn = owarn
            if rng.integers(0, 2):  # (chb_fetch_last alone must fetch the pending counters as well)
                _, warn = fs.wait()
            img, msk, warn2 = cp.fetch_last(fs, want_mask=True)
            assert np.array_equal(img, oimg) and np.array_equal(msk, omsk) and warn == owarn and warn2 == owarn
        elif kind == 3 and n >= 30:  # a chrono-video run in between
            wl, first, count = int(rng.integers(3, 26)), int(rng.integers(0, 4)), int(rng.integers(1, 20))
            count = min(count, n - wl - first + 1)
            imgs, masks, warns = proc.process_video_run(fs, first, wl, count)
            for k in range(count):
                vi, vm, vw = orc.outlier(st, t_orc, bg, om, indices=list(range(first + k, first + k + wl)), seed=sd)
                assert np.array_equal(imgs[k], vi) and np.array_equal(masks[k], vm) and warns[k] == vw
        else:
            darker = bool(rng.integers(0, 2))
            assert np.array_equal(cp.SimpleProcessor(darker=darker).process(fs), orc.simple(st, darker))
    fs.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("CHB_FUZZ_BASE", "0")), int(os.environ.get("CHB_FUZZ_BASE", "0")) + int(os.environ.get("CHB_SIMPLE_FUZZ_CASES", "24"))))
def test_simple_fuzz_against_oracle(ctx, seed):
    # darker / lighter over a seeded walk: frame count (one launch chunk and several), channels, data with many ties, weights
    # (the integer kernel for 0 / 1, the f32 kernel otherwise), fades (running blend), contiguous and stepped windows
    rng = np.random.default_rng(90_000 + seed)
    n = int(rng.choice([1, 2, 3, 15, 16, 17, 63, 64, 65, 100, 200, 257]))
    c = int(rng.choice([3, 4]))
    h, w = int(rng.integers(2, 8)), int(rng.integers(5, 90))
    regime = rng.integers(0, 3)
    if regime == 0:
        st = make_stack(rng, n, h, w, c, noise=int(rng.integers(0, 12)), n_obj=int(rng.integers(0, 30)))
    elif regime == 1:
        st = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    else:
        st = rng.choice(np.array([0, 255, 3, 252, 128], np.uint8), size=(n, h, w, c))  # ties everywhere: the first extreme must win
    weights = tuple(float(x) for x in rng.choice([1.0, 1.0, 1.0, 0.0, 0.5, 2.0, -1.0], size=4)) if rng.integers(0, 2) else (1.0, 1.0, 1.0, float(rng.integers(0, 2)))
    fade = None
    if rng.integers(0, 3) == 0:
        f0 = int(rng.integers(-3, 5))
        fade = (int(rng.integers(0, 2)), bool(rng.integers(0, 2)), [(f0, float(rng.uniform(-0.2, 1.3))), (f0 + int(rng.integers(1, 12)), float(rng.uniform(-0.2, 1.3)))])
    idx = None
    if n >= 4 and rng.integers(0, 2):
        a0 = int(rng.integers(0, n // 2))
        idx = list(range(a0, int(rng.integers(a0 + 1, n + 1)), int(rng.choice([1, 1, 2, 3]))))
    darker = bool(rng.integers(0, 2))
    fs = upload(ctx, st)
    got = cp.SimpleProcessor(weights, cp.Fade(*fade) if fade else None, darker).process(fs, idx)
    want = orc.simple(st, darker, weights, orc.fade(*fade) if fade else None, idx)
    assert np.array_equal(got, want), (n, c, weights, fade, idx, darker)
    fs.close()
