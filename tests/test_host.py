"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/chrono_b200.h declares, the host
arithmetic it exports matches the oracle, option grammars match the reference's FromStr impls, and compute entry
points fail loudly without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as orc

import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "chrono_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(chb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 25
    l = _lib.lib()
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert l.chb_version() == 200


def test_no_cpu_fallback():
    if _has_gpu():
        pytest.skip("device present")
    with pytest.raises(_lib.ChbError) as e:
        cp.Context()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_threshold_matches_oracle():
    for absolute, mn, mx in ((True, 0.05, 0.2), (False, 3.0, 5.0), (True, 0.1, 0.1), (True, 0.033, 0.7)):
        t, o = cp.Threshold(absolute, mn, mx), orc.threshold(absolute, mn, mx)
        assert (np.float32(t.min), np.float32(t.max)) == (np.float32(o.min), np.float32(o.max))
        assert np.float32(t.scale) == np.float32(o.scale) or (np.isinf(t.scale) and np.isinf(o.scale))
    t = cp.Threshold.from_str("abs/0.05/0.2")  # CLI default, src/cli.rs:192
    assert t.absolute and abs(t.min - 12.75) < 1e-5
    assert not cp.Threshold.from_str("relative/3").absolute and cp.Threshold.from_str("rel/3").max == 3.0
    with pytest.raises(cp.ParseOptionError):
        cp.Threshold.from_str("foo/1/2")


def test_fade_matches_oracle():
    frames = [(-3, 0.2), (0, 1.0), (7, 0.0), (9, 0.5)]
    f = cp.Fade(cp.FadeMode.REPEAT, True, frames)
    fo, keep = orc.fade(1, True, frames)
    assert np.array_equal(f.values, keep) and f.offset == fo.offset
    for fr in range(-30, 30):
        assert np.float32(f.get(fr)) == np.float32(orc.lib().orc_fade_get(fo, fr))
    s = cp.Fade.from_str("clamp/abs/0,0/10,1")  # src/options.rs:352
    assert len(s.values) == 11 and s.absolute and s.mode == cp.FadeMode.CLAMP
    assert cp.Fade.none().get(123) == 1.0
    with pytest.raises(cp.ParseOptionError):
        cp.Fade.from_str("clamp/abs/0;0")
    with pytest.raises(cp.ParseEnumError):
        cp.Fade.from_str("mirror/abs/0,0/10,1")


def test_enum_grammars():
    assert cp.OutlierSelectionMode.from_str("forward") == cp.OutlierSelectionMode.ALL_FORWARD
    assert cp.OutlierSelectionMode.from_str("backward") == cp.OutlierSelectionMode.ALL_BACKWARD
    assert [int(cp.BackgroundMode.from_str(s)) for s in ("first", "random", "average", "median")] == [0, 1, 2, 3]
    assert cp.SelectionMode.from_str("darker") == cp.SelectionMode.DARKER
    for bad, cls in (("brightest", cp.SelectionMode), ("mean", cp.BackgroundMode), ("all", cp.OutlierSelectionMode)):
        with pytest.raises(cp.ParseEnumError):
            cls.from_str(bad)
    fr = cp.FrameRange.from_str("-4/./2")
    assert (fr.start, fr.end, fr.step) == (-4, None, 2) and fr.range() is None
    assert cp.FrameRange.from_str("0/25/1").range() == 25
    with pytest.raises(cp.ParseOptionError):
        cp.FrameRange.from_str("0/25")


def test_crop_create_matches_oracle():
    rng = np.random.default_rng(43)
    off = rng.integers(-8, 9, size=(50, 2))
    off[0] = 0
    a, b = cp.crop_create(off, 1920, 1080), orc.crop_create(off, 1920, 1080)
    assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    assert cp.crop_create(np.zeros((5, 2)), 10, 10) is None


@pytest.mark.parametrize("vin,vout,count", [((0, 25, 1), (None, None, 1), 1800), ((None, 10, 1), (None, None, 1), 40),
                                            ((-12, 3, 3), (0, 60, 2), 60), ((-5, None, 1), (None, None, 1), 33)])
def test_video_windows_match_oracle(vin, vout, count):
    got = cp.video_windows(count, cp.FrameRange(*vin), cp.FrameRange(*vout))
    n, ws, we, num = orc.video_windows(count, vin, vout, cap=8192)
    want = [(int(num[i]), list(range(int(ws[i]), int(we[i]), vin[2]))) for i in range(n) if ws[i] < we[i]]
    assert got == want


def test_synth_generator_recipe_s1():
    # recipe of src/util/create_example_data.rs:10-49
    f = cp.synth_frame_host(1, 42, 3, 25, 1024, 768, 3)
    assert f.shape == (768, 1024, 3)
    assert f[..., 2].min() >= 140 and f[..., 2].max() < 150
    bg = f[10, 10]
    assert 240 <= bg[0] < 250 and 240 <= bg[1] < 250
    cx, cy = 100 + 3 * 10, 768 // 3 + 3 * 5
    assert (f[cy - 8:cy + 9, cx - 8:cx + 9, 0] == 0).all() and f[cy, cx, 1] >= 240
    assert (f[700 - 8:700 + 9, 1000 - 8:1000 + 9, 0] == 0).all()
    # rows are generated independently of the band they are requested in (row-sharded ranks agree)
    band = cp.synth_frame_host(1, 42, 3, 25, 1024, 768, 3, row0=200, rows=50)
    assert np.array_equal(band, f[200:250])


def test_sample_positions_are_sorted_distinct():
    p = cp.sample_positions(9, 200, 40)
    assert len(set(p.tolist())) == 40 and (np.diff(p) > 0).all() and p.min() >= 0 and p.max() < 200
    assert np.array_equal(p, cp.sample_positions(9, 200, 40))


def test_shake_option_parsing_mirrors_the_reference():
    # src/shake.rs:58-119
    p = cp.ShakeParams.from_str("20/10")
    assert (p.anchor_radius, p.search_radius) == (20, 10)
    assert cp.ShakeAnchor.from_str("120/-3").anchor == (120, -3)
    for bad, cls, msg in [("20", cp.ShakeParams, "expected <rad>/<search-rad>"), ("a/3", cp.ShakeParams, "Unexpected format in shake parameter"),
                          ("1/2/3", cp.ShakeAnchor, "expected x/y"), ("x/2", cp.ShakeAnchor, "expected x/y")]:
        with pytest.raises(cp.ParseOptionError) as e:
            cls.from_str(bad)
        assert msg in str(e.value)
