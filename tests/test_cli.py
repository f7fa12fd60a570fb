"""The C++ host mirror (include/chrono_b200.hpp) and the CLI-compatible driver tools/chrono_b200_cli (PPM frames)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as orc
from test_oracle import BG, OM, make_stack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tools", "chrono_b200_cli")


def build_cli():
    src = os.path.join(ROOT, "tools", "chrono_b200_cli.cpp")
    if not os.path.exists(CLI) or os.path.getmtime(CLI) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", CLI, src, "-L" + os.path.join(ROOT, "chrono_photo_b200"), "-lchrono_b200",
                               "-Wl,-rpath,$ORIGIN/../chrono_photo_b200"])


def write_ppm(path, img):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def read_ppm(path):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=120)


def test_cli_argument_errors_need_no_gpu(tmp_path):
    build_cli()
    assert run("--output", "x.ppm").returncode == 1  # --pattern is required (src/cli.rs:23-24)
    p = run("--pattern", str(tmp_path / "*.ppm"), "--output", str(tmp_path / "o.ppm"), "--mode", "brightest")
    assert p.returncode == 1 and "Not a pixel selection mode" in p.stderr
    p = run("--pattern", str(tmp_path / "*.ppm"), "--output", str(tmp_path / "o.ppm"), "--threshold", "foo/1")
    assert p.returncode == 1 and "Not a pixel outlier detection mode" in p.stderr
    p = run("--pattern", str(tmp_path / "*.ppm"), "--output", str(tmp_path / "o.ppm"))
    assert p.returncode == 1 and "Unable to process search pattern" in p.stderr


@pytest.mark.gpu
def test_cli_photo_video_and_option_file(tmp_path):
    build_cli()
    rng = np.random.default_rng(91)
    st = make_stack(rng, 14, 20, 28, 3, n_obj=25)
    for i, fr in enumerate(st):
        write_ppm(tmp_path / f"image-{i:05d}.ppm", fr)
    pat = str(tmp_path / "image-*.ppm")
    # outlier photo with explicit policies + blend mask
    out, blend = str(tmp_path / "out.ppm"), str(tmp_path / "blend.ppm")
    p = run("--pattern", pat, "--output", out, "--output-blend", blend, "-t", "abs/0.05/0.2", "-b", "first", "-l", "forward", "--slice", "rows/4")
    assert p.returncode == 0, p.stderr
    oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG["first"], OM["forward"])
    assert np.array_equal(read_ppm(out), oimg) and np.array_equal(read_ppm(blend), omsk)
    # darker with weights, --frames start/end/step, through an option file with a quoted pattern (src/main.rs:33-43)
    optf = tmp_path / "run.chrono"
    optf.write_text(f'--pattern "{pat}"\n--output {tmp_path}/dark.ppm\n--mode darker --weights 1 0.5 0.25 0\n--frames 2/12/3 --threshold abs/0.1\n')
    p = run(str(optf))
    assert p.returncode == 0, p.stderr
    assert "not used" in p.stdout and "--threshold" in p.stdout  # unused-flag warning (src/cli.rs:233-239)
    sel = st[2:12][::3]
    assert np.array_equal(read_ppm(tmp_path / "dark.ppm"), orc.simple(sel, True, weights=(1, 0.5, 0.25, 0)))
    # video: --video-in -4/1/1 -> one output frame per window, numbered from v_lower (src/main.rs:230-331)
    p = run("--pattern", pat, "--output", str(tmp_path / "vid.ppm"), "--mode", "lighter", "--video-in", "-4/1/1", "--video-out", "2/9/2")
    assert p.returncode == 0, p.stderr
    n, ws, we, num = orc.video_windows(14, (-4, 1, 1), (2, 9, 2))
    for i in range(n):
        if ws[i] < we[i]:
            got = read_ppm(tmp_path / f"vid-{num[i]:05d}.ppm")
            assert np.array_equal(got, orc.simple(st, False, indices=list(range(ws[i], we[i]))))
    # outlier video: --video-in 0/5/1 -> a growing head (single windows), then a run of 5-frame windows composited by the
    # sliding-window kernel in one call (chb_outlier_video); every frame and blend mask equals the oracle's
    p = run("--pattern", pat, "--output", str(tmp_path / "ov.ppm"), "--output-blend", str(tmp_path / "ob.ppm"), "--video-in", "0/5/1",
            "-t", "abs/0.05/0.2", "-b", "first", "-l", "extreme")
    assert p.returncode == 0, p.stderr
    n, ws, we, num = orc.video_windows(14, (0, 5, 1), (None, None, 1))
    checked = 0
    for i in range(n):
        if ws[i] < we[i]:
            idx = list(range(ws[i], we[i]))
            oimg, omsk, _ = orc.outlier(st, orc.threshold(True, 0.05, 0.2), BG["first"], OM["extreme"], indices=idx)
            assert np.array_equal(read_ppm(tmp_path / f"ov-{num[i]:05d}.ppm"), oimg), idx
            assert np.array_equal(read_ppm(tmp_path / f"ob-{num[i]:05d}.ppm"), omsk), idx
            checked += 1
    assert checked >= 10
