"""The C++ host mirror (include/chrono_b200.hpp) and the CLI-compatible driver tools/chrono_b200_cli (PPM frames)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as orc
from test_oracle import BG, OM, make_stack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tools", "chrono_b200_cli")


CONVERT = os.path.join(ROOT, "tools", "imageio_convert")
HEADERS = [os.path.join(ROOT, "include", h) for h in ("chrono_b200.hpp", "chrono_b200_imageio.hpp", "chrono_b200.h")]


def _stale(binary, src):
    return not os.path.exists(binary) or os.path.getmtime(binary) < max(os.path.getmtime(d) for d in [src] + HEADERS)


def build_cli():
    src = os.path.join(ROOT, "tools", "chrono_b200_cli.cpp")
    if _stale(CLI, src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", CLI, src, "-L" + os.path.join(ROOT, "chrono_photo_b200"), "-lchrono_b200", "-lz",
                               "-Wl,-rpath,$ORIGIN/../chrono_photo_b200"])


def build_convert():
    src = os.path.join(ROOT, "tools", "imageio_convert.cpp")
    if _stale(CONVERT, src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", CONVERT, src, "-L" + os.path.join(ROOT, "chrono_photo_b200"), "-lchrono_b200", "-lz",
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


def test_driver_image_files_round_trip_without_gpu(tmp_path):
    """The driver's frame reader / save_image (include/chrono_b200_imageio.hpp; src/main.rs:520-571) against Pillow: PNG with every
    row filter and RGBA in, PNG / TIFF / BMP / PPM out, byte for byte."""
    from PIL import Image
    build_convert()
    rng = np.random.default_rng(5)
    for c, mode in ((3, "RGB"), (4, "RGBA")):
        img = rng.integers(0, 256, size=(37, 53, c), dtype=np.uint8)
        img[5:20, 3:40] = (np.arange(37)[:, None, None] * 3 + np.arange(c)[None, None, :] * 40).astype(np.uint8)[:15]  # smooth area: Paeth / Avg rows
        src = tmp_path / f"in{c}.png"
        Image.fromarray(img, mode).save(src, optimize=True)  # Pillow chooses filters adaptively per row
        for ext in ("png", "tif", "bmp") + (("ppm",) if c == 3 else ()):
            out = tmp_path / f"out{c}.{ext}"
            p = subprocess.run([CONVERT, str(src), str(out)], capture_output=True, text=True, timeout=60)
            assert p.returncode == 0, p.stderr
            assert p.stdout.split() == ["53", "37", str(c)]
            back = np.asarray(Image.open(out).convert(mode))
            if ext == "bmp" and c == 4:  # a 32-bit BMP with the plain 40-byte header: readers treat the fourth byte as padding
                back, want = back[:, :, :3], img[:, :, :3]
            else:
                want = img
            assert np.array_equal(back, want), (c, ext)
    # error behaviour: 16-bit and grey frames are "Not an 8 bit image" for the path (as_flat_samples_u8 + Rgb8/Rgba8 only)
    Image.fromarray(rng.integers(0, 256, size=(8, 8), dtype=np.uint8), "L").save(tmp_path / "grey.png")
    p = subprocess.run([CONVERT, str(tmp_path / "grey.png"), str(tmp_path / "x.png")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "Not an 8 bit image" in p.stderr
    p = subprocess.run([CONVERT, str(tmp_path / "in3.png"), str(tmp_path / "noext")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "Expects an extension" in p.stderr


@pytest.mark.gpu
def test_cli_shake_png_frames_jpeg_output(tmp_path):
    """--shake / --shake-anchors (src/main.rs:61-84): PNG frames shifted by known offsets are analysed on the GPU, cropped by their
    Crop origins during the upload and composited; the result equals the oracle on the cropped frames. JPEG output decodes to
    the same image within the codec's tolerance; JPEG frames go through nvJPEG ingest."""
    from PIL import Image
    import io
    build_cli()
    rng = np.random.default_rng(17)
    n, h, w = 9, 48, 64
    base = rng.integers(0, 256, size=(h + 16, w + 16, 3), dtype=np.uint8)
    offs = [(0, 0)] + [(int(rng.integers(-3, 4)), int(rng.integers(-3, 4))) for _ in range(n - 1)]
    frames = np.stack([base[8 - oy:8 - oy + h, 8 - ox:8 - ox + w] for ox, oy in offs]).copy()
    for i in range(n):  # a moving object on top of the static scene
        frames[i, 20:26, 5 + 5 * i:11 + 5 * i] = 255 - frames[i, 20:26, 5 + 5 * i:11 + 5 * i]
    for i, fr in enumerate(frames):
        Image.fromarray(fr, "RGB").save(tmp_path / f"f-{i:03d}.png")
    anchors = [(16, 12), (48, 36)]
    got_offs = orc.shake_analyze(frames, anchors, 5, 4)
    out, blend = str(tmp_path / "out.png"), str(tmp_path / "blend.png")
    p = run("--pattern", str(tmp_path / "f-*.png"), "--output", out, "--output-blend", blend, "--shake", "5/4", "--shake-anchors", "16/12", "48/36",
            "-b", "first", "-l", "extreme", "--shake-threads", "2")
    assert p.returncode == 0, p.stderr
    assert "Camera shake detected" in p.stdout
    xs, ys = [o[0] for o in got_offs], [o[1] for o in got_offs]
    xmin, xmax, ymin, ymax = min(0, min(xs)), max(0, max(xs)), min(0, min(ys)), max(0, max(ys))
    cw, ch = w + xmin - xmax, h + ymin - ymax  # Crop::create, src/shake.rs:136-176
    cropped = np.stack([frames[i, -ymin + ys[i]:-ymin + ys[i] + ch, -xmin + xs[i]:-xmin + xs[i] + cw] for i in range(n)])
    oimg, omsk, _ = orc.outlier(cropped, orc.threshold(True, 0.05, 0.2), BG["first"], OM["extreme"])
    assert np.array_equal(np.asarray(Image.open(out)), oimg) and np.array_equal(np.asarray(Image.open(blend)), omsk)
    # JPEG output (nvJPEG encode at --quality) and JPEG frames in (nvJPEG decode on the device), on a smooth scene
    smooth = np.stack([np.clip(np.add.outer(np.arange(h) * 2, np.arange(w))[:, :, None] + np.array([0, 30, 60]) + 3 * i, 0, 255).astype(np.uint8) for i in range(5)])
    for i, fr in enumerate(smooth):
        Image.fromarray(fr, "RGB").save(tmp_path / f"s-{i:03d}.png")
        Image.fromarray(fr, "RGB").save(tmp_path / f"j-{i:03d}.jpg", quality=95, subsampling=0)
    p = run("--pattern", str(tmp_path / "s-*.png"), "--output", str(tmp_path / "dark.jpg"), "--mode", "darker", "--quality", "100")
    assert p.returncode == 0, p.stderr
    dj = np.asarray(Image.open(tmp_path / "dark.jpg").convert("RGB")).astype(int)
    assert dj.shape == (h, w, 3) and np.abs(dj - orc.simple(smooth, True).astype(int)).mean() < 3.0  # lossy codec, 4:2:0 chroma
    p = run("--pattern", str(tmp_path / "j-*.jpg"), "--output", str(tmp_path / "light.png"), "--mode", "lighter")
    assert p.returncode == 0, p.stderr
    decoded = np.stack([np.asarray(Image.open(tmp_path / f"j-{i:03d}.jpg").convert("RGB")) for i in range(5)])
    lj = np.asarray(Image.open(tmp_path / "light.png")).astype(int)
    assert np.abs(lj - orc.simple(decoded, False).astype(int)).max() <= 3  # decoders differ in the last bits (IDCT rounding)


@pytest.mark.gpu
def test_cli_row_bands_equal_the_whole_image(tmp_path):
    """Out of core (the reference's time slices, src/slicer.rs:19-41): a series that does not fit the device is composited in row
    bands, each a stack of its own. Forced here with CHRONO_B200_BAND_ROWS: same files as the one-stack run, including
    --background random, whose per-pixel draws are keyed by the global pixel index (pixel_offset)."""
    build_cli()
    rng = np.random.default_rng(23)
    st = make_stack(rng, 21, 29, 44, 3, n_obj=30)
    for i, fr in enumerate(st):
        write_ppm(tmp_path / f"image-{i:05d}.ppm", fr)
    pat = str(tmp_path / "image-*.ppm")
    for extra in (["-b", "random", "-l", "extreme"], ["-b", "first", "-l", "backward", "-t", "rel/2/4"], ["--mode", "lighter"]):
        whole, banded = str(tmp_path / "whole.ppm"), str(tmp_path / "banded.ppm")
        wb, bb = str(tmp_path / "whole_b.ppm"), str(tmp_path / "banded_b.ppm")
        blend = [] if "--mode" in extra else ["--output-blend"]
        p = run("--pattern", pat, "--output", whole, *(blend + [wb] if blend else []), *extra)
        assert p.returncode == 0, p.stderr
        env = dict(os.environ, CHRONO_B200_BAND_ROWS="7")
        p = subprocess.run([CLI, "--pattern", pat, "--output", banded, *(blend + [bb] if blend else []), *extra], capture_output=True, text=True, timeout=120, env=env)
        assert p.returncode == 0, p.stderr
        assert "row bands of 7 rows" in p.stdout
        assert np.array_equal(read_ppm(whole), read_ppm(banded)), extra
        if blend:
            assert np.array_equal(read_ppm(wb), read_ppm(bb)), extra
    # a video needs the whole clip resident
    env = dict(os.environ, CHRONO_B200_BAND_ROWS="7")
    p = subprocess.run([CLI, "--pattern", pat, "--output", str(tmp_path / "v.ppm"), "--video-in", "0/5/1"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 1 and "whole clip resident" in p.stderr
