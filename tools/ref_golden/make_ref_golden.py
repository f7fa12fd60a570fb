"""Produces tests/golden_ref/*.npz: outputs of the UNMODIFIED reference (mlange-42/chrono-photo, its own CLI binary) over
the committed golden stacks (tests/golden/*.npz). Needs a box with cargo (this image has none -- no cargo, no rustc, no
network -- which is why the repository ships without these files and tests/test_golden_ref.py reports "parity unpinned").

    cargo build --release --manifest-path /path/to/chrono-photo/Cargo.toml
    python tools/ref_golden/make_ref_golden.py --bin /path/to/chrono-photo/target/release/chrono-photo

For every case the stack's frames are written as lossless PNG files, the reference runs through its public command line
(src/cli.rs) with the case's flags, and its PNG outputs (--output, --output-blend) are read back. Only deterministic
option sets are used: `--background random` and `--sample` draw from an OS-seeded RNG in the reference
(src/chrono.rs:68,157,357,553) and cannot be pinned by any vector.
Commit the resulting tests/golden_ref/*.npz; tests/test_golden_ref.py then checks the oracle (CPU) and the CUDA path (GPU)
against them bit for bit.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from pngio import read_png, write_png  # noqa: E402

# name -> (golden stack file, frame indices or None, reference command-line flags, outputs wanted)
CASES = {
    "outlier_abs_first_extreme": ("outlier_abs_first_extreme.npz", None,
                                  ["--mode", "outlier", "--threshold", "abs/0.05/0.2", "--background", "first", "--outlier", "extreme"], ("image", "mask")),
    "outlier_abs_first_forward": ("outlier_abs_random_forward.npz", None,
                                  ["--mode", "outlier", "--threshold", "abs/0.05/0.2", "--background", "first", "--outlier", "forward"], ("image", "mask")),
    "outlier_abs_median_last": ("outlier_abs_random_forward.npz", None,
                                ["--mode", "outlier", "--threshold", "abs/0.05/0.2", "--background", "median", "--outlier", "last"], ("image", "mask")),
    "outlier_abs_average_average_rgba": ("outlier_abs_average_average_rgba.npz", None,
                                         ["--mode", "outlier", "--threshold", "abs/0.04/0.15", "--background", "average", "--outlier", "average"], ("image", "mask")),
    "outlier_rel_median_backward": ("outlier_rel_median_backward.npz", None,
                                    ["--mode", "outlier", "--threshold", "rel/3.0/5.0", "--background", "median", "--outlier", "backward",
                                     "--weights", "1", "1", "1", "0"], ("image", "mask")),
    "outlier_rel_first_last_window_fade": ("outlier_rel_first_last_window_fade.npz", list(range(4, 37, 3)),
                                           ["--mode", "outlier", "--threshold", "rel/2.5/4.0", "--background", "first", "--outlier", "last",
                                            "--weights", "1", "0.5", "0.5", "0", "--fade", "repeat/rel/(0,1.0)/(6,0.0)/(9,0.5)"], ("image", "mask")),
    "outlier_abs_first_first_fade": ("outlier_abs_first_first_sample.npz", None,
                                     ["--mode", "outlier", "--threshold", "abs/0.05/0.2", "--background", "first", "--outlier", "first",
                                      "--fade", "clamp/abs/(0,0.0)/(10,1.0)"], ("image", "mask")),
    "simple_darker": ("simple.npz", None, ["--mode", "darker"], ("image",)),
    "simple_lighter": ("simple.npz", None, ["--mode", "lighter"], ("image",)),
    "simple_darker_weighted": ("simple.npz", None, ["--mode", "darker", "--weights", "1", "0.5", "0.25", "0"], ("image",)),
    "simple_lighter_fade_window": ("simple.npz", list(range(2, 20, 2)), ["--mode", "lighter", "--fade", "clamp/rel/(0,1.0)/(8,0.0)"], ("image",)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bin", required=True, help="path of the reference's chrono-photo binary (cargo build --release)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden_ref"))
    ap.add_argument("--keep", action="store_true", help="keep the temporary frame directories")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name, (stack_file, indices, flags, wanted) in CASES.items():
        stack = np.load(os.path.join(ROOT, "tests", "golden", stack_file))["stack"]
        frames = stack if indices is None else stack[indices]
        tmp = tempfile.mkdtemp(prefix="chrono_ref_")
        for i, fr in enumerate(frames):
            write_png(os.path.join(tmp, f"frame-{i:05d}.png"), fr)
        out_png, blend_png = os.path.join(tmp, "out.png"), os.path.join(tmp, "blend.png")
        cmd = [args.bin, "--pattern", os.path.join(tmp, "frame-*.png"), "--output", out_png, "--temp-dir", tmp] + flags
        if "mask" in wanted:
            cmd += ["--output-blend", blend_png]
        print(" ".join(cmd))
        subprocess.run(cmd, check=True, stdin=subprocess.DEVNULL)
        res = {"stack_file": stack_file, "indices": np.array(indices if indices is not None else [], np.int32), "flags": np.array(flags),
               "image": read_png(out_png)}
        if "mask" in wanted:
            res["mask"] = read_png(blend_png)
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **res)
        if not args.keep:
            shutil.rmtree(tmp, ignore_errors=True)
    print("wrote", len(CASES), "reference vectors to", args.out)


if __name__ == "__main__":
    main()
