"""Generates tests/golden/*.npz: small seeded frame stacks with the oracle's outputs for a spread of options.

The reference cannot be built in this environment (no cargo/rustc) and ships no image fixtures, so these vectors are
produced by the CPU oracle (oracle/chrono_oracle.c), itself pinned by the reference's known-answer vectors and by the
independent numpy restatement (tests/test_oracle.py). They freeze today's agreed answers: any later change of the
oracle, the kernels or the host code that alters a byte shows up as a diff against a committed file.

    python tests/golden/make_golden.py        # rewrites the fixtures (review the diff before committing)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as orc  # noqa: E402
from test_oracle import BG, OM, make_stack  # noqa: E402

CASES = [
    # name, frames, h, w, c, threshold, bg, om, weights, fade, indices, sample_pos
    ("abs_first_extreme", 25, 12, 16, 3, (True, 0.05, 0.2), "first", "extreme", (1, 1, 1, 1), None, None, None),
    ("abs_random_forward", 25, 12, 16, 3, (True, 0.05, 0.2), "random", "forward", (1, 1, 1, 1), None, None, None),
    ("abs_average_average_rgba", 18, 8, 12, 4, (True, 0.04, 0.15), "average", "average", (1, 1, 1, 1), None, None, None),
    ("rel_median_backward", 31, 10, 14, 3, (False, 3.0, 5.0), "median", "backward", (1, 1, 1, 0), None, None, None),
    ("rel_first_last_window_fade", 40, 8, 12, 3, (False, 2.5, 4.0), "first", "last", (1, 0.5, 0.5, 0), (1, False, [(0, 1.0), (6, 0.0), (9, 0.5)]),
     list(range(4, 37, 3)), None),
    ("abs_first_first_sample", 30, 8, 12, 3, (True, 0.05, 0.2), "first", "first", (1, 1, 1, 1), (0, True, [(0, 0.0), (10, 1.0)]), None,
     [0, 3, 4, 9, 11, 17, 20, 21, 28]),
]


def main():
    for i, (name, n, h, w, c, thr, bg, om, wts, fade, idx, spos) in enumerate(CASES):
        rng = np.random.default_rng(1000 + i)
        st = make_stack(rng, n, h, w, c, noise=5, n_obj=12)
        f = orc.fade(*fade) if fade else None
        img, msk, warn, dbg = orc.outlier(st, orc.threshold(*thr), BG[bg], OM[om], wts, f, idx, spos, seed=77, want_debug=True)
        np.savez_compressed(os.path.join(HERE, f"outlier_{name}.npz"), stack=st, image=img, mask=msk, warnings=np.int64(warn),
                            median=dbg["median"], q1=dbg["q1"], q3=dbg["q3"], n_outliers=dbg["n_outliers"])
    rng = np.random.default_rng(2000)
    st = make_stack(rng, 21, 10, 14, 3, n_obj=10)
    fade = (0, False, [(0, 1.0), (8, 0.0)])
    np.savez_compressed(os.path.join(HERE, "simple.npz"), stack=st, darker=orc.simple(st, True), lighter=orc.simple(st, False),
                        darker_weighted=orc.simple(st, True, weights=(1, 0.5, 0.25, 0)),
                        lighter_fade_window=orc.simple(st, False, fade_=orc.fade(*fade), indices=list(range(2, 20, 2))))
    print("wrote", len(CASES) + 1, "fixtures to", HERE)


if __name__ == "__main__":
    main()
