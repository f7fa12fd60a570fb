"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every tier of K1 (streaming, queue compaction,
solver and histogram tiers, per-frame path), a chrono-video run, both K2 kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import chrono_photo_b200 as cp
from test_oracle import make_stack
rng = np.random.default_rng(1)
ctx = cp.Context([0])
for n, c in ((25, 3), (200, 3), (70, 4), (300, 3)):
    st = make_stack(rng, n, 6, 70, c, noise=5, n_obj=30)
    fs = cp.FrameStack(ctx, 70, 6, c, n)
    fs.upload_all(st)
    for thr in (cp.Threshold.abs(0.05, 0.2), cp.Threshold.rel(3.0, 5.0)):
        for bg, om in ((0, 2), (1, 4), (2, 3), (3, 5)):
            cp.OutlierProcessor(thr, bg, om, seed=3).process(fs)
            cp.OutlierProcessor(thr, bg, om, seed=3).process(fs, list(range(1, n - 1, 2)))
    cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2, sample_count=max(3, n // 3)).process(fs)
    if n >= 70:  # chrono-video runs (video_kernel + video_exact_kernel): 20 windows of 9 frames, 37 of 25 (relative thresholds too)
        cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2).process_video_run(fs, 3, 9, 20)
        cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 1, 4, seed=3).process_video_run(fs, 5, 25, 37)
        cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 2, 3).process_video_run(fs, 0, 24, 18)
    # an interleaved row-block shard (blocks of 2 rows, 3 ranks): global pixel indices for the per-pixel draws
    cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 1, 2, seed=3, pixel_offset=2 * 70, block_pixels=2 * 70, block_skip=4 * 70).process(fs)
    cp.SimpleProcessor(darker=True).process(fs)
    cp.SimpleProcessor((1, 0.5, 0.25, 0), cp.Fade(0, False, [(0, 1.0), (9, 0.0)]), False).process(fs, list(range(0, n, 3)))
    fs.close()
# noise at the threshold / iid bytes: tiles classified frame by frame inside the streaming kernel and in the iterative tier (dense pass)
for kind in (4, 3):
    fs = cp.FrameStack(ctx, 256, 8, 3, 200)
    fs.fill_synthetic(kind, 42)
    for bg, om in ((0, 2), (1, 4), (2, 3), (3, 5), (0, 0), (0, 1)):
        cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), bg, om, seed=3).process(fs)
    cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2).process(fs, list(range(5, 190)))
    fs.close()
print("sanitize run done")
