"""Scratch: repeats one compositing call on a resident full-size stack and compares every result with the first one (any
difference is a race) and rows of it with the oracle. python tools/determinism.py kind reps inline_min"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as orc
import chrono_photo_b200 as cp
kind, reps, inline_min = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
H, W, N = 4000, 6000, 200
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(kind, 42)
cp.set_tuning("inline_min", inline_min)
proc = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
ref_img, ref_msk = proc.process(fs)
ref_img, ref_msk = ref_img.copy(), ref_msk.copy()
bad_runs = 0
for i in range(reps):
    for _ in range(3):
        proc.enqueue_device(fs)
    img, msk = proc.process(fs)
    d = (img != ref_img).any(axis=2) | (msk != ref_msk).any(axis=2)
    if d.any():
        bad_runs += 1
        ys, xs = np.nonzero(d)
        tiles = sorted(set(((ys * W + xs) // 32).tolist()))
        print(f"run {i}: {int(d.sum())} pixels differ from run 0, tiles {tiles[:8]} ({len(tiles)} tiles), first {list(zip(ys[:4].tolist(), xs[:4].tolist()))}")
print(f"kind {kind} inline_min {inline_min}: {bad_runs} of {reps} runs differ from the first")
# which is right? rows of the first result against the oracle
thr = orc.threshold(True, 0.05, 0.2)
for r0 in (0, 742, 1992, 3984):
    st = orc.synth_frames(kind, 42, N, W, H, rows=16, row0=r0)
    oimg, omsk, _ = orc.outlier(st, thr, 0, 2, n_threads=os.cpu_count())
    print("rows", r0, "first run vs oracle:", int(((oimg != ref_img[r0:r0 + 16]).any(axis=2) | (omsk != ref_msk[r0:r0 + 16]).any(axis=2)).sum()))
