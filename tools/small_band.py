"""Scratch: kernel time of small row bands (what one rank of a strong-scaling run owns)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
ctx = cp.Context([0])
for rows in (4000, 2000, 1000, 500, 250):
    fs = cp.FrameStack(ctx, 6000, rows, 3, 200); fs.fill_synthetic(2, 42, 0, 4000)
    p = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
    ms = [p.process_device(fs) for _ in range(6)]
    t0 = time.perf_counter()
    for _ in range(20): p.enqueue_device(fs)
    fs.wait(); dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"rows {rows}: launch ms {min(ms[1:]):.4f} (ideal {3.10*rows/4000:.4f})  enqueue-loop ms/step {dt:.4f}")
    fs.close()
