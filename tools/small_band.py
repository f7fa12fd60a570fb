"""Scratch: kernel time of small row bands (what one rank of a strong-scaling run owns): per-call device time, streaming kernel
alone, and the enqueue loop's time per step (launch overhead shows there)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib
ctx = cp.Context([0])
def run(rows, row0):
    fs = cp.FrameStack(ctx, 6000, rows, 3, 200); fs.fill_synthetic(2, 42, row0, 4000)
    p = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
    ms, main = [], []
    for _ in range(6):
        ms.append(p.process_device(fs)); main.append(float(_lib.lib().chb_last_main_kernel_ms()))
    t0 = time.perf_counter()
    for _ in range(20): p.enqueue_device(fs)
    fs.wait(); dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"rows {rows} at {row0}: call ms {min(ms[1:]):.4f} main {min(main[1:]):.4f} tiers {min(ms[1:]) - min(main[1:]):.4f} (ideal call {2.455*rows/4000:.4f})  enqueue-loop ms/step {dt:.4f}  slow {int(_lib.lib().chb_last_slow_pixels())}")
    fs.close()
for rows in (4000, 2000, 1000):
    run(rows, 0)
for k in range(8):
    run(500, 500 * k)
