"""Scratch: chrono-video on the c5 clip: per-window launches vs the sliding-window run kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib
H, W, N = 1064, 1904, int(sys.argv[1]) if len(sys.argv) > 1 else 1800
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 400
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(2, 42)
L = _lib.lib()
for name, p, wl in [("abs 25", cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2), 25), ("rel 25", cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 0, 4), 25),
                    ("abs 50", cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2), 50), ("abs 10", cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2), 10)]:
    idx = list(range(100, 100 + wl))
    ms1 = min(p.process_device(fs, idx) for _ in range(3))
    s1, h1 = L.chb_last_slow_pixels(), L.chb_last_hard_pixels()
    cnt = min(nwin, N - wl - 100)
    msr = min(p.process_video_run_device(fs, 100, wl, cnt) for _ in range(2))
    print(f"{name}: per-window launch {ms1:.4f} ms (slow {100*s1/(H*W):.2f}% hard {100*h1/(H*W):.2f}%) | run of {cnt}: {msr:.2f} ms = {msr/cnt:.4f} ms/window"
          f" (slow {100*L.chb_last_slow_pixels()/(H*W*cnt):.3f}% hard band {100*L.chb_last_hard_pixels()/(H*W*cnt):.3f}%)"
          f" alg GB/s {H*W*3*(wl+2)*cnt/msr/1e6:.0f}")
