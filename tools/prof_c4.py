import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
H, W, N = 540, 3840, 1000
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(2, 42, 800, 2160)
p = cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 0, 4)
for _ in range(3): print(p.process_device(fs))
