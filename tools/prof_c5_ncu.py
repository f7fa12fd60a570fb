import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
H, W, N = 1064, 1904, 300
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(2, 42)
p = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
for _ in range(2): print(p.process_video_run_device(fs, 96, 25, 128))
