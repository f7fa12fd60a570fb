"""Scratch: a few launches of one workload (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
H, W, N, rel = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kind = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ctx = cp.Context([0]); fs = cp.FrameStack(ctx, W, H, 3, N); fs.fill_synthetic(kind, 42)
p = cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 0, 4) if rel else cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
for _ in range(3): print(p.process_device(fs))
