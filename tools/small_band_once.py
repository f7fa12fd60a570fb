"""Scratch: a few calls on one eighth of config 3 (rows [3000, 3500): the band with most disc pixels) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
ctx = cp.Context([0])
fs = cp.FrameStack(ctx, 6000, 500, 3, 200); fs.fill_synthetic(2, 42, 3000, 4000)
p = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)
for _ in range(4): print(p.process_device(fs))
