"""Scratch timing of the kernels on synthetic stacks (not the bench): python tools/quick_time.py [H] [W] [N] [kind]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib

H = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200
kinds = [int(k) for k in sys.argv[4].split(",")] if len(sys.argv) > 4 else [2]
ctx = cp.Context([0])
t0 = time.time()
fs = cp.FrameStack(ctx, W, H, 3, N)
print(f"stack {N}x{H}x{W}x3 = {fs.device_bytes()/1e9:.2f} GB, create {time.time()-t0:.2f}s")
for kind in kinds:
    t0 = time.time()
    fs.fill_synthetic(kind, 42)
    print(f"kind {kind}: fill {time.time()-t0:.2f}s")
    alg = H * W * 3 * (N + 2)
    for name, proc in [("outlier abs first/extreme", cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2)),
                       ("outlier abs median/extreme", cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 3, 2)),
                       ("outlier rel first/forward", cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), 0, 4)),
                       ("darker", cp.SimpleProcessor(darker=True)), ("lighter", cp.SimpleProcessor(darker=False))]:
        ms = [proc.process_device(fs) for _ in range(4)]
        best = min(ms[1:])
        slow = _lib.lib().chb_last_slow_pixels() if "outlier" in name else 0
        hard = _lib.lib().chb_last_hard_pixels() if "outlier" in name else 0
        print(f"  {name:28s} ms {['%.3f' % m for m in ms]}  {alg/best/1e6:8.0f} GB/s  {N*H*W/best/1e6:9.1f} Gpf/s  slow px {slow} ({100*slow/(H*W):.3f}%) hard {100*hard/(H*W):.2f}%")
fs.close()
