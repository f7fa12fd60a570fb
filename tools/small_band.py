"""Scratch: kernel time of the shards one rank of a strong-scaling run owns -- contiguous bands and interleaved row blocks
(sharding.InterleavedShard): per-call device time, streaming kernel alone, and the enqueue loop's time per step (launch
overhead shows there). `hard_drains` 0 / 1: with / without the exact-path launch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chrono_photo_b200 as cp
from chrono_photo_b200 import _lib
from chrono_photo_b200.sharding import InterleavedShard, interleave_block_rows
ctx = cp.Context([0])
def run(rows, row0, shard=None, tag=""):
    fs = cp.FrameStack(ctx, 6000, rows, 3, 200)
    if shard: fs.fill_synthetic(2, 42, **shard.fill_args())
    else: fs.fill_synthetic(2, 42, row0, 4000)
    p = cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), 0, 2, **(shard.processor_args() if shard else {}))
    ms, main = [], []
    for _ in range(6):
        ms.append(p.process_device(fs)); main.append(float(_lib.lib().chb_last_main_kernel_ms()))
    t0 = time.perf_counter()
    for _ in range(20): p.enqueue_device(fs)
    fs.wait(); dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{tag}rows {rows} at {row0}: call ms {min(ms[1:]):.4f} main {min(main[1:]):.4f} tiers {min(ms[1:]) - min(main[1:]):.4f} (ideal call {2.455*rows/4000:.4f})  enqueue-loop ms/step {dt:.4f}  slow {int(_lib.lib().chb_last_slow_pixels())}", flush=True)
    fs.close()
for hd, wr in ((1, 1), (1, 0)):
    _lib.lib().chb_set_tuning(b"hard_drains", hd)
    _lib.lib().chb_set_tuning(b"hard_window", wr)
    print(f"== hard_drains {hd} hard_window {wr}")
    for rows in (4000, 2000, 1000):
        run(rows, 0)
    for k in (3, 6, 7):
        run(500, 500 * k)
    for G in (2, 4, 8):
        B = interleave_block_rows(4000, G)
        for g in ((0, G - 1) if G < 8 else (0, 3, 7)):
            sh = InterleavedShard(4000, 6000, g, G, B)
            run(sh.rows, sh.row0, sh, tag=f"interleaved G={G} g={g} B={B}: ")
