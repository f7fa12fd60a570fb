"""Worker of tests/test_dist_cpu.py: world_size-2 gloo run of the row-shard + band-gather host logic. The per-band
compute is done by the oracle here (CPU test infrastructure); on a GPU box bench.py runs the same logic over NCCL with
the CUDA path."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as orc  # noqa: E402
from chrono_photo_b200.sharding import gather_bands, max_over_ranks, shard_rows  # noqa: E402
from chrono_photo_b200 import synth_frame_host  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, H, W = 12, 37, 29  # H not divisible by the world size on purpose
    row0, rows = shard_rows(H, rank, world)
    band = np.stack([synth_frame_host(1, 7, f, n, W, H, 3, row0=row0, rows=rows) for f in range(n)])
    thr = orc.threshold(True, 0.05, 0.2)
    # background random: the per-pixel RNG is keyed by the GLOBAL pixel index, so shards must agree with the full image
    img, msk, warn = orc.outlier(band, thr, 1, 2, seed=11, pixel_offset=row0 * W)
    dark = orc.simple(band, True)
    full_img = gather_bands(img, H, dist)
    full_msk = gather_bands(msk, H, dist)
    full_dark = gather_bands(dark, H, dist)
    slowest = max_over_ranks(1.0 + rank, dist)
    assert slowest == float(world), slowest
    if rank == 0:
        st = np.stack([synth_frame_host(1, 7, f, n, W, H, 3) for f in range(n)])
        ref_img, ref_msk, _ = orc.outlier(st, thr, 1, 2, seed=11)
        assert np.array_equal(full_img, ref_img) and np.array_equal(full_msk, ref_msk)
        assert np.array_equal(full_dark, orc.simple(st, True))
        print("DIST_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
