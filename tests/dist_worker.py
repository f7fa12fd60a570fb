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
    # ---- strong-scaling shards: interleaved row blocks (rank g owns the blocks with index = g mod G). The oracle composites a
    # shard block by block (pixel_offset = the block's first global pixel), which is what block_pixels / block_skip make the CUDA
    # path do in one call; the gathered shards are de-interleaved on rank 0.
    from chrono_photo_b200.sharding import InterleavedShard, deinterleave, interleave_block_rows
    H2 = 36
    B = interleave_block_rows(H2, world, 4)
    sh = InterleavedShard(H2, W, rank, world, B)
    rows_g = sh.global_rows()
    assert sh.processor_args()["pixel_offset"] == rows_g[0] * W and len(rows_g) == H2 // world
    full2 = np.stack([synth_frame_host(1, 7, f, n, W, H2, 3) for f in range(n)])
    mine = full2[:, rows_g]
    parts_i, parts_m = [], []
    for b0 in range(0, sh.rows, B):
        bi, bm, _ = orc.outlier(np.ascontiguousarray(mine[:, b0:b0 + B]), thr, 1, 2, seed=11, pixel_offset=int(rows_g[b0]) * W)
        parts_i.append(bi); parts_m.append(bm)
    import torch
    mine_img = torch.from_numpy(np.concatenate(parts_i))
    mine_msk = torch.from_numpy(np.concatenate(parts_m))
    g_i = [torch.empty_like(mine_img) for _ in range(world)] if rank == 0 else None
    g_m = [torch.empty_like(mine_msk) for _ in range(world)] if rank == 0 else None
    dist.gather(mine_img, g_i, dst=0)
    dist.gather(mine_msk, g_m, dst=0)
    if rank == 0:
        ref_img, ref_msk, _ = orc.outlier(full2, thr, 1, 2, seed=11)
        assert np.array_equal(deinterleave(torch.stack(g_i), B).numpy(), ref_img)
        assert np.array_equal(deinterleave(torch.stack(g_m), B).numpy(), ref_msk)
        print("DIST_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
