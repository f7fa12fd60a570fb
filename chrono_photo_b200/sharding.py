"""Row sharding of one image over ranks / devices, and the final band gather (the only exchange on the path).

Every output pixel depends only on its own column of N samples (src/chrono.rs:169-192, src/simple.rs:102), so GPU g of G
owns rows [H*g/G, H*(g+1)/G) of every frame and no collective runs during compositing. Inside one process the library
gathers bands by per-device D2H copies into the caller's image; across processes (one rank per GPU under torchrun) the
ranks gather their bands on rank 0 with torch.distributed (NCCL on GPU boxes, gloo in the CPU tests).
"""
import numpy as np


def shard_rows(height, rank, world):
    """(row0, rows) of `rank`: the same rule chb_stack_create applies across the devices of a context."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    r0 = height * rank // world
    return r0, height * (rank + 1) // world - r0


class InterleavedShard:
    """Interleaved row-block shard of `rank`: the image is cut into blocks of `block_rows` rows and rank g of G owns the
    blocks with index = g mod G, so that objects (which are compact in the image) spread evenly over the GPUs -- with
    contiguous bands the GPU whose band holds the most object pixels defines the step. The rank's stack holds its blocks back
    to back; `processor_args()` are the chb_outlier_params fields that key the per-pixel RNG by the GLOBAL pixel index."""

    def __init__(self, height, width, rank, world, block_rows):
        if world < 1 or not (0 <= rank < world) or block_rows < 1 or height % (block_rows * world) != 0:
            raise ValueError("interleaved shards need height % (block_rows * world) == 0")
        self.height, self.width, self.rank, self.world, self.block_rows = height, width, rank, world, block_rows
        self.rows = height // world                      # rows this rank owns
        self.row0 = rank * block_rows                    # image row of its first row
        self.block_skip_rows = (world - 1) * block_rows  # rows between the end of one of its blocks and the start of the next

    def global_rows(self):
        """Image row of every local row."""
        r = np.arange(self.rows)
        return self.row0 + r + (r // self.block_rows) * self.block_skip_rows

    def processor_args(self):
        return dict(pixel_offset=self.row0 * self.width, block_pixels=self.block_rows * self.width,
                    block_skip=self.block_skip_rows * self.width)

    def fill_args(self):
        return dict(row0_global=self.row0, full_height=self.height, block_rows=self.block_rows, block_skip_rows=self.block_skip_rows)


def interleave_block_rows(height, world, target=32):
    """Largest block height <= target with height % (block * world) == 0 (None: the image does not divide; use shard_rows)."""
    if world < 1 or height % world != 0:
        return None
    per = height // world
    for b in range(min(target, per), 0, -1):
        if per % b == 0:
            return b
    return None


def deinterleave(parts, block_rows):
    """parts: array-like [world][rows, W, C] in rank order (torch tensor or numpy) -> the (H, W, C) image."""
    world, rows = parts.shape[0], parts.shape[1]
    nb = rows // block_rows
    x = parts.reshape(world, nb, block_rows, *parts.shape[2:])
    x = x.permute(1, 0, 2, 3, 4) if hasattr(x, "permute") else x.transpose(1, 0, 2, 3, 4)
    return x.reshape(world * rows, *parts.shape[2:])


def gather_bands(band, height, dist=None, dst=0):
    """Gathers per-rank (rows_r, W, C) uint8 bands into the full (H, W, C) image on rank `dst` (None elsewhere).
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert band.shape[0] == height
        return band
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    w, c = band.shape[1], band.shape[2]
    max_rows = max(shard_rows(height, r, world)[1] for r in range(world))
    buf = torch.zeros((max_rows, w, c), dtype=torch.uint8, device=dev)
    buf[: band.shape[0]] = torch.from_numpy(np.ascontiguousarray(band)).to(dev)
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst)
        out = np.empty((height, w, c), dtype=np.uint8)
        for r in range(world):
            r0, rows = shard_rows(height, r, world)
            out[r0:r0 + rows] = parts[r][:rows].cpu().numpy()
        return out
    dist.gather(buf, None, dst=dst)
    return None


def max_over_ranks(value, dist=None):
    """Timing reduction of the bench contract: the slowest rank defines the step."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
