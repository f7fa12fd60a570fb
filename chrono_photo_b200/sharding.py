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
