"""Host-side plumbing of the row-slab sharded pressure solve (one process per GPU).

The data path needs no collective: slab halos and dot products move through peer memory inside
the CG kernels (include/fsb.h, "multi-GPU").  What the host has to do once per context is hand
every rank the CUDA IPC handles of all ranks; this module does that with torch.distributed
(NCCL on GPUs, gloo in the CPU tests).  torch is plumbing only.
"""
import numpy as np

from .capi import SHARD_BLOB_BYTES


def slab_rows(ny, world, rank):
    """Rows [lo, hi) rank `rank` iterates on: the even split fsb_shard_connect uses."""
    return ny * rank // world, ny * (rank + 1) // world


def gather_blobs(blob, dist=None, device=None):
    """All-gather one fixed-size uint8 blob per rank; returns a (world, len) numpy array in rank
    order.  `device`: where the staging tensors live (cuda for NCCL, cpu for gloo)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def connect(sim, dist=None, device=None):
    """Export this rank's handles, exchange them, connect the slabs.  Collective: every rank of
    the default process group must call it.  Returns this rank's (row_lo, row_hi)."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return 0, sim.ny
    blobs = gather_blobs(sim.shard_export(), dist, device)
    assert blobs.shape == (world, SHARD_BLOB_BYTES)
    sim.shard_connect(rank, world, blobs)
    dist.barrier()  # nobody starts a solve before every rank has opened its peers
    lo, hi = sim.shard_rows()
    assert (lo, hi) == slab_rows(sim.ny, world, rank)
    return lo, hi
