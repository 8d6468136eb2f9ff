"""Multi-GPU sharding of query batches: queries are independent and read-only
against the navmesh, so a batch of N queries is cut into `world` contiguous slices, one per
rank / GPU, each rank answers its slice against its own replica of the navmesh, and the
results are concatenated in rank order.  There is no collective on the data path; the optional
`gather` below (torch.distributed all_gather of the per-rank result arrays) only exists for
callers that want the whole result on every rank.

Slices are contiguous and balanced (sizes differ by at most one, larger slices first), aligned
to `granule` queries so that units that must stay on one device (an env's try_step -> find_path
chain, the G goals of one multi-goal start) are never split.  Random-point streams are keyed by
the GLOBAL query index (`query0` of hbn_random_points), so results do not depend on `world`.
"""
from __future__ import annotations

import numpy as np


def shard_slices(n: int, world: int, granule: int = 1):
    """[(begin, end)] * world: contiguous, balanced, granule-aligned cover of range(n)."""
    if world <= 0 or granule <= 0 or n < 0:
        raise ValueError("world, granule must be positive and n non-negative")
    if n % granule:
        raise ValueError("n must be a multiple of granule")
    units = n // granule
    base, extra = divmod(units, world)
    out, b = [], 0
    for r in range(world):
        e = b + (base + (1 if r < extra else 0)) * granule
        out.append((b, e))
        b = e
    return out


def shard_slice(n: int, rank: int, world: int, granule: int = 1):
    return shard_slices(n, world, granule)[rank]


def gather(local: np.ndarray, n: int, rank: int, world: int, granule: int = 1, group=None) -> np.ndarray:
    """Concatenate the per-rank result arrays (first axis = queries) on every rank.
    Uses torch.distributed (gloo for host arrays, nccl for CUDA tensors)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    slices = shard_slices(n, world, granule)
    is_t = isinstance(local, torch.Tensor)
    t = local if is_t else torch.from_numpy(np.ascontiguousarray(local))
    assert t.shape[0] == slices[rank][1] - slices[rank][0], "local result does not match this rank's slice"
    parts = [torch.empty((e - b,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for b, e in slices]
    if len({p.shape for p in parts}) == 1:
        dist.all_gather(parts, t.contiguous(), group=group)
    else:  # ragged slices: one broadcast per rank
        for r, p in enumerate(parts):
            if r == rank:
                p.copy_(t)
            dist.broadcast(p, src=r, group=group)
    out = torch.cat(parts, 0)
    return out if is_t else out.numpy()
