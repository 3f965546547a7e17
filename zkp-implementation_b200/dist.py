"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo on CPU
for the host-logic tests).

* MSM shards by point range: rank g owns bases/scalars [g*n/G, (g+1)*n/G), runs a full
  Pippenger on its shard and contributes one un-normalised XYZZ partial (192 bytes).  NCCL has no
  group-addition reduce op, so the partials are all-gathered as bytes and folded on the host
  (G - 1 additions + one inversion).  No data-path collective besides that.
* Batches of independent polynomials shard whole across ranks (no collective at all).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous point range of `rank` (the first n % world ranks get one extra point)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(partial: np.ndarray, group=None, device=None) -> np.ndarray:
    """All-gather one 24 x u64 XYZZ record per rank -> (world, 24) array, identical on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


def msm_sharded(engine, scalars_dev, bases_dev, n_local: int, group=None, device=None):
    """Point-range-sharded MSM: local Pippenger, all-gather of partials, host fold.

    Returns (xy limbs, infinity flag) of the full sum on every rank."""
    partial = engine.msm_partial_dev(scalars_dev, bases_dev, n_local)
    allp = gather_partials(partial, group=group, device=device)
    return engine.fold_partials(allp)


def batch_shard(batch: int, rank: int, world: int) -> range:
    """Whole-polynomial sharding of a batch of independent NTTs: polynomial i goes to rank i % world."""
    return range(rank, batch, world)
