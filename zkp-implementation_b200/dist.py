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


class DistNtt:
    """Four-step Fr NTT of size 2^log_n over `world` GPUs (one process each), SURVEY.md 8e.

    N is viewed as 2^r rows x 2^s columns (r = ``rows_log``, s = log_n - r).  Rank g holds

    * layout A (input of forward / output of inverse): its column block, row-major --
      ``a[j1, c] = x[j1 * 2^s + g * 2^s / G + c]``;
    * layout B (output of forward / input of inverse): its rows of the result --
      ``b[k', k2] = X[(g * 2^r / G + k') + 2^r * k2]``.

    forward = stage kernel (column transforms of size 2^r on the local block + omega_N^(col * k) twiddles)
    -> all-to-all transpose -> batched local transforms of size 2^s.  Two transports for the transpose:

    * ``p2p=False``: `torch.distributed.all_to_all_single` (NCCL over NVLink; gloo in the CPU tests) with a
      small permute kernel on the receiving / sending side;
    * ``p2p=True``: the exchange is fused into the stage kernel -- every rank opens the peers' exchange
      buffers through CUDA IPC and the kernel stores (forward) / loads (inverse) rows directly in peer
      memory over NVLink; a barrier replaces the collective.

    Buffers are flat int64 torch tensors of 4 words per Fr element (device tensors on GPUs, CPU tensors on
    the emulator)."""

    def __init__(self, engine, log_n: int, rank: int, world: int, group=None, p2p: bool = False):
        self.eng, self.log_n, self.rank, self.world, self.group, self.p2p = engine, log_n, rank, world, group, p2p
        self.r = engine.ntt_dist_rows_log(log_n, world)
        if self.r == 0:
            raise ValueError(f"2^{log_n} cannot be split over {world} ranks")
        self.s = log_n - self.r
        self.local = 1 << (log_n - (world.bit_length() - 1))
        self.rows_local = (1 << self.r) // world
        self._exch = 0
        self._peers = None
        self._bind_stream()
        if p2p:
            self._open_peers()

    def _bind_stream(self) -> None:
        """Stream contract: the engine's kernels (stage, permute, local transforms, copies) and torch's
        collectives / allocations must be ordered on ONE stream.  The engine launches asynchronously on its own
        stream by default, so every transform first binds it to torch's current stream on this device (a no-op
        on the CPU emulator, whose calls are synchronous)."""
        import torch

        if torch.cuda.is_available() and self.eng.is_cuda():
            self.eng.set_stream(torch.cuda.current_stream(self.eng.device).cuda_stream)

    # -- layouts (host helpers for tests / callers that hold the natural-order vector) ---------------
    def layout_a(self, x: np.ndarray) -> np.ndarray:
        """Natural-order vector (N x 4 u64) -> this rank's layout-A block."""
        cols = (1 << self.s) // self.world
        m = x.reshape(1 << self.r, 1 << self.s, 4)
        return np.ascontiguousarray(m[:, self.rank * cols:(self.rank + 1) * cols, :]).reshape(-1, 4)

    def layout_b(self, x: np.ndarray) -> np.ndarray:
        """Natural-order vector X (N x 4 u64) -> this rank's layout-B block."""
        m = x.reshape(1 << self.s, 1 << self.r, 4)  # [k2][k1]
        blk = m[:, self.rank * self.rows_local:(self.rank + 1) * self.rows_local, :]
        return np.ascontiguousarray(blk.transpose(1, 0, 2)).reshape(-1, 4)

    # -- transports ----------------------------------------------------------------------------------
    def _open_peers(self) -> None:
        import torch
        import torch.distributed as dist

        self._exch = self.eng.dev_alloc(self.local * 32)
        h = torch.tensor(list(self.eng.ipc_export(self._exch)), dtype=torch.uint8, device=f"cuda:{self.eng.device}")
        hs = [torch.empty_like(h) for _ in range(self.world)]
        dist.all_gather(hs, h, group=self.group)
        self._peers = []
        for g, t in enumerate(hs):
            self._peers.append(self._exch if g == self.rank else self.eng.ipc_open(bytes(t.cpu().tolist())))

    def close(self) -> None:
        if self._peers:
            for g, p in enumerate(self._peers):
                if g != self.rank:
                    self.eng.ipc_close(p)
            self._peers = None
        if self._exch:
            self.eng.dev_free(self._exch)
            self._exch = 0

    def _sync_ranks(self) -> None:
        import torch
        import torch.distributed as dist

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _all_to_all(self, src):
        import torch
        import torch.distributed as dist

        dst = torch.empty_like(src)
        dist.all_to_all_single(dst, src, group=self.group)
        return dst

    # -- transforms ----------------------------------------------------------------------------------
    def forward(self, data, coset: Optional[int] = None):
        """layout A (consumed) -> layout B.  Returns the tensor holding B (p2p: a fresh tensor copied out of
        the exchange buffer; use `forward_into_exchange` to keep it there)."""
        import torch

        e = self.eng
        self._bind_stream()
        if self.p2p:
            self.forward_into_exchange(data, coset)
            out = torch.empty_like(data)
            e.dev_copy(out, self._exch, self.local * 32)
            return out
        e.ntt_dist_stage_dev(data, self.log_n, self.rank, self.world, False, coset)
        recv = self._all_to_all(data)
        out = torch.empty_like(recv)
        e.ntt_dist_permute_dev(recv, out, self.log_n, self.world, False)
        e.ntt_dev(out, self.s, batch=self.rows_local)
        return out

    def forward_into_exchange(self, data, coset: Optional[int] = None) -> int:
        """p2p transport: layout A in `data` -> layout B left in this rank's exchange buffer (device pointer)."""
        e = self.eng
        self._bind_stream()
        e.ntt_dist_stage_dev(data, self.log_n, self.rank, self.world, False, coset, peers=self._peers)
        self._sync_ranks()  # every rank's rows have landed
        e.ntt_dev(self._exch, self.s, batch=self.rows_local)
        return self._exch

    def inverse(self, data, coset: Optional[int] = None):
        """layout B (consumed) -> layout A."""
        import torch

        e = self.eng
        self._bind_stream()
        if self.p2p:
            e.dev_copy(self._exch, data, self.local * 32)
            out = torch.empty_like(data)
            self.inverse_from_exchange(out, coset)
            return out
        e.ntt_dev(data, self.s, batch=self.rows_local, inverse=True)
        send = torch.empty_like(data)
        e.ntt_dist_permute_dev(data, send, self.log_n, self.world, True)
        recv = self._all_to_all(send)
        e.ntt_dist_stage_dev(recv, self.log_n, self.rank, self.world, True, coset)
        return recv

    def inverse_from_exchange(self, out, coset: Optional[int] = None) -> None:
        """p2p transport: layout B in the exchange buffer -> layout A written to `out`."""
        e = self.eng
        self._bind_stream()
        e.ntt_dev(self._exch, self.s, batch=self.rows_local, inverse=True)
        self._sync_ranks()  # peers' rows are final before anyone pulls them
        e.ntt_dist_stage_dev(out, self.log_n, self.rank, self.world, True, coset, peers=self._peers)
        self._sync_ranks()  # nobody overwrites an exchange buffer that is still being read
