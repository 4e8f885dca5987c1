"""One-process-per-GPU sharding of the prover hot path over torch.distributed (NCCL over NVLink on the B200 box; gloo on
CPU for the host-logic tests).

The reference has no multi-GPU layer at all; this is the B200-native addition the north star asks for:
  * column level  — a proof is ~100 independent MSM(n) and several hundred independent NTTs: `column_owner` deals whole
    columns to ranks, the SRS table is replicated, results are all-gathered (96 B per commitment).  No data-path
    collective inside an op.
  * inside one MSM — `ShardedMsm`: the (scalar, base) pairs are split into contiguous slices, each rank runs the local
    Pippenger pipeline on its slice, the per-rank XYZZ partial sums (128 B each) are all-gathered and added locally in
    rank order (NCCL has no group-law reduction), so every rank ends with the same normalised point.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment (torchrun).  Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def column_owner(index: int, world: int) -> int:
    """Round-robin deal of independent units (columns / ops) to ranks."""
    return index % world


def my_columns(count: int, rank: int, world: int):
    return [i for i in range(count) if column_owner(i, world) == rank]


def slice_bounds(n: int, rank: int, world: int):
    """Contiguous split of n pairs across ranks (remainder to the low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_columns(local: torch.Tensor, counts):
    """All-gather variable-count per-rank results [m_r, w] -> list of per-rank tensors (padded exchange)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [local]
    mx = max(counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[:c] for o, c in zip(outs, counts)]


def interleave_columns(per_rank, total: int):
    """Inverse of the round-robin deal: per_rank[r][j] is global column r + j*world."""
    world = len(per_rank)
    out = [None] * total
    for r in range(world):
        for j in range(per_rank[r].shape[0]):
            out[r + j * world] = per_rank[r][j]
    return torch.stack(out) if total else per_rank[0][:0]


class ShardedMsm:
    """MSM split across ranks by base slice; one all-gather of XYZZ partials + a local add."""

    def __init__(self, d_bases_full_or_slice: torch.Tensor, n_total: int, already_sliced: bool = False, window_bits: int = 0):
        from . import device as dev
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_total = n_total
        self.lo, self.hi = slice_bounds(n_total, self.rank, self.world)
        sl = d_bases_full_or_slice if already_sliced else d_bases_full_or_slice[self.lo:self.hi].contiguous()
        assert sl.shape[0] == self.hi - self.lo
        self.bases = dev.DeviceBases(sl, window_bits)

    def partial(self, scalars_local: torch.Tensor) -> torch.Tensor:
        """scalars_local [batch, hi-lo, 4] -> this rank's XYZZ partials [batch, 16]."""
        from . import device as dev
        return dev.msm_batch(self.bases, scalars_local)

    def combine(self, partial: torch.Tensor) -> torch.Tensor:
        """all-gather [batch,16] partials and add them in rank order -> [batch,16], identical on every rank."""
        from . import device as dev
        if self.world == 1:
            return partial
        outs = [torch.empty_like(partial) for _ in range(self.world)]
        dist.all_gather(outs, partial)
        stacked = torch.stack(outs, dim=1).contiguous()      # [batch, world, 16]
        return dev.g1_sum(stacked)

    def __call__(self, scalars_full: torch.Tensor) -> np.ndarray:
        """scalars_full [batch, n_total, 4] (every rank holds the column) -> normalised Jacobian wire [batch, 12]."""
        from . import device as dev
        if scalars_full.dim() == 2:
            scalars_full = scalars_full.unsqueeze(0)
        local = scalars_full[:, self.lo:self.hi].contiguous()
        return dev.normalize(self.combine(self.partial(local)))
