"""Multi-GPU plumbing: cells shard embarrassingly (every cell is an independent series in the
reference — skdownscale/pointwise_models/core.py:86-96, 137-141), so each rank fits and predicts
its own contiguous range of the flattened cell axis with NO data-path collective.  The only
exchange is the optional gather of the predicted field at the end (NCCL over NVLink on GPUs;
gloo works for CPU tensors, which is how the host logic is tested).

torch.distributed is plumbing here: one process per GPU, launched with torchrun.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def cell_range(n_cells: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) of the flattened cell axis owned by ``rank``; the first
    ``n_cells % world_size`` ranks take one extra cell."""
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(n_cells, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_cells(field: torch.Tensor, world_size: int | None = None, rank: int | None = None) -> torch.Tensor:
    """This rank's column block of a ``[..., cells]`` array (a view, no copy)."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    a, b = cell_range(field.shape[-1], world_size, rank)
    return field[..., a:b]


def gather_cells(local: torch.Tensor, n_cells: int, group=None, stacked: bool = False) -> torch.Tensor:
    """All-gather the per-rank column blocks ``[T, (k,) C_rank]`` into the full ``[T, (k,) n_cells]``
    field on every rank (one NCCL all-gather over NVLink + one local re-interleave).

    ``stacked=True`` skips the re-interleave and returns the collective's own layout
    ``[world, T, (k,) C_max]`` (rank-major; uneven shards zero-padded to the widest)."""
    world = dist.get_world_size(group)
    spans = [cell_range(n_cells, world, r) for r in range(world)]
    wmax = max(b - a for a, b in spans)
    lead = tuple(local.shape[:-1])
    mine = local
    if local.shape[-1] != wmax or not local.is_contiguous():
        mine = torch.zeros(lead + (wmax,), dtype=local.dtype, device=local.device)
        mine[..., : local.shape[-1]] = local
    # concatenation form along dim 0 (accepted by both NCCL and gloo), viewed as [world, ...] afterwards
    flat = torch.empty((world * mine.shape[0],) + tuple(mine.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, mine, group=group)
    full = flat.view((world,) + tuple(mine.shape))
    if stacked:
        return full
    if all(b - a == wmax for a, b in spans):
        return full.movedim(0, -2).reshape(lead + (world * wmax,))     # [.., world, C_rank] → cells contiguous
    return torch.cat([full[r][..., : (b - a)] for r, (a, b) in enumerate(spans)], dim=-1)
