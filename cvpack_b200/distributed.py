"""
Multi-GPU plumbing: one process per GPU, frames sharded by rank, results gathered over
``torch.distributed`` (NCCL on GPUs, gloo on CPU for tests).

Frames (or metadynamics walkers) are independent -- every reference evaluation is a single-context
call (``/root/reference/cvpack/utils.py:140``) -- so the batch dimension is split contiguously
across ranks and the data path needs no collective at all.  The only exchange is the optional
all-gather of results: values ``[B]`` (tiny) and, when asked for, gradients, either dense
``[B, N, 3]`` or restricted to the atoms that can carry a non-zero gradient (``compact=True``),
which is what makes the gather cheap for CVs that touch a small part of a large system.
"""

from __future__ import annotations

import typing as t

import torch
import torch.distributed as dist

from .batch import shard_frames


def _world(group: t.Optional[t.Any]) -> t.Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def local_shard(num_frames: int, group: t.Optional[t.Any] = None) -> t.Tuple[int, int]:
    """``[begin, end)`` of the frames owned by the calling rank."""
    rank, world = _world(group)
    return shard_frames(num_frames, rank, world)


def all_gather_frames(
    local: torch.Tensor, num_frames: int, group: t.Optional[t.Any] = None
) -> torch.Tensor:
    """
    Concatenate per-rank results along the frame dimension.  ``local`` holds the calling rank's
    frames ``[b_local, ...]`` (shards from :func:`local_shard`, possibly of unequal length);
    returns ``[num_frames, ...]`` on every rank.
    """
    rank, world = _world(group)
    if world == 1:
        return local
    counts = [shard_frames(num_frames, r, world) for r in range(world)]
    sizes = [end - begin for begin, end in counts]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((num_frames,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # unequal shards: pad to the largest, gather, trim
    largest = max(sizes)
    padded = torch.zeros((largest,) + tail, dtype=local.dtype, device=local.device)
    padded[: sizes[rank]] = local
    out = torch.empty((world * largest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * largest : r * largest + sizes[r]] for r in range(world)]
    return torch.cat(pieces, dim=0)


def all_gather_gradients(
    local_gradient: torch.Tensor,
    num_frames: int,
    active_atoms: t.Optional[torch.Tensor] = None,
    group: t.Optional[t.Any] = None,
) -> torch.Tensor:
    """
    Gather gradients ``[b_local, N, 3]`` from all ranks.  With ``active_atoms`` (sorted indices of
    the atoms that can carry a non-zero gradient, ``cvp_active_atoms``) only ``[b, n_active, 3]``
    crosses the links and the dense ``[num_frames, N, 3]`` tensor is rebuilt locally.
    """
    if active_atoms is None:
        return all_gather_frames(local_gradient, num_frames, group)
    index = active_atoms.to(local_gradient.device, dtype=torch.long)
    compact = local_gradient.index_select(1, index)
    gathered = all_gather_frames(compact, num_frames, group)
    dense = torch.zeros(
        (num_frames,) + tuple(local_gradient.shape[1:]),
        dtype=local_gradient.dtype,
        device=local_gradient.device,
    )
    dense.index_copy_(1, index, gathered)
    return dense
