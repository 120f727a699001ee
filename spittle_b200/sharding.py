"""Multi-GPU partitioning of the path: independent clips, static round-robin, no collective.

SURVEY.md 8(e): every clip / stream is independent end to end, so rank r of W simply owns the clips
i with i % W == r, runs its own engine on its own GPU, and the host gathers texts by clip index.
torch.distributed is used only to move the (tiny) result strings and timings; NVLink carries nothing.
"""
from __future__ import annotations

from typing import List, Sequence


def owned_indices(n_items: int, world_size: int, rank: int) -> List[int]:
    return list(range(rank, n_items, world_size))


def gather_by_index(local: Sequence, n_items: int, world_size: int, rank: int, dist=None) -> List:
    """Every rank contributes results for owned_indices(); returns the full list on every rank."""
    idx = owned_indices(n_items, world_size, rank)
    assert len(idx) == len(local)
    if dist is None or world_size == 1:
        out = [None] * n_items
        for i, v in zip(idx, local):
            out[i] = v
        return out
    gathered = [None] * world_size
    dist.all_gather_object(gathered, list(zip(idx, local)))
    out = [None] * n_items
    for part in gathered:
        for i, v in part:
            out[i] = v
    return out
