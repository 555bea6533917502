"""Multi-GPU decomposition of the flame-front path (one process per GPU, torch.distributed).

Frames shard naturally:

* within a clip: contiguous frame ranges, one per rank (the contiguous variant of the
  reference's ``distribute_indices``, src/photron/parallel.py:101-113), each prefixed by a
  one-frame halo read from host memory - the only cross-frame dependency is the one-frame
  look-back of the frame difference (scripts/process_videos.py:397-399,469);
* across clips: whole videos round-robin over ranks (src/photron/parallel.py:173-208).

The single real exchange is tiny: an all-reduce(min) of the first exit frame (one int32)
followed by an all-gather of the per-rank position arrays.  It replaces the reference's
pickled ``comm.gather`` (scripts/process_videos.py:1533-1541) and turns the per-rank ``break``
(:1494) into a global truncation (README.md:145-149).  NCCL on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._cabi import FF_NO_EXIT, FF_POS_DROPPED


def contiguous_range(total: int, rank: int, size: int) -> Tuple[int, int]:
    """[start, stop) of rank's block; the first ``total % size`` ranks get one extra item."""
    if size <= 0 or not 0 <= rank < size:
        raise ValueError(f"bad rank/size {rank}/{size}")
    base, extra = divmod(max(0, total), size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def assign_videos(n_videos: int, rank: int, size: int, weights: Optional[Sequence[int]] = None) -> List[int]:
    """Videos handled by ``rank``.  Default: round-robin like ``process_videos``
    (src/photron/parallel.py:192).  With ``weights`` (e.g. bytes per video): greedy
    longest-first balancing, deterministic on every rank."""
    if weights is None:
        return list(range(rank, n_videos, size))
    if len(weights) != n_videos:
        raise ValueError("weights must have one entry per video")
    loads = [0] * size
    owner = [0] * n_videos
    for v in sorted(range(n_videos), key=lambda i: (-weights[i], i)):
        r = min(range(size), key=lambda k: (loads[k], k))
        owner[v] = r
        loads[r] += weights[v]
    return [v for v in range(n_videos) if owner[v] == rank]


@dataclass
class GatheredRange:
    first_exit_t: torch.Tensor      # int32[1] global index of the first exit frame, or FF_NO_EXIT
    pos: torch.Tensor               # int32[total] positions of the whole clip, truncated
    counts: Optional[torch.Tensor]  # int32[total] above-noise counts (if provided)

    @property
    def first_exit(self) -> int:    # synchronises on the device scalar
        return int(self.first_exit_t.item())


class RangeExchange:
    """The exit-min + result-gather step over a process group."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.size = dist.get_world_size(group) if self.active else 1

    def my_range(self, total: int) -> Tuple[int, int]:
        return contiguous_range(total, self.rank, self.size)

    def exit_min(self, first_exit: torch.Tensor) -> torch.Tensor:
        """In-place all-reduce(min) of the int32[1] first-exit scalar."""
        if self.active and self.size > 1:
            dist.all_reduce(first_exit, op=dist.ReduceOp.MIN, group=self.group)
        return first_exit

    def gather_ranges(self, local: torch.Tensor, total: int, fill: int = FF_POS_DROPPED) -> torch.Tensor:
        """All-gather per-rank int32 blocks (contiguous_range layout) into int32[total]."""
        if not (self.active and self.size > 1):
            return local
        base, extra = divmod(total, self.size)
        width = base + (1 if extra else 0)
        padded = torch.full((width,), fill, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
        out = torch.empty(self.size * width, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        if extra == 0:
            return out
        parts = []
        for r in range(self.size):
            a, b = contiguous_range(total, r, self.size)
            parts.append(out[r * width: r * width + (b - a)])
        return torch.cat(parts)

    def finish(self, pos_local: torch.Tensor, first_exit_local: torch.Tensor, total: int,
               truncate, counts_local: Optional[torch.Tensor] = None) -> GatheredRange:
        """exit-min, local truncation against the GLOBAL exit frame, then the gathers.

        ``truncate(pos_local, first_frame, first_exit)`` applies the truncation in place; on
        GPUs it is ``FlameFrontEngine.truncate`` (the ff_truncate kernel)."""
        start, _ = self.my_range(total)
        fe = self.exit_min(first_exit_local)
        truncate(pos_local, start, fe)
        pos = self.gather_ranges(pos_local, total)
        counts = None if counts_local is None else self.gather_ranges(counts_local, total, fill=0)
        return GatheredRange(fe, pos, counts)
