"""MPIVideoProcessor - kept for API parity with the reference (src/photron/parallel.py:16-302).

The reference's only parallelism is this mpi4py wrapper (round-robin frames over CPU ranks).
The B200 path replaces it with contiguous frame ranges / whole videos per GPU and NCCL for
the one real exchange (``high_speed_image_processing_b200.sharding``); this class remains so
that scripts constructing it keep running, including the serial ``comm=None`` mode every
method degrades to.  ``comm`` may be any object with mpi4py's communicator methods.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Tuple, TypeVar

import numpy as np

from .collection import VideoCollection

T = TypeVar("T")


def split_indices(total_count: int, rank: int, size: int, distribution: str = "round_robin") -> List[int]:
    """The reference's two index partitions (src/photron/parallel.py:99-115)."""
    if distribution == "round_robin":
        return list(range(rank, total_count, size)) if total_count > 0 else []
    if distribution == "contiguous":
        base, extra = divmod(total_count, size)
        start = rank * base + min(rank, extra)
        return list(range(start, start + base + (1 if rank < extra else 0)))
    raise ValueError(f"Unknown distribution strategy: {distribution}")


class MPIVideoProcessor:
    def __init__(self, comm=None):
        self._comm = comm
        self._rank = comm.Get_rank() if comm is not None else 0
        self._size = comm.Get_size() if comm is not None else 1

    @property
    def rank(self) -> int:
        return self._rank

    @property
    def size(self) -> int:
        return self._size

    @property
    def is_root(self) -> bool:
        return self._rank == 0

    @property
    def is_parallel(self) -> bool:
        return self._comm is not None and self._size > 1

    def distribute_indices(self, total_count: int, distribution: str = "round_robin") -> List[int]:
        return split_indices(total_count, self._rank, self._size, distribution)

    def _gather_sorted(self, local: List[Tuple[int, Any]], gather_results: bool):
        if not (gather_results and self._comm is not None):
            return local
        parts = self._comm.gather(local, root=0)
        if not self.is_root:
            return None
        merged = [item for part in parts for item in part]
        merged.sort(key=lambda item: item[0])
        return merged

    def process_collection(self, collection: VideoCollection, process_func: Callable[[np.ndarray, int], T],
                           gather_results: bool = True,
                           distribution: str = "round_robin") -> Optional[List[Tuple[int, T]]]:
        local = [(g, process_func(collection.get_global_frame(g), g))
                 for g in self.distribute_indices(collection.total_frames, distribution)]
        return self._gather_sorted(local, gather_results)

    def process_videos(self, collection: VideoCollection, process_video_func: Callable[[Any, int], T],
                       gather_results: bool = True) -> Optional[List[Tuple[int, T]]]:
        local = [(v, process_video_func(collection[v], v)) for v in self.distribute_indices(len(collection))]
        return self._gather_sorted(local, gather_results)

    def broadcast(self, data: Any, root: int = 0) -> Any:
        return self._comm.bcast(data, root=root) if self._comm is not None else data

    def gather(self, data: Any, root: int = 0) -> Optional[List[Any]]:
        return self._comm.gather(data, root=root) if self._comm is not None else [data]

    def scatter(self, data: Optional[List[Any]], root: int = 0) -> Any:
        if self._comm is not None:
            return self._comm.scatter(data, root=root)
        return data[0] if data else None

    def barrier(self) -> None:
        if self._comm is not None:
            self._comm.Barrier()

    def reduce_sum(self, data: np.ndarray, root: int = 0) -> Optional[np.ndarray]:
        if self._comm is None:
            return data
        from mpi4py import MPI
        if self.is_root:
            out = np.zeros_like(data)
            self._comm.Reduce(data, out, op=MPI.SUM, root=root)
            return out
        self._comm.Reduce(data, None, op=MPI.SUM, root=root)
        return None

    def allreduce_sum(self, data: np.ndarray) -> np.ndarray:
        if self._comm is None:
            return data
        from mpi4py import MPI
        out = np.zeros_like(data)
        self._comm.Allreduce(data, out, op=MPI.SUM)
        return out

    def __repr__(self) -> str:
        return f"<MPIVideoProcessor rank={self._rank}/{self._size} mode={'parallel' if self.is_parallel else 'serial'}>"
