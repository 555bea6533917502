"""VideoCollection - several PhotonVideo objects behind one global frame index
(reference: src/photron/collection.py:15-395).  Config 5 of BASELINE.json shards a
collection by whole videos across GPUs (see ``sharding.assign_videos``)."""
from __future__ import annotations

from bisect import bisect_right
from pathlib import Path
from typing import Any, Callable, Iterator, List, Optional, Set, Tuple, Union

import numpy as np

from .timing import SpatialCalibration
from .video import PhotonVideo


class VideoCollection:
    def __init__(self, videos: List[PhotonVideo], metadata_fields: Optional[Set[str]] = None):
        self._videos = videos
        self._metadata_fields = metadata_fields
        self._build_index()

    def _build_index(self) -> None:
        # offsets[i] = global index of the first frame of video i; offsets[-1] = total
        offsets = [0]
        for v in self._videos:
            offsets.append(offsets[-1] + len(v))
        self._cumulative_lengths = offsets
        self._total_frames = offsets[-1]

    # ---- constructors -------------------------------------------------------------------
    @classmethod
    def from_directory(cls, directory: Union[str, Path], pattern: str = "*.cihx", recursive: bool = False,
                       metadata_fields: Optional[Set[str]] = None,
                       calibration: Optional[SpatialCalibration] = None,
                       trigger_frame: Optional[int] = None) -> "VideoCollection":
        root = Path(directory)
        if not root.exists():
            raise FileNotFoundError(f"Directory not found: {directory}")
        files = sorted(root.rglob(pattern) if recursive else root.glob(pattern))
        videos = []
        for f in files:
            try:
                videos.append(PhotonVideo(str(f), metadata_fields=metadata_fields, calibration=calibration,
                                          trigger_frame=trigger_frame))
            except Exception as exc:   # one bad file must not sink the batch (reference :112-114)
                print(f"Warning: Could not load {f}: {exc}")
        return cls(videos, metadata_fields)

    @classmethod
    def from_files(cls, filepaths: List[Union[str, Path]], metadata_fields: Optional[Set[str]] = None,
                   calibration: Optional[SpatialCalibration] = None,
                   trigger_frame: Optional[int] = None) -> "VideoCollection":
        videos = [PhotonVideo(str(fp), metadata_fields=metadata_fields, calibration=calibration,
                              trigger_frame=trigger_frame) for fp in filepaths]
        return cls(videos, metadata_fields)

    # ---- container protocol ---------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._videos)

    def __iter__(self) -> Iterator[PhotonVideo]:
        return iter(self._videos)

    def __getitem__(self, idx: int) -> PhotonVideo:
        return self._videos[idx]

    @property
    def videos(self) -> List[PhotonVideo]:
        return list(self._videos)

    @property
    def total_frames(self) -> int:
        return self._total_frames

    @property
    def filepaths(self) -> List[Path]:
        return [v.filepath for v in self._videos]

    # ---- global indexing ----------------------------------------------------------------------
    def _resolve_global_index(self, global_idx: int) -> Tuple[int, int]:
        if global_idx < 0:
            global_idx += self._total_frames
        if not 0 <= global_idx < self._total_frames:
            raise IndexError(f"Global frame index {global_idx} out of range [0, {self._total_frames})")
        # first video whose end offset exceeds the index (skips zero-length videos like the
        # reference's linear scan does)
        vid = bisect_right(self._cumulative_lengths, global_idx) - 1
        return vid, global_idx - self._cumulative_lengths[vid]

    def global_to_local(self, global_idx: int) -> Tuple[int, int]:
        return self._resolve_global_index(global_idx)

    def local_to_global(self, video_idx: int, local_idx: int) -> int:
        if video_idx < 0 or video_idx >= len(self._videos):
            raise IndexError(f"Video index {video_idx} out of range")
        return self._cumulative_lengths[video_idx] + local_idx

    def get_global_frame(self, global_idx: int) -> np.ndarray:
        vid, local = self._resolve_global_index(global_idx)
        return self._videos[vid][local]

    def get_global_time(self, global_idx: int) -> float:
        vid, local = self._resolve_global_index(global_idx)
        return self._videos[vid].get_time(local)

    # ---- bulk helpers ----------------------------------------------------------------------------
    def map_frames(self, func: Callable[[np.ndarray, int, int], Any],
                   frame_indices: Optional[List[int]] = None,
                   video_indices: Optional[List[int]] = None) -> List[Any]:
        out: List[Any] = []
        if frame_indices is not None:
            for g in frame_indices:
                vid, local = self._resolve_global_index(g)
                out.append(func(self._videos[vid][local], vid, local))
            return out
        for vid in (video_indices if video_indices is not None else range(len(self._videos))):
            video = self._videos[vid]
            for local in range(len(video)):
                out.append(func(video[local], vid, local))
        return out

    def iter_frames(self) -> Iterator[Tuple[np.ndarray, int, int, float]]:
        for vid, video in enumerate(self._videos):
            for local in range(len(video)):
                yield video[local], vid, local, video.get_time(local)

    def set_calibration_all(self, scale: float, units: str = "m", origin_x: float = 0.0,
                            origin_y: float = 0.0) -> "VideoCollection":
        for v in self._videos:
            v.set_calibration(scale, units, origin_x, origin_y)
        return self

    def set_trigger_frame_all(self, frame_index: int) -> "VideoCollection":
        for v in self._videos:
            v.set_trigger_frame(frame_index)
        return self

    def summary(self) -> str:
        lines = [f"VideoCollection: {len(self)} videos, {self.total_frames} total frames", "-" * 60]
        lines += [f"  [{i}] {v.filepath.name}: {len(v)} frames @ {v.frame_rate} fps"
                  for i, v in enumerate(self._videos)]
        return "\n".join(lines)

    def close_all(self) -> None:
        for v in self._videos:
            v.close()

    def __enter__(self) -> "VideoCollection":
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.close_all()

    def __repr__(self) -> str:
        return f"<VideoCollection videos={len(self)} total_frames={self.total_frames}>"
