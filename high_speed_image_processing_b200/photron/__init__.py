"""Photron video access (mirror of the reference package ``src/photron``,
src/photron/__init__.py:14-112): ``open_video``, ``open_collection`` and the classes behind
them, backed by raw-MRAW frame stores that the B200 engine consumes directly."""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Set, Union

from .collection import VideoCollection
from .metadata import MetadataConfig
from .parallel import MPIVideoProcessor
from .timing import SpatialCalibration, TimingInfo, parse_cihx_xml
from .video import PhotonVideo, PhotonVideoFloat64


def open_video(filepath: str, metadata_fields: Optional[Set[str]] = None, trigger_frame: Optional[int] = None,
               calibration: Optional[SpatialCalibration] = None) -> PhotonVideo:
    """Open one .cihx/.cih recording."""
    return PhotonVideo(filepath, metadata_fields=metadata_fields, trigger_frame=trigger_frame,
                       calibration=calibration)


def open_collection(source: Union[str, List[str]], pattern: str = "*.cihx", recursive: bool = False,
                    metadata_fields: Optional[Set[str]] = None, trigger_frame: Optional[int] = None,
                    calibration: Optional[SpatialCalibration] = None) -> VideoCollection:
    """Open a directory (globbed with ``pattern``) or an explicit list of files."""
    if isinstance(source, (str, Path)) and Path(source).is_dir():
        return VideoCollection.from_directory(source, pattern=pattern, recursive=recursive,
                                              metadata_fields=metadata_fields, trigger_frame=trigger_frame,
                                              calibration=calibration)
    if isinstance(source, list):
        return VideoCollection.from_files(source, metadata_fields=metadata_fields, trigger_frame=trigger_frame,
                                          calibration=calibration)
    raise ValueError("source must be a directory path or list of file paths")


__all__ = ["PhotonVideo", "PhotonVideoFloat64", "VideoCollection", "MetadataConfig", "MPIVideoProcessor",
           "SpatialCalibration", "TimingInfo", "parse_cihx_xml", "open_video", "open_collection"]
