"""high_speed_image_processing_b200 - B200-native flame-front pipeline for Photron video.

Drop-in for the per-frame flame-front path of Nadexterbrown/High-Speed-Image-Processing:
the ``photron`` sub-package mirrors the reference's ``src/photron`` API (src/__init__.py:29-61),
``process_videos`` mirrors ``scripts/process_videos.py``, and the compute lives in
``csrc/`` (hand-written sm_100a CUDA behind the C-ABI of ``include/flamefront.h``).
"""
from .photron import (MetadataConfig, MPIVideoProcessor, PhotonVideo, PhotonVideoFloat64,
                      SpatialCalibration, TimingInfo, VideoCollection, open_collection, open_video,
                      parse_cihx_xml)

__version__ = "0.1.0"

__all__ = ["PhotonVideo", "PhotonVideoFloat64", "VideoCollection", "MetadataConfig", "MPIVideoProcessor",
           "SpatialCalibration", "TimingInfo", "parse_cihx_xml", "open_video", "open_collection"]
