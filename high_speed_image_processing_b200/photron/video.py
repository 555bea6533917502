"""PhotonVideo - PIMS-style access to one Photron recording.

Public surface (names, argument meaning, exceptions) follows the reference class
(src/photron/video.py:275-750) so scripts written against it keep working; the backing store
is different: instead of an eagerly unpacked pyMRAW array it holds a
:class:`~high_speed_image_processing_b200.mraw.FrameStore` over the raw .mraw bytes, which the
flame-front engine streams to the GPU as-is (``raw_frames`` / ``bits``) and which decodes
frames on demand (12-bit: CUDA unpack kernel) for ``video[i]``.
"""
from __future__ import annotations

from datetime import datetime
from pathlib import Path
from typing import Any, Dict, Iterator, Optional, Set, Tuple, Union

import numpy as np

from .. import mraw as _mraw
from .metadata import MetadataConfig
from .timing import SpatialCalibration, TimingInfo, parse_cihx_xml


class PhotonVideo:
    def __init__(self, filepath: str, metadata_fields: Optional[Set[str]] = None, validate: bool = True,
                 trigger_frame: Optional[int] = None, calibration: Optional[SpatialCalibration] = None):
        self._filepath = Path(filepath)
        if validate and not self._filepath.exists():
            raise FileNotFoundError(f"Video file not found: {filepath}")

        self._images, self._raw_info = _mraw.load_video(str(self._filepath))   # seam B1

        self._metadata_config = (MetadataConfig.for_processing() if metadata_fields is None
                                 else MetadataConfig(fields=metadata_fields))
        self._metadata = self._metadata_config.filter_metadata(self._raw_info)

        info = self._raw_info
        self._len = int(info.get("Total Frame", len(self._images)))
        self._frame_shape = (int(info.get("Image Height", self._images.shape[1])),
                             int(info.get("Image Width", self._images.shape[2])))
        self._dtype = self._images.dtype

        # CIHX XML carries the timing fields the decode seam does not publish.
        self._cihx_metadata: Dict[str, Any] = {}
        if self._filepath.suffix.lower() == ".cihx":
            self._cihx_metadata = parse_cihx_xml(self._filepath)
        cihx = self._cihx_metadata
        cihx_ok = cihx.get("record_rate", 0) > 0          # "parsed successfully" in the reference
        frame_rate = cihx["record_rate"] if cihx_ok else int(info.get("Record Rate(fps)", 0))
        start_frame = cihx.get("start_frame", 0) if cihx_ok else int(info.get("Start Frame", 0))
        trig = trigger_frame if trigger_frame is not None else int(info.get("Trigger Frame", 0))
        self._timing = TimingInfo(
            frame_rate=frame_rate, trigger_frame=trig, start_frame=start_frame, pre_trigger_frames=trig,
            recording_datetime=cihx.get("recording_datetime"), recorded_frame=cihx.get("recorded_frame", 0),
            skip_frame=cihx.get("skip_frame", 1))
        self._calibration = calibration

    # ---- identity / metadata ------------------------------------------------------------
    @property
    def filepath(self) -> Path:
        return self._filepath

    @property
    def metadata(self) -> dict:
        return dict(self._metadata)

    @property
    def raw_metadata(self) -> dict:
        return dict(self._raw_info)

    @property
    def cihx_metadata(self) -> Dict[str, Any]:
        return dict(self._cihx_metadata)

    @property
    def recording_datetime(self) -> Optional[datetime]:
        return self._timing.recording_datetime

    @property
    def has_absolute_timing(self) -> bool:
        return self._timing.has_absolute_timing

    @property
    def frame_rate(self) -> int:
        return self._timing.frame_rate

    fps = frame_rate

    @property
    def frame_shape(self) -> Tuple[int, int]:
        return self._frame_shape

    @property
    def height(self) -> int:
        return self._frame_shape[0]

    @property
    def width(self) -> int:
        return self._frame_shape[1]

    @property
    def dtype(self) -> np.dtype:
        return self._dtype

    @property
    def bit_depth(self) -> int:
        return int(self._raw_info.get("EffectiveBit Depth", 16))

    @property
    def shutter_speed(self) -> float:
        return float(self._raw_info.get("Shutter Speed(s)", 0.0))

    exposure_time = shutter_speed

    @property
    def duration(self) -> float:
        return len(self) / self.frame_rate if self.frame_rate > 0 else 0.0

    @property
    def timing(self) -> TimingInfo:
        return self._timing

    @property
    def trigger_frame(self) -> int:
        return self._timing.trigger_frame

    # ---- raw access for the GPU engine --------------------------------------------------
    @property
    def storage_bits(self) -> int:
        """Bits per stored pixel in the .mraw file (8, 12 packed, or 16): the CIH 'Color Bit'."""
        return self._require_open().bits

    @property
    def frame_store(self) -> "_mraw.FrameStore":
        return self._require_open()

    def raw_frames(self, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
        """uint8 view (no copy) of the stored bytes of frames [start, stop)."""
        stop = self._len if stop is None else stop
        if not 0 <= start <= stop <= self._len:
            raise IndexError(f"frame range [{start}, {stop}) out of range [0, {self._len}]")
        return self._require_open().raw_frames(start, stop)

    def pin_memory(self) -> "PhotonVideo":
        """Stage the recording's raw bytes in page-locked host memory (B200 extension; no
        counterpart in the reference): later ``process_video`` calls then stream it with
        asynchronous PCIe copies."""
        self._require_open().pin_memory()
        return self

    def _require_open(self) -> "_mraw.FrameStore":
        if self._images is None:
            raise ValueError("video is closed")
        return self._images

    # ---- calibration / trigger ------------------------------------------------------------
    @property
    def calibration(self) -> Optional[SpatialCalibration]:
        return self._calibration

    @calibration.setter
    def calibration(self, value: Optional[SpatialCalibration]) -> None:
        self._calibration = value

    def set_calibration(self, scale: float, units: str = "m", origin_x: float = 0.0,
                        origin_y: float = 0.0) -> "PhotonVideo":
        self._calibration = SpatialCalibration(scale=scale, units=units, origin_x=origin_x, origin_y=origin_y)
        return self

    def set_trigger_frame(self, frame_index: int) -> "PhotonVideo":
        t = self._timing
        self._timing = TimingInfo(
            frame_rate=t.frame_rate, trigger_frame=frame_index, start_frame=t.start_frame,
            pre_trigger_frames=frame_index, recording_datetime=t.recording_datetime,
            recorded_frame=t.recorded_frame, skip_frame=t.skip_frame)
        return self

    # ---- frame access -----------------------------------------------------------------------
    def __len__(self) -> int:
        return self._len

    def __getitem__(self, key: Union[int, slice]) -> np.ndarray:
        if isinstance(key, int):
            if key < 0:
                key = self._len + key
            if not 0 <= key < self._len:
                raise IndexError(f"Frame index {key} out of range [0, {self._len})")
            return np.array(self._require_open()[key])
        if isinstance(key, slice):
            return np.array(self._require_open()[key])
        raise TypeError(f"Indices must be integers or slices, not {type(key).__name__}")

    def __iter__(self) -> Iterator[np.ndarray]:
        for i in range(self._len):
            yield self[i]

    # ---- time ---------------------------------------------------------------------------------
    def get_time(self, frame_index: int) -> float:
        return self._timing.frame_to_time(frame_index)

    def get_absolute_time(self, frame_index: int) -> float:
        return self._timing.frame_to_absolute_time(frame_index)

    def get_datetime(self, frame_index: int) -> Optional[datetime]:
        return self._timing.frame_to_datetime(frame_index)

    def get_frame_at_time(self, time_seconds: float) -> np.ndarray:
        if self.frame_rate <= 0:
            raise ValueError("Cannot get frame by time: frame rate is 0")
        index = self._timing.time_to_frame(time_seconds)
        return self[max(0, min(index, self._len - 1))]

    def get_time_range(self, start: float, end: float) -> np.ndarray:
        if self.frame_rate <= 0:
            raise ValueError("Cannot get frames by time: frame rate is 0")
        lo = max(0, self._timing.time_to_frame(start))
        hi = min(self._len, self._timing.time_to_frame(end) + 1)
        return self[lo:hi]

    def pixels_to_physical(self, pixels: float) -> float:
        if self._calibration is None:
            raise ValueError("No calibration set. Use set_calibration() first.")
        return self._calibration.pixels_to_physical(pixels)

    def physical_to_pixels(self, physical: float) -> float:
        if self._calibration is None:
            raise ValueError("No calibration set. Use set_calibration() first.")
        return self._calibration.physical_to_pixels(physical)

    def to_float64(self, normalize: bool = True) -> "PhotonVideoFloat64":
        return PhotonVideoFloat64(self, normalize=normalize)

    # ---- lifetime -------------------------------------------------------------------------------
    def close(self) -> None:
        self._images = None

    def __enter__(self) -> "PhotonVideo":
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.close()

    def __repr__(self) -> str:
        return (f"<PhotonVideo '{self._filepath.name}' frames={len(self)} shape={self.frame_shape} "
                f"dtype={self.dtype} fps={self.frame_rate}>")


class PhotonVideoFloat64:
    """float64 view of a PhotonVideo, optionally scaled to [0, 1] by the effective bit depth
    (reference: src/photron/video.py:753-795)."""

    def __init__(self, video: PhotonVideo, normalize: bool = True):
        self._video = video
        self._normalize = normalize
        self._max_value = (2 ** video.bit_depth) - 1

    def _convert(self, frame: np.ndarray) -> np.ndarray:
        out = frame.astype(np.float64)
        if self._normalize:
            out /= self._max_value
        return out

    def __len__(self) -> int:
        return len(self._video)

    def __getitem__(self, key: Union[int, slice]) -> np.ndarray:
        return self._convert(self._video[key])

    def __iter__(self) -> Iterator[np.ndarray]:
        for frame in self._video:
            yield self._convert(frame)

    @property
    def frame_rate(self) -> int:
        return self._video.frame_rate

    @property
    def frame_shape(self) -> Tuple[int, int]:
        return self._video.frame_shape
