"""Timing and spatial-calibration value types, and the CIHX timing-metadata reader.

API and arithmetic mirror the reference (src/photron/video.py:31-150 parse_cihx_xml,
:153-183 SpatialCalibration, :186-272 TimingInfo) so ``Time_s`` comes out bit-identical:
both conversions are a single correctly-rounded float64 division of Python ints.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass
from datetime import datetime, timedelta
from pathlib import Path
from typing import Any, Dict, Optional



def _int_field(parent: Optional[ET.Element], tag: str) -> Optional[int]:
    if parent is None:
        return None
    node = parent.find(tag)
    if node is None or not node.text:
        return None
    return int(node.text)


def parse_cihx_xml(filepath) -> Dict[str, Any]:
    """Timing metadata of a CIHX file (reference: src/photron/video.py:31-150).

    Never raises: on any failure the defaults below are returned (and a warning printed),
    which makes ``PhotonVideo`` fall back to the decode seam's info dict."""
    meta: Dict[str, Any] = {
        "recording_datetime": None,
        "record_rate": 0,
        "recorded_frame": 0,
        "start_frame": 0,
        "total_frame": 0,
        "skip_frame": 1,
        "irig_enabled": False,
        "shutter_speed_ns": 0,
    }
    try:
        data = Path(filepath).read_bytes()
        head = data.find(b"<?xml")
        if head < 0:
            head = data.find(b"<cih>")
        if head < 0:
            return meta
        tail = data.find(b"</cih>", head)
        if tail < 0:
            return meta
        root = ET.fromstring(data[head:tail + len(b"</cih>")].decode("utf-8", errors="ignore"))

        finfo = root.find("fileInfo")
        if finfo is not None:
            date, time = finfo.find("date"), finfo.find("time")
            if date is not None and time is not None:
                try:
                    meta["recording_datetime"] = datetime.strptime(f"{date.text} {time.text}",
                                                                   "%Y/%m/%d %H:%M:%S")
                except ValueError:
                    pass

        frame = root.find("frameInfo")
        for key, tag in (("recorded_frame", "recordedFrame"), ("total_frame", "totalFrame"),
                         ("start_frame", "startFrame"), ("skip_frame", "skipFrame")):
            val = _int_field(frame, tag)
            if val is not None:
                meta[key] = val

        rec = root.find("recordInfo")
        val = _int_field(rec, "recordRate")
        if val is not None:
            meta["record_rate"] = val
        val = _int_field(rec, "shutterSpeedNsec")
        if val is not None:
            meta["shutter_speed_ns"] = val

        dev = root.find("deviceInfo")
        val = _int_field(dev, "irig")
        if val is not None:
            meta["irig_enabled"] = val != 0
        if meta["record_rate"] == 0:
            val = _int_field(dev, "recordRate")
            if val is not None:
                meta["record_rate"] = val
    except Exception as exc:  # same contract as the reference: swallow and use defaults
        print(f"Warning: Failed to parse CIHX XML: {exc}")
    return meta


@dataclass
class SpatialCalibration:
    """Pixels <-> physical units (reference: src/photron/video.py:153-183)."""
    scale: float
    units: str = "m"
    origin_x: float = 0.0
    origin_y: float = 0.0

    def pixels_to_physical(self, pixels: float) -> float:
        return pixels * self.scale

    def physical_to_pixels(self, physical: float) -> float:
        return physical / self.scale

    def x_to_physical(self, x_pixels: float) -> float:
        return (x_pixels - self.origin_x) * self.scale

    def y_to_physical(self, y_pixels: float) -> float:
        return (y_pixels - self.origin_y) * self.scale


@dataclass
class TimingInfo:
    """Frame index <-> time (reference: src/photron/video.py:186-272)."""
    frame_rate: int
    trigger_frame: int = 0
    start_frame: int = 0
    pre_trigger_frames: int = 0
    recording_datetime: Optional[datetime] = None
    recorded_frame: int = 0
    skip_frame: int = 1

    def frame_to_time(self, frame_index: int) -> float:
        """Seconds relative to the trigger frame (negative before it)."""
        if self.frame_rate <= 0:
            return 0.0
        return (frame_index - self.trigger_frame) / self.frame_rate

    def frame_to_absolute_time(self, frame_index: int) -> float:
        """Seconds from recording start: (start_frame + index*skip_frame) / rate."""
        if self.frame_rate <= 0:
            return 0.0
        return (self.start_frame + (frame_index * self.skip_frame)) / self.frame_rate

    def frame_to_datetime(self, frame_index: int) -> Optional[datetime]:
        if self.recording_datetime is None or self.frame_rate <= 0:
            return None
        return self.recording_datetime + timedelta(seconds=self.frame_to_absolute_time(frame_index))

    def time_to_frame(self, time_seconds: float) -> int:
        if self.frame_rate <= 0:
            return 0
        return int(time_seconds * self.frame_rate) + self.trigger_frame

    @property
    def has_absolute_timing(self) -> bool:
        return self.recording_datetime is not None and self.frame_rate > 0
