"""Photron CIH / CIHX + MRAW loader - the seam the reference fills with ``pyMRAW.load_video``
(call site src/photron/video.py:332; pyMRAW itself is NOT part of the reference tree).

``load_video(path)`` returns ``(frames, info)`` like pyMRAW does, with two differences that
make it B200-native:

* ``frames`` is a :class:`FrameStore`, an array-like ``[N,H,W]`` view over the *raw* .mraw
  bytes (memory-mapped, never unpacked on the CPU).  Indexing it yields NumPy arrays; for
  packed 12-bit data the decode runs in the CUDA unpack kernel (``ff_unpack``) - there is no
  CPU decoder in this package, so 12-bit frame access without a GPU raises.
* the raw packed bytes are exposed (``FrameStore.raw``) so the flame-front engine can stream
  them to the GPU without ever materialising uint16 frames.

``info`` carries the keys pyMRAW publishes and the reference consumes
(src/photron/video.py:343-346,360,367,378,468,473; src/photron/metadata.py:38-61).
The 12-bit layout (3 bytes -> 2 pixels) is pyMRAW's; see ``oracle/flame_oracle.py``.
"""
from __future__ import annotations

import warnings
import xml.etree.ElementTree as ET
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

SUPPORTED_FILE_FORMATS = ("mraw",)
SUPPORTED_EFFECTIVE_BIT_SIDE = ("lower", "higher")
SUPPORTED_BITS = (8, 12, 16)


# --------------------------------------------------------------------------------------
# header parsing
# --------------------------------------------------------------------------------------
def _text(node: Optional[ET.Element], path: str, default: Optional[str] = None) -> Optional[str]:
    if node is None:
        return default
    found = node.find(path)
    if found is None or found.text is None:
        return default
    return found.text.strip()


def extract_cih_xml(data: bytes) -> Optional[str]:
    """The ``<cih>...</cih>`` document embedded after a CIHX file's binary preamble."""
    start = data.find(b"<cih>")
    if start < 0:
        return None
    end = data.find(b"</cih>", start)
    if end < 0:
        return None
    return data[start:end + len(b"</cih>")].decode("utf-8", errors="ignore")


def _cihx_info(path: Path) -> Dict[str, Any]:
    xml = extract_cih_xml(path.read_bytes())
    if xml is None:
        raise ValueError(f"{path}: no <cih> XML block found")
    root = ET.fromstring(xml)
    info: Dict[str, Any] = {
        "Date": _text(root, "fileInfo/date", ""),
        "Camera Type": _text(root, "deviceInfo/deviceName", ""),
        "Record Rate(fps)": float(_text(root, "recordInfo/recordRate", "0")),
        "Shutter Speed(s)": float(_text(root, "recordInfo/shutterSpeed", "0")),
        "Total Frame": int(_text(root, "frameInfo/totalFrame", "0")),
        "Original Total Frame": int(_text(root, "frameInfo/recordedFrame", "0")),
        "Image Width": int(_text(root, "imageDataInfo/resolution/width", "0")),
        "Image Height": int(_text(root, "imageDataInfo/resolution/height", "0")),
        "File Format": _text(root, "imageFileInfo/fileFormat", ""),
        "EffectiveBit Depth": int(_text(root, "imageDataInfo/effectiveBit/depth", "0")),
        "EffectiveBit Side": _text(root, "imageDataInfo/effectiveBit/side", ""),
        "Color Bit": int(_text(root, "imageDataInfo/colorInfo/bit", "0")),
        "Comment Text": _text(root, "basicInfo/comment", "") or "",
    }
    return info


_CIH_NUMERIC = {
    "Record Rate(fps)": float, "Shutter Speed(s)": float, "Total Frame": int,
    "Original Total Frame": int, "Image Width": int, "Image Height": int,
    "EffectiveBit Depth": int, "Color Bit": int, "Start Frame": int, "Trigger Frame": int,
}


def _cih_info(path: Path) -> Dict[str, Any]:
    """Legacy text header: ``Key : Value`` lines, ``#`` comments."""
    info: Dict[str, Any] = {}
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if line.startswith("#") or ":" not in line:
                continue
            key, value = line.split(":", 1)
            key, value = key.strip(), value.strip()
            if not key or key in info:
                continue
            conv = _CIH_NUMERIC.get(key)
            if conv is not None:
                try:
                    if key == "Shutter Speed(s)" and "/" in value:      # written as "1/20000"
                        num, den = value.split("/", 1)
                        info[key] = float(num) / float(den)
                    else:
                        info[key] = conv(float(value)) if conv is int else conv(value)
                except ValueError:
                    info[key] = value
            else:
                info[key] = value
    return info


def get_cih(path: Union[str, Path]) -> Dict[str, Any]:
    """Parse a .cih / .cihx header and validate it the way the decode seam expects."""
    path = Path(path)
    ext = path.suffix.lower()
    if ext == ".cihx":
        info = _cihx_info(path)
    elif ext == ".cih":
        info = _cih_info(path)
    else:
        raise ValueError(f"Unsupported configuration file ({ext or 'no extension'}); expected .cih or .cihx")
    fmt = str(info.get("File Format", ""))
    if fmt.lower() not in SUPPORTED_FILE_FORMATS:
        raise ValueError(f"Unexpected File Format: {fmt!r}")
    side = str(info.get("EffectiveBit Side", ""))
    if side.lower() not in SUPPORTED_EFFECTIVE_BIT_SIDE:
        raise ValueError(f"Unexpected EffectiveBit Side: {side!r}")
    bits = int(info.get("Color Bit", 0))
    if bits not in SUPPORTED_BITS:
        raise ValueError(f"only 8-, 12- and 16-bit MRAW files are supported (Color Bit = {bits})")
    if int(info.get("Original Total Frame", 0)) > int(info.get("Total Frame", 0)):
        warnings.warn(f"Clipped footage! (Total frame: {info['Total Frame']}, "
                      f"Original total frame: {info['Original Total Frame']})")
    return info


# --------------------------------------------------------------------------------------
# frame store
# --------------------------------------------------------------------------------------
class FrameStore:
    """Array-like ``[N,H,W]`` over raw MRAW bytes (file-backed memmap or any uint8 buffer)."""

    def __init__(self, raw: np.ndarray, n_frames: int, height: int, width: int, bits: int):
        if bits not in SUPPORTED_BITS:
            raise ValueError(f"unsupported bit depth {bits}")
        px = height * width
        if bits == 12 and (n_frames * px) % 2:
            raise ValueError("packed 12-bit data needs an even total pixel count")
        self.bits = bits
        self.frame_px = px
        self.frame_bytes = px * bits // 8 if not (bits == 12 and px % 2) else None
        need = n_frames * px * bits // 8
        raw = raw.reshape(-1)
        if raw.dtype != np.uint8:
            raw = raw.view(np.uint8)
        if raw.size < need:
            raise ValueError(f"MRAW data too short: {raw.size} bytes, need {need}")
        self.raw = raw[:need]
        self.shape: Tuple[int, int, int] = (n_frames, height, width)
        self.dtype = np.dtype(np.uint8) if bits == 8 else np.dtype(np.uint16)
        self.ndim = 3

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def nbytes_raw(self) -> int:
        return int(self.raw.size)

    def pin_memory(self) -> "FrameStore":
        """Copy the raw bytes into page-locked host memory (once), so that the engine's chunked
        H2D copies run asynchronously at PCIe speed instead of through the driver's pageable
        staging.  The file read (page cache -> pinned buffer) happens here, off the timed path."""
        if getattr(self, "_pinned", None) is None:
            import torch
            pinned = torch.empty(self.raw.size, dtype=torch.uint8, pin_memory=True)
            dst = pinned.numpy()
            step = 256 << 20
            for a in range(0, self.raw.size, step):        # bounded slices: no second full copy in RAM
                dst[a:a + step] = self.raw[a:a + step]
            self._pinned = pinned                           # keeps the allocation alive
            self.raw = dst
        return self

    @property
    def is_pinned(self) -> bool:
        return getattr(self, "_pinned", None) is not None

    def raw_frames(self, start: int, stop: int) -> np.ndarray:
        """uint8 view of the packed bytes of frames [start, stop)."""
        if self.frame_bytes is None:
            raise ValueError("odd-sized 12-bit frames do not start on byte boundaries")
        return self.raw[start * self.frame_bytes: stop * self.frame_bytes]

    def _decode(self, start: int, stop: int) -> np.ndarray:
        n, h, w = stop - start, self.shape[1], self.shape[2]
        if n <= 0:
            return np.empty((0, h, w), dtype=self.dtype)
        if self.bits == 8:
            return np.array(self.raw_frames(start, stop).reshape(n, h, w))
        if self.bits == 16:
            return np.array(self.raw_frames(start, stop).view("<u2").reshape(n, h, w))
        # packed 12-bit: decode on the GPU (stage-1 kernel); no CPU decoder exists here
        from .engine import get_engine
        eng = get_engine()
        return eng.unpack(eng.upload(self.raw_frames(start, stop)), n, h, w, 12).cpu().numpy()

    def __getitem__(self, key):
        n = self.shape[0]
        if isinstance(key, (int, np.integer)):
            k = int(key)
            if k < 0:
                k += n
            if not 0 <= k < n:
                raise IndexError(f"index {key} is out of bounds for axis 0 with size {n}")
            return self._decode(k, k + 1)[0]
        if isinstance(key, slice):
            start, stop, step = key.indices(n)
            if step == 1:
                return self._decode(start, stop)
            idx = range(start, stop, step)
            if len(idx) == 0:
                return np.empty((0,) + self.shape[1:], dtype=self.dtype)
            lo, hi = min(idx), max(idx) + 1
            block = self._decode(lo, hi)
            return block[[i - lo for i in idx]]
        raise TypeError(f"FrameStore indices must be integers or slices, not {type(key).__name__}")


def load_images(mraw: Union[str, Path], height: int, width: int, n_frames: int, bit: int = 16) -> FrameStore:
    raw = np.memmap(str(mraw), dtype=np.uint8, mode="r")
    return FrameStore(raw, n_frames, height, width, int(bit))


def load_video(cih_file: Union[str, Path]) -> Tuple[FrameStore, Dict[str, Any]]:
    """Drop-in for ``pyMRAW.load_video``: (frames, info)."""
    cih_file = Path(cih_file)
    info = get_cih(cih_file)
    mraw_file = cih_file.with_suffix(".mraw")
    if not mraw_file.exists():
        raise FileNotFoundError(f"MRAW data file not found: {mraw_file}")
    frames = load_images(mraw_file, int(info["Image Height"]), int(info["Image Width"]),
                         int(info["Total Frame"]), int(info["Color Bit"]))
    return frames, info
