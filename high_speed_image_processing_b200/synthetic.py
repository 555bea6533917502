"""Synthetic Photron recordings (CIHX + MRAW) for tests and benchmarks.

The reference ships no sample data, so this writer defines the inputs
every parity test and benchmark runs on: integer background noise
``clip(round(N(40, 4^2)))``, a flame-free frame 0, and a bright region behind a front that
moves left-to-right with a Gaussian-CDF (erfc) leading edge.  "nova" style adds an exponential
tail behind the front (peaked profile); "mini" keeps a plateau.

Two renderers produce the same *model*: NumPy (per-frame seeded, chunk-independent; used by
tests, where CUDA results and oracle results are compared on the identical array) and torch
(on the GPU, for multi-GB benchmark clips).  They use different random streams.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from pathlib import Path
from typing import Optional

import numpy as np


@dataclass(frozen=True)
class SyntheticSpec:
    width: int = 512
    height: int = 64
    n_frames: int = 500
    bits: int = 12                 # storage: 8, 12 (packed) or 16
    effective_bits: int = 12
    style: str = "nova"            # "nova" (decaying tail) | "mini" (plateau)
    t_enter: float = 20.0          # frame at which the front is at x = x_enter
    x_enter: float = 0.0
    velocity: float = 1.0          # px / frame
    amplitude: float = 2500.0
    edge_sigma: float = 3.0
    tail_length: float = 120.0     # nova: e-folding length behind the front (px)
    curvature_px: float = 6.0      # front lags by this many px at the top/bottom rows
    noise_mean: float = 40.0
    noise_std: float = 4.0
    seed: int = 1000
    record_rate: int = 160000
    start_frame: int = 500
    skip_frame: int = 1
    shutter_ns: int = 2500
    date: str = "2023/10/4"
    time: str = "14:29:21"
    camera: str = "FASTCAM NOVA S16"

    @property
    def max_value(self) -> int:
        return (1 << min(self.effective_bits, 8 if self.bits == 8 else 16)) - 1

    @property
    def frame_bytes(self) -> int:
        return self.width * self.height * self.bits // 8

    def front_position(self, t: float) -> float:
        return self.x_enter + self.velocity * (t - self.t_enter)

    def exit_frame(self, margin_px: int = 10) -> int:
        """First frame whose ideal front position reaches width - margin."""
        return int(math.ceil(self.t_enter + (self.width - margin_px - self.x_enter) / self.velocity))


# Named workloads of BASELINE.json:configs (frame counts can be overridden for tests).
def config_spec(name: str, n_frames: Optional[int] = None, seed: Optional[int] = None) -> SyntheticSpec:
    name = name.upper()
    if name == "C1":    # 512x64 x 500, half_maximum
        spec = SyntheticSpec(width=512, height=64, n_frames=500, style="nova", t_enter=20.0, velocity=1.1,
                             seed=1000)
    elif name == "C2":  # Nova-style 1024x128 x 20000, half_maximum: long empty lead-in, front enters late
        spec = SyntheticSpec(width=1024, height=128, n_frames=20000, style="nova", t_enter=18900.0,
                             velocity=1.0, seed=2000)
    elif name == "C3":  # Mini-style 1024x256 x 20000, threshold, exit ~ frame 15000
        spec = SyntheticSpec(width=1024, height=256, n_frames=20000, style="mini", t_enter=15000.0 - 1014.0,
                             velocity=1.0, seed=3000, camera="FASTCAM Mini AX200", record_rate=100000)
    elif name == "C4":  # 1024x1024 x 5000, gradient, diff retained
        spec = SyntheticSpec(width=1024, height=1024, n_frames=5000, style="mini", t_enter=3900.0, velocity=1.0,
                             curvature_px=24.0, seed=4000)
    else:
        raise ValueError(f"unknown config {name!r}")
    if n_frames is not None:
        # keep the flame inside shorter clips: scale the entry time with the clip length
        scale = n_frames / spec.n_frames
        spec = replace(spec, n_frames=n_frames, t_enter=max(2.0, spec.t_enter * scale))
    if seed is not None:
        spec = replace(spec, seed=seed)
    return spec


# --------------------------------------------------------------------------------------
# NumPy renderer
# --------------------------------------------------------------------------------------
def _flame_numpy(spec: SyntheticSpec, t: int) -> np.ndarray:
    from scipy.special import erfc
    xf = spec.front_position(float(t))
    h, w = spec.height, spec.width
    if xf < -6.0 * spec.edge_sigma:
        return np.zeros((h, w), dtype=np.float64)
    rows = (np.arange(h, dtype=np.float64) - (h // 2)) / max(1.0, h / 2.0)
    front = xf - spec.curvature_px * rows ** 2                      # [h]
    u = np.arange(w, dtype=np.float64)[None, :] - front[:, None]    # x - x_f(row)
    img = spec.amplitude * 0.5 * erfc(u / (math.sqrt(2.0) * spec.edge_sigma))
    if spec.style == "nova":
        img = img * np.exp(-np.maximum(-u, 0.0) / spec.tail_length)
    elif spec.style != "mini":
        raise ValueError(f"unknown style {spec.style!r}")
    return img


def render_frames(spec: SyntheticSpec, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
    """Frames [start, stop) as an integer array [n,H,W] (uint8 for 8-bit storage else uint16).
    Each frame has its own seed, so any sub-range equals the same slice of the full render."""
    stop = spec.n_frames if stop is None else stop
    dtype = np.uint8 if spec.bits == 8 else np.uint16
    out = np.empty((stop - start, spec.height, spec.width), dtype=dtype)
    scale = 1.0 if spec.bits != 8 else 255.0 / 4095.0
    for i, t in enumerate(range(start, stop)):
        rng = np.random.default_rng([spec.seed, t])
        img = rng.normal(spec.noise_mean, spec.noise_std, size=(spec.height, spec.width))
        if t > 0:                                    # frame 0 stays flame-free (reference :1360)
            img = img + _flame_numpy(spec, t)
        out[i] = np.clip(np.rint(img * scale), 0, spec.max_value).astype(dtype)
    return out


def pack_frames(frames: np.ndarray, bits: int) -> np.ndarray:
    """[n,H,W] integer frames -> the exact bytes of a .mraw file (uint8 1-D)."""
    if bits == 8:
        return np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1)
    if bits == 16:
        return np.ascontiguousarray(frames.astype("<u2")).view(np.uint8).reshape(-1)
    if bits == 12:
        px = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1)
        if px.size % 2:
            raise ValueError("packed 12-bit needs an even number of pixels")
        if px.size and int(px.max()) > 0xFFF:
            raise ValueError("pixel value exceeds 12 bits")
        p0, p1 = px[0::2], px[1::2]
        out = np.empty((p0.size, 3), dtype=np.uint8)
        out[:, 0] = p0 >> 4
        out[:, 1] = ((p0 & 15) << 4) | (p1 >> 8)
        out[:, 2] = p1 & 255
        return out.reshape(-1)
    raise ValueError(f"unsupported bit depth {bits}")


def render_packed(spec: SyntheticSpec, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
    return pack_frames(render_frames(spec, start, stop), spec.bits)


# --------------------------------------------------------------------------------------
# CIHX / CIH writers
# --------------------------------------------------------------------------------------
def cihx_bytes(spec: SyntheticSpec, total_frames: Optional[int] = None) -> bytes:
    """A CIHX file: binary preamble, then the <cih> XML document with every element the
    reference (src/photron/video.py:86-144) and the decode seam read, then a binary tail."""
    n = spec.n_frames if total_frames is None else total_frames
    xml = f"""<?xml version="1.0" encoding="utf-8"?>
<cih>
  <fileInfo><version>1.0</version><date>{spec.date}</date><time>{spec.time}</time></fileInfo>
  <basicInfo><comment>synthetic flame front ({spec.style})</comment></basicInfo>
  <deviceInfo><deviceName>{spec.camera}</deviceName><irig>0</irig></deviceInfo>
  <recordInfo>
    <recordRate>{spec.record_rate}</recordRate>
    <shutterSpeed>{1.0 / (1e9 / spec.shutter_ns):.10f}</shutterSpeed>
    <shutterSpeedNsec>{spec.shutter_ns}</shutterSpeedNsec>
  </recordInfo>
  <frameInfo>
    <totalFrame>{n}</totalFrame><recordedFrame>{n}</recordedFrame>
    <startFrame>{spec.start_frame}</startFrame><skipFrame>{spec.skip_frame}</skipFrame>
  </frameInfo>
  <imageFileInfo><fileFormat>MRaw</fileFormat></imageFileInfo>
  <imageDataInfo>
    <resolution><width>{spec.width}</width><height>{spec.height}</height></resolution>
    <effectiveBit><depth>{spec.effective_bits}</depth><side>Lower</side></effectiveBit>
    <colorInfo><type>Mono</type><bit>{spec.bits}</bit></colorInfo>
  </imageDataInfo>
</cih>"""
    preamble = b"\x00\x01\x02\x03CIHX\xff\xfe" + bytes(range(16, 64))
    return preamble + xml.encode("utf-8") + b"\x00\x00\xde\xad\xbe\xef"


def cih_text(spec: SyntheticSpec, total_frames: Optional[int] = None) -> str:
    n = spec.n_frames if total_frames is None else total_frames
    return "\n".join([
        "#Camera Information Header",
        f"Date : {spec.date}",
        f"Camera Type : {spec.camera}",
        f"Record Rate(fps) : {spec.record_rate}",
        f"Shutter Speed(s) : 1/{int(round(1e9 / spec.shutter_ns))}",
        f"Total Frame : {n}",
        f"Original Total Frame : {n}",
        f"Start Frame : {spec.start_frame}",
        "Trigger Frame : 0",
        f"Image Width : {spec.width}",
        f"Image Height : {spec.height}",
        "File Format : MRaw",
        f"EffectiveBit Depth : {spec.effective_bits}",
        "EffectiveBit Side : Lower",
        f"Color Bit : {spec.bits}",
        "Comment Text : synthetic",
        "END", ""])


def write_clip(directory, stem: str, spec: SyntheticSpec, frames: Optional[np.ndarray] = None,
               header: str = "cihx", chunk: int = 256) -> Path:
    """Write ``<stem>.cihx`` (or .cih) + ``<stem>.mraw`` and return the header path."""
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    mraw = directory / f"{stem}.mraw"
    with open(mraw, "wb") as fh:
        if frames is not None:
            fh.write(pack_frames(frames, spec.bits).tobytes())
        else:
            for a in range(0, spec.n_frames, chunk):
                fh.write(render_packed(spec, a, min(spec.n_frames, a + chunk)).tobytes())
    n = spec.n_frames if frames is None else len(frames)
    if header == "cihx":
        path = directory / f"{stem}.cihx"
        path.write_bytes(cihx_bytes(spec, n))
    elif header == "cih":
        path = directory / f"{stem}.cih"
        path.write_text(cih_text(spec, n))
    else:
        raise ValueError("header must be 'cihx' or 'cih'")
    return path


# --------------------------------------------------------------------------------------
# torch renderer (GPU) for benchmark-sized clips
# --------------------------------------------------------------------------------------
def render_packed_torch(spec: SyntheticSpec, device, start: int = 0, stop: Optional[int] = None,
                        chunk: Optional[int] = None, out=None):
    """Packed bytes of frames [start, stop) as a uint8 tensor on ``device`` (same model as the
    NumPy renderer, torch's Philox stream).  Synthetic-data generation only - not the hot path.

    The noise is drawn per ALIGNED block of ``chunk`` frames (seeded with the block's first frame), so any
    sub-range equals the same slice of the full render - ranks that render their own frame ranges, a halo
    frame rendered on its own and a whole-clip render on one GPU all see the same recording."""
    import torch
    stop = spec.n_frames if stop is None else stop
    n = stop - start
    fb = spec.frame_bytes
    if out is None:
        out = torch.empty(n * fb, dtype=torch.uint8, device=device)
    h, w = spec.height, spec.width
    if chunk is None:                                    # a function of the frame shape only
        chunk = max(16, min(512, (1 << 27) // (h * w)))
    gen = torch.Generator(device=device)
    rows = (torch.arange(h, device=device, dtype=torch.float32) - (h // 2)) / max(1.0, h / 2.0)
    lag = spec.curvature_px * rows ** 2                                    # [h]
    xs = torch.arange(w, device=device, dtype=torch.float32)
    inv = 1.0 / (math.sqrt(2.0) * spec.edge_sigma)
    for blk in range((start // chunk) * chunk, stop, chunk):
        a, b = max(start, blk), min(stop, blk + chunk)
        gen.manual_seed(spec.seed * 1_000_003 + blk)
        img = torch.randn((chunk, h, w), generator=gen, device=device, dtype=torch.float32)[a - blk:b - blk]
        img.mul_(spec.noise_std).add_(spec.noise_mean)
        t = torch.arange(a, b, device=device, dtype=torch.float32)
        xf = spec.x_enter + spec.velocity * (t - spec.t_enter)              # [n]
        active = (xf >= -6.0 * spec.edge_sigma) & (t > 0)
        if bool(active.any()):
            u = xs[None, None, :] - (xf[:, None, None] - lag[None, :, None])
            flame = spec.amplitude * 0.5 * torch.special.erfc(u * inv)
            if spec.style == "nova":
                flame = flame * torch.exp(-torch.clamp(-u, min=0.0) / spec.tail_length)
            img += flame * active[:, None, None]
        if spec.bits == 8:
            img.mul_(255.0 / 4095.0)
        px = torch.clamp(torch.round(img), 0, spec.max_value).to(torch.int32).reshape(-1)
        dst = out[(a - start) * fb:(b - start) * fb]
        if spec.bits == 8:
            dst.copy_(px.to(torch.uint8))
        elif spec.bits == 16:
            pair = torch.stack((px & 255, px >> 8), dim=1).to(torch.uint8)
            dst.copy_(pair.reshape(-1))
        else:
            p0, p1 = px[0::2], px[1::2]
            trip = torch.stack((p0 >> 4, ((p0 & 15) << 4) | (p1 >> 8), p1 & 255), dim=1).to(torch.uint8)
            dst.copy_(trip.reshape(-1))
    return out
