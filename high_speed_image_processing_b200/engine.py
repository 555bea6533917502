"""FlameFrontEngine - the Python face of libflamefront.so.

PyTorch supplies device memory and streams; every computation on frames happens in the
hand-written sm_100a kernels behind the C-ABI (include/flamefront.h).  The few float64
scalars the reference derives per clip (centre-row mean/std, thresholds, the empty-frame
fraction) are evaluated on the host with the reference's own NumPy expressions and handed to
the kernels as exact integer bounds - see ``derive_kernel_bounds``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi
from ._cabi import (FF_DIFF_F32, FF_DIFF_F64, FF_DIFF_NONE, FF_DIFF_U16, FF_METHOD, FF_NO_EXIT, FF_PX_F64,
                    FF_PX_U8, FF_PX_U16, HostArgs, RangeArgs, RangeHooks)

DETECTION_METHODS = tuple(FF_METHOD)  # ("threshold", "gradient", "half_maximum")
INT32_MAX = 2**31 - 1
INT32_MIN = -(2**31)

_DIFF_DTYPES = {
    None: (FF_DIFF_NONE, None),
    "uint16": (FF_DIFF_U16, torch.uint16),
    "float32": (FF_DIFF_F32, torch.float32),
    "float64": (FF_DIFF_F64, torch.float64),
}


def frame_nbytes(height: int, width: int, bits: int) -> int:
    if bits not in (8, 12, 16):
        raise ValueError(f"unsupported bit depth {bits} (8, 12 or 16)")
    px = height * width
    if bits == 12 and px % 2:
        raise ValueError("packed 12-bit frames need an even pixel count")
    return px * bits // 8


@dataclass
class DetectionParams:
    """Tunables of the per-frame path.  Defaults are the reference's
    (scripts/process_videos.py:112,169,174,1459) except ``exit_margin_px`` = 10
    (README.md:146; HEAD's FlameDetectorConfig uses 15, :193)."""
    method: str = "half_maximum"
    use_frame_diff: bool = True
    frame_diff_threshold: float = 5.0
    min_gradient_strength: float = 10.0
    min_run_px: int = 1
    exit_margin_px: int = 10
    min_signal_fraction: float = 0.0005

    def __post_init__(self) -> None:
        if self.method not in FF_METHOD:
            raise ValueError(
                f"unknown detection_method {self.method!r}; options: {', '.join(DETECTION_METHODS)}")
        if self.min_run_px < 1:
            raise ValueError("min_run_px must be >= 1")


@dataclass
class ClipScalars:
    """Per-clip scalars taken from frame 0 (scripts/process_videos.py:1357-1370, :1458)."""
    background: float
    centerline_mean: float
    centerline_std: float
    centerline_max: float
    flame_threshold: float
    noise_threshold: float

    @staticmethod
    def from_frame0_stats(bg_max: int, centerline: np.ndarray) -> "ClipScalars":
        background = float(bg_max)                              # :1358 float(np.max(frame))
        line = np.asarray(centerline).astype(np.float64)        # :1362
        mean = np.mean(line)                                    # :1363
        std = np.std(line)                                      # :1364
        mx = np.max(line)                                       # :1365
        thr = max(mean + 5 * std, mx * 2.0)                     # :1367-1370
        noise = max(10.0, background * 0.5)                     # :1458
        return ClipScalars(background, float(mean), float(std), float(mx), float(thr), float(noise))


@dataclass
class KernelBounds:
    empty_thr: int
    min_signal_count: int
    diff_thr: int
    threshold_floor: int
    grad2_bound: int


def _clamp_i32(v: int) -> int:
    return max(INT32_MIN, min(INT32_MAX, int(v)))


def min_signal_count(n_pixels: int, fraction: float) -> int:
    """Smallest integer count c for which ``c / n_pixels < fraction`` is False, evaluated with
    the same float64 division/compare as is_empty_frame (scripts/process_videos.py:759-763)."""
    c = max(0, int(fraction * n_pixels) - 2)
    while c <= n_pixels and (c / n_pixels < fraction):
        c += 1
    while c > 0 and not ((c - 1) / n_pixels < fraction):
        c -= 1
    return c


def derive_kernel_bounds(scalars: ClipScalars, params: DetectionParams, n_pixels: int) -> KernelBounds:
    """Turn the reference's float comparisons on integer-valued data into exact integer ones:
    for integer v and real t:  v > t <=> v > floor(t);  v < t <=> v < ceil(t)."""
    return KernelBounds(
        empty_thr=_clamp_i32(math.floor(scalars.noise_threshold)),
        min_signal_count=min_signal_count(n_pixels, params.min_signal_fraction),
        diff_thr=_clamp_i32(math.ceil(params.frame_diff_threshold)),
        threshold_floor=_clamp_i32(math.floor(scalars.flame_threshold)),
        grad2_bound=_clamp_i32(math.ceil(-2.0 * params.min_gradient_strength)),
    )


class PendingScalars:
    """The clip's frame-0 statistics, still on the device: the clip-scalar block and the centre row of frame 0
    that the prep kernel wrote.  They are copied to the host and the float64 statistics are evaluated with
    NumPy (the reference's own expressions, scripts/process_videos.py:1362-1370) only when somebody asks for
    them - nothing is queued and the host never waits inside ``process_range``.  For the threshold method the
    value the kernels used (computed on the device in NumPy's order of operations) is compared with NumPy's
    and a difference raises."""

    def __init__(self, scal_dev: torch.Tensor, line_dev: torch.Tensor, check_threshold: bool):
        self._scal_dev, self._line_dev, self._check = scal_dev, line_dev, check_threshold
        self._value: Optional[ClipScalars] = None
        self.block: Optional[np.ndarray] = None       # the int32[16] clip-scalar block as the kernels saw it

    def get(self) -> ClipScalars:
        if self._value is None:
            block = self._scal_dev.cpu().numpy()
            line = self._line_dev.cpu().numpy()
            value = ClipScalars.from_frame0_stats(int(block[0]), line)
            if self._check:
                dev_floor = int(block[1])
                if dev_floor != _clamp_i32(math.floor(value.flame_threshold)):
                    raise RuntimeError(
                        f"flame threshold evaluated on the device (floor {dev_floor}) differs from NumPy's "
                        f"({value.flame_threshold!r}): this NumPy sums in another order than the prep kernel "
                        "assumes - pass scalars=/bg_dev= from clip_scalars() instead of frame0")
            self._value, self.block = value, block
            self._scal_dev = self._line_dev = None
        return self._value


class RangeResult:
    """Device-resident outputs of one contiguous frame range."""

    def __init__(self, first_frame: int, pos: torch.Tensor, counts: torch.Tensor, first_exit: torch.Tensor,
                 diff: Optional[torch.Tensor] = None, profiles: Optional[torch.Tensor] = None,
                 decoded: Optional[torch.Tensor] = None, scalars=None):
        self.first_frame = first_frame
        self.pos = pos                    # int32[n]  (>=0, -1 none, -2 dropped after truncate)
        self.counts = counts              # int32[n]  above-noise pixel count
        self.first_exit = first_exit      # int32[1]  global frame index or FF_NO_EXIT
        self.diff = diff                  # [n,H,W]
        self.profiles = profiles          # int32[n,W]
        self.decoded = decoded            # uint16[n,H,W]
        self._scalars = scalars           # ClipScalars | PendingScalars | None

    @property
    def scalars(self) -> Optional[ClipScalars]:
        """Per-clip scalars (host float64).  Synchronises on two tiny device-to-host copies the first time."""
        if isinstance(self._scalars, PendingScalars):
            self._scalars = self._scalars.get()
        return self._scalars


@dataclass
class HeadRangeResult:
    """Device-resident outputs of the HEAD-parity detector on one frame range."""
    first_frame: int
    track: torch.Tensor               # int32[n,5]: final, min_gradient, rightmost_sobel, search_start, search_end
    flags: torch.Tensor               # uint8[n]: 0 not processed, 1 detected on a difference image, 2 no prior
    stop: torch.Tensor                # int32[3]: exit frame or FF_NO_EXIT, last detection frame, last position
    lines: Optional[torch.Tensor]     # float64[n,2,W] (Sobel / gradient centre rows) when kept
    scalars: Optional[ClipScalars] = None


@dataclass
class HostResult:
    first_frame: int
    pos: Optional[np.ndarray]         # None when the results stayed on the device (range block)
    counts: Optional[np.ndarray]
    first_exit: int                   # this range's own first exit frame (global index) or FF_NO_EXIT
    frames_done: int
    bytes_uploaded: int = 0           # host -> device bytes actually moved (the early exit keeps it small)


SCALAR_BLOCK = 16                     # int32 words of the clip-scalar block ([0] background scalar)
MAX_DEVICE_STATS_WIDTH = 4096         # widest centre row whose float64 statistics prep_kernel evaluates


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class FlameFrontEngine:
    """One engine per GPU.  All methods launch on the current torch CUDA stream."""

    def __init__(self, device: Union[int, str, torch.device, None] = None,
                 host_chunk_bytes: int = 64 << 20, host_copy_threads: Optional[int] = None):
        self._lib = _cabi.load()                      # raises if the .so is missing
        if not torch.cuda.is_available():
            raise RuntimeError("FlameFrontEngine needs a CUDA device (no CPU fallback exists)")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("FlameFrontEngine runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._host_chunk_bytes = int(host_chunk_bytes)
        # ff_process_host copies pageable (memory-mapped) sources into pinned bounce buffers with a
        # thread pool, by default half the host's cores; under torchrun the ranks share the cores
        self._copy_threads: Optional[int] = host_copy_threads
        if host_copy_threads is None and "FF_HOST_COPY_THREADS" not in os.environ:
            ranks = os.environ.get("LOCAL_WORLD_SIZE", "1")
            if ranks.isdigit() and int(ranks) > 1:
                self._copy_threads = max(1, min(8, (os.cpu_count() or 2) // int(ranks)))
        self._host_ctx: Optional[C.c_void_p] = None
        self._ws: Optional[torch.Tensor] = None        # ff_process_range workspace (kept zero-filled by the kernels)
        self.launches = 0                             # kernels launched through this engine
        self._side_stream = None
        self._pinned = {}
        self._stream_events = None                    # bench hook: [(start, stop)] around every 4th ff_process_range
        self._stream_event_tick = 0
        self._stream_events_every = 4

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check_dev(self, t: torch.Tensor, name: str) -> None:
        if t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")

    def close(self) -> None:
        if self._host_ctx is not None:
            self._lib.ff_host_ctx_destroy(self._host_ctx)
            self._host_ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ stage 1
    def unpack(self, packed: torch.Tensor, n_frames: int, height: int, width: int, bits: int) -> torch.Tensor:
        """Decode ``n_frames`` device-resident frames to [n,H,W] (uint16; uint8 for 8-bit)."""
        self._check_dev(packed, "packed")
        need = n_frames * frame_nbytes(height, width, bits)
        if packed.dtype != torch.uint8 or packed.numel() < need:
            raise ValueError(f"packed must be uint8 with at least {need} bytes")
        out = torch.empty((n_frames, height, width), dtype=torch.uint8 if bits == 8 else torch.uint16,
                          device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_unpack(packed.data_ptr(), out.data_ptr(), n_frames, height, width, bits,
                                            self._stream()), "ff_unpack")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ stage 2a
    def background(self, frame0: torch.Tensor, height: int, width: int, bits: int):
        """Launch the background reduction on frame 0.  Returns device tensors
        ``(bg_max int32[1], centerline uint16[W])`` without synchronising."""
        self._check_dev(frame0, "frame0")
        if frame0.dtype != torch.uint8 or frame0.numel() < frame_nbytes(height, width, bits):
            raise ValueError("frame0 must be a uint8 tensor holding one whole frame")
        bg = torch.empty(SCALAR_BLOCK, dtype=torch.int32, device=self.device)[:1]     # [0] of a clip-scalar block
        line = torch.empty(width, dtype=torch.uint16, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_background(frame0.data_ptr(), height, width, bits, bg.data_ptr(),
                                                line.data_ptr(), self._stream()), "ff_background")
        self.launches += 1
        return bg, line

    def _fetch_frame0_stats_async(self, bg_dev: torch.Tensor, line_dev: torch.Tensor):
        """Copy the clip-scalar words (``bg_dev``: int32[1] or the whole int32[16] block) and the centre row
        to pinned host memory on a side stream that depends only on what is on the current stream so far,
        so later work on the main stream does not delay them.  Returns (event, scalars_host, line_host)."""
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(self.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        side = self._side_stream
        side.wait_event(ready)
        key = (bg_dev.numel(), line_dev.numel())
        if key not in self._pinned:       # reused: the previous values were consumed under a sync
            self._pinned[key] = (torch.empty(key[0], dtype=torch.int32).pin_memory(),
                                 torch.empty(key[1], dtype=torch.uint16).pin_memory())
        bg_host, line_host = self._pinned[key]
        with torch.cuda.stream(side):
            bg_host.copy_(bg_dev, non_blocking=True)
            line_host.copy_(line_dev, non_blocking=True)
            bg_dev.record_stream(side)
            line_dev.record_stream(side)
            done = torch.cuda.Event()
            done.record(side)
        return done, bg_host, line_host

    def clip_scalars(self, frame0: torch.Tensor, height: int, width: int, bits: int):
        """Background reduction + host-side float64 statistics.  Synchronises on two tiny D2H
        copies.  Returns ``(ClipScalars, bg_dev)``."""
        bg, line = self.background(frame0, height, width, bits)
        bg_host = int(bg.cpu().item())
        line_host = line.cpu().numpy()
        return ClipScalars.from_frame0_stats(bg_host, line_host), bg

    # ------------------------------------------------------------------ stages 2b-4
    def _workspace(self, n_bytes: int) -> torch.Tensor:
        """The zero-filled workspace of ``ff_process_range`` (every call leaves it zero-filled again)."""
        if self._ws is None or self._ws.numel() < n_bytes:
            self._ws = torch.zeros(max(n_bytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self._ws

    def process_range(self, frames: torch.Tensor, n_frames: int, height: int, width: int, bits: int,
                      params: DetectionParams, scalars: Optional[ClipScalars] = None,
                      bg_dev: Optional[torch.Tensor] = None, *, frame0: Optional[torch.Tensor] = None,
                      first_frame: int = 0, halo: Optional[torch.Tensor] = None,
                      skip: Optional[torch.Tensor] = None, diff_dtype: Optional[str] = None,
                      keep_profiles: bool = False, keep_decoded: bool = False,
                      first_exit: Optional[torch.Tensor] = None, truncate: bool = True,
                      partial: Optional[torch.Tensor] = None, pos_out: Optional[torch.Tensor] = None,
                      counts_out: Optional[torch.Tensor] = None, hooks: Optional[RangeHooks] = None,
                      init_first_exit: Optional[bool] = None, want_scalars: bool = True) -> RangeResult:
        """Stages 2-4 on a device-resident contiguous frame range in one C-ABI call (``ff_process_range``):
        prep kernel -> ONE range kernel for counts-only work (stream + count + empty-frame decision +
        detection + exit min + truncation), or stream kernel -> detect kernel when an image output is
        retained.  Nothing on the GPU waits for the host.

        Either pass ``scalars``/``bg_dev`` from :meth:`clip_scalars`, or pass ``frame0`` (the clip's first
        frame, packed, on the device; defaults to ``frames[0]`` when ``first_frame == 0``): then the prep
        kernel derives the background scalar - and, for the threshold method, the float64 flame threshold in
        NumPy's order of operations - on the device; the host repeats the statistics with NumPy from the
        copied-back centre row (``RangeResult.scalars``) and raises if the two thresholds ever differed.

        ``pos_out`` / ``counts_out`` / ``first_exit`` let the caller supply the int32 output arrays - e.g.
        the views of a ``sharding.RangeExchange`` block together with its ``hooks`` (the kernels then wait
        for the peers before writing and publish the finished block to them)."""
        self._check_dev(frames, "frames")
        for name, t in (("pos_out", pos_out), ("counts_out", counts_out)):
            if t is not None:
                self._check_dev(t, name)
                if t.dtype != torch.int32 or t.numel() < n_frames:
                    raise ValueError(f"{name} must be int32 with at least {n_frames} elements")
        fb = frame_nbytes(height, width, bits)
        if frames.dtype != torch.uint8 or frames.numel() < n_frames * fb:
            raise ValueError(f"frames must be uint8 with at least {n_frames * fb} bytes")
        if halo is not None:
            self._check_dev(halo, "halo")
            if halo.dtype != torch.uint8 or halo.numel() < fb:
                raise ValueError("halo must be a uint8 tensor holding one whole frame")
        if skip is not None:
            self._check_dev(skip, "skip")
            if skip.dtype != torch.uint8 or skip.numel() != n_frames:
                raise ValueError("skip must be uint8[n_frames]")
        if diff_dtype not in _DIFF_DTYPES:
            raise ValueError(f"diff_dtype must be one of {list(_DIFF_DTYPES)}")
        if params.method == "gradient" and width < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least 2 elements are required.")  # np.gradient's own error
        if (scalars is None) != (bg_dev is None):
            raise ValueError("pass scalars and bg_dev together (from clip_scalars), or neither")
        diff_code, diff_torch = _DIFF_DTYPES[diff_dtype]
        diff_thr = _clamp_i32(math.ceil(params.frame_diff_threshold))
        if diff_code == FF_DIFF_U16 and diff_thr < 0:
            raise ValueError("uint16 difference images need frame_diff_threshold >= 0")

        if scalars is None and params.method == "threshold" and width > MAX_DEVICE_STATS_WIDTH:
            # a centre row wider than the prep kernel's shared memory: statistics on the host first
            if frame0 is None:
                if first_frame != 0:
                    raise ValueError("frame0 (the clip's first frame) is required for a sub-range")
                frame0 = frames[:fb]
            scalars, bg_dev = self.clip_scalars(frame0, height, width, bits)
        line_dev = None
        if scalars is None:
            if frame0 is None:
                if first_frame != 0:
                    raise ValueError("frame0 (the clip's first frame) is required for a sub-range")
                frame0 = frames[:fb]
            self._check_dev(frame0, "frame0")
            if frame0.dtype != torch.uint8 or frame0.numel() < fb:
                raise ValueError("frame0 must be a uint8 tensor holding one whole frame")
            scal_dev = torch.empty(SCALAR_BLOCK, dtype=torch.int32, device=self.device)
            line_dev = torch.empty(width, dtype=torch.uint16, device=self.device) if want_scalars else None
            empty_thr, threshold_floor = -1, 0        # derived on the device (:1458, :1367-1370)
        else:
            self._check_dev(bg_dev, "bg_dev")
            scal_dev = bg_dev
            empty_thr = _clamp_i32(math.floor(scalars.noise_threshold))
            threshold_floor = _clamp_i32(math.floor(scalars.flame_threshold))

        ws_bytes, n_partial, fused = C.c_int64(0), C.c_int64(0), C.c_int(0)
        _cabi.check(self._lib.ff_process_range_plan(n_frames, height, width, bits, diff_code, int(keep_decoded),
                                                    int(keep_profiles), C.byref(ws_bytes), C.byref(n_partial),
                                                    C.byref(fused)), "ff_process_range_plan")
        ws = self._workspace(ws_bytes.value)
        if n_partial.value and (partial is None or partial.numel() < n_partial.value):
            partial = torch.empty(n_partial.value, dtype=torch.int32, device=self.device)
        pos = pos_out[:n_frames] if pos_out is not None else torch.empty(n_frames, dtype=torch.int32,
                                                                         device=self.device)
        counts = counts_out[:n_frames] if counts_out is not None else torch.empty(n_frames, dtype=torch.int32,
                                                                                  device=self.device)
        if init_first_exit is None:
            init_first_exit = first_exit is None
        if first_exit is None:
            first_exit = torch.empty(1, dtype=torch.int32, device=self.device)
        diff = None if diff_torch is None else torch.empty((n_frames, height, width), dtype=diff_torch,
                                                           device=self.device)
        profiles = torch.zeros((n_frames, width), dtype=torch.int32, device=self.device) if keep_profiles else None
        decoded = torch.empty((n_frames, height, width), dtype=torch.uint16,
                              device=self.device) if keep_decoded else None
        kb_min = min_signal_count(height * width, params.min_signal_fraction)
        args = RangeArgs(
            frames_dev=frames.data_ptr(), halo_dev=_ptr(halo), frame0_dev=None if scalars is not None else frame0.data_ptr(),
            n_frames=n_frames, first_frame=first_frame, height=height, width=width, bits=bits,
            method=FF_METHOD[params.method], use_frame_diff=int(params.use_frame_diff), min_run_px=params.min_run_px,
            exit_margin_px=params.exit_margin_px, diff_thr=diff_thr,
            grad2_bound=_clamp_i32(math.ceil(-2.0 * params.min_gradient_strength)), empty_thr=empty_thr,
            threshold_floor=threshold_floor, min_signal_count=kb_min, skip_dev=_ptr(skip),
            scalars_dev=scal_dev.data_ptr(), centerline_dev=_ptr(line_dev), pos_out_dev=pos.data_ptr(),
            count_out_dev=counts.data_ptr(), first_exit_dev=first_exit.data_ptr(),
            init_first_exit=int(bool(init_first_exit)), truncate=int(bool(truncate)), diff_out_dev=_ptr(diff),
            diff_dtype=diff_code, decoded_out_dev=_ptr(decoded), profile_out_dev=_ptr(profiles),
            partial_dev=_ptr(partial) if n_partial.value else None, workspace_dev=ws.data_ptr(),
            hooks=C.pointer(hooks) if hooks is not None else None)
        st = self._stream()
        with torch.cuda.device(self.device):
            ev = self._stream_events
            if ev is not None:            # bench hook: time every 4th call (a timing event between two kernels costs
                self._stream_event_tick += 1          # the stream a few microseconds, so not around every step)
                if self._stream_events_every > 1 and self._stream_event_tick % self._stream_events_every != 1:
                    ev = None
            if ev is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = self._lib.ff_process_range(C.byref(args), st)
            if rc != 0:
                self._ws = None           # a failed launch sequence may leave workspace words set
            _cabi.check(rc, "ff_process_range")
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        has_prep = scalars is None or init_first_exit or hooks is not None
        self.launches += (1 if has_prep else 0) + (1 if fused.value else 2)
        if scalars is None and want_scalars:
            scalars = PendingScalars(scal_dev, line_dev, params.method == "threshold")
        return RangeResult(first_frame, pos, counts, first_exit, diff, profiles, decoded, scalars)

    # ------------------------------------------------------------------ HEAD-parity detector
    def head_lines(self, frames: torch.Tensor, n_frames: int, height: int, width: int, bits: int, params, *,
                   frame0: Optional[torch.Tensor] = None, first_frame: int = 0,
                   halo: Optional[torch.Tensor] = None, skip: Optional[torch.Tensor] = None):
        """The image part of the HEAD detector on a device-resident frame range - everything that
        shards by frame range: streaming kernel (above-noise counts) -> ``ff_head_lines``
        (difference, 3x3 opening, Gaussian, Sobel / gradient on the centre band, float64 in SciPy's
        operation order).  Returns ``(lines float64[n,2,W], flags uint8[n], pending)``; pass
        ``pending`` to :meth:`head_scalars` after the remaining launches to get the ``ClipScalars``."""
        from .head import gaussian_weights
        self._check_dev(frames, "frames")
        fb = frame_nbytes(height, width, bits)
        if frames.dtype != torch.uint8 or frames.numel() < n_frames * fb:
            raise ValueError(f"frames must be uint8 with at least {n_frames * fb} bytes")
        if width < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least 2 elements are required.")
        if halo is not None:
            self._check_dev(halo, "halo")
        if skip is not None:
            self._check_dev(skip, "skip")
            if skip.dtype != torch.uint8 or skip.numel() != n_frames:
                raise ValueError("skip must be uint8[n_frames]")
        if frame0 is None:
            if first_frame != 0:
                raise ValueError("frame0 (the clip's first frame) is required for a sub-range")
            frame0 = frames[:fb]
        bg_dev, line_dev = self.background(frame0, height, width, bits)
        pending = self._fetch_frame0_stats_async(bg_dev, line_dev)
        diff_thr = _clamp_i32(math.ceil(params.frame_diff_threshold))
        n_elems, tiles = C.c_int64(0), C.c_int(0)
        _cabi.check(self._lib.ff_partial_len(n_frames, height, width, bits, C.byref(n_elems), C.byref(tiles)),
                    "ff_partial_len")
        partial = torch.empty(max(1, n_elems.value), dtype=torch.int32, device=self.device)
        lines = torch.empty((n_frames, 2, width), dtype=torch.float64, device=self.device)
        flags = torch.empty(n_frames, dtype=torch.uint8, device=self.device)
        scratch = torch.empty(n_frames + 4, dtype=torch.int32, device=self.device)
        weights = np.ascontiguousarray(gaussian_weights(params.gaussian_sigma), dtype=np.float64)
        radius = (weights.size - 1) // 2
        st = self._stream()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_stream_frames(
                frames.data_ptr(), _ptr(halo), n_frames, height, width, bits, bg_dev.data_ptr(), -1, diff_thr,
                _ptr(skip), partial.data_ptr(), None, FF_DIFF_NONE, None, st), "ff_stream_frames")
            _cabi.check(self._lib.ff_head_lines(
                frames.data_ptr(), _ptr(halo), n_frames, height, width, bits, bg_dev.data_ptr(),
                partial.data_ptr(), min_signal_count(height * width, params.min_signal_fraction), diff_thr,
                int(params.morphology_kernel_size), weights.ctypes.data_as(C.POINTER(C.c_double)), radius, _ptr(skip),
                lines.data_ptr(),
                flags.data_ptr(), scratch.data_ptr(), st), "ff_head_lines")
        self.launches += 3      # stream, flags, band
        return lines, flags, pending

    @staticmethod
    def head_scalars(pending) -> ClipScalars:
        """The clip's float64 frame-0 statistics once the side-stream copies of :meth:`head_lines` landed."""
        done, bg_host, line_host = pending
        done.synchronize()
        return ClipScalars.from_frame0_stats(int(bg_host.item()), line_host.numpy())

    def process_head(self, frames: torch.Tensor, n_frames: int, height: int, width: int, bits: int,
                     params, frame_rate: float, calibration: float, *, frame0: Optional[torch.Tensor] = None,
                     first_frame: int = 0, halo: Optional[torch.Tensor] = None,
                     skip: Optional[torch.Tensor] = None, keep_lines: bool = False,
                     tracker_state=(-1, -1)) -> "HeadRangeResult":
        """The detector the reference runs at HEAD (scripts/process_videos.py:350-465) on a
        device-resident frame range: :meth:`head_lines` (the image pipeline) -> ``ff_head_track``
        (velocity-constrained search window, candidate selection, exit stop; :meth:`head_track_lines`).
        ``params`` is a ``head.HeadParams``."""
        from .head import max_displacement_px
        lines, flags, pending = self.head_lines(frames, n_frames, height, width, bits, params, frame0=frame0,
                                                first_frame=first_frame, halo=halo, skip=skip)
        track, stop = self.head_track_lines(lines, flags, first_frame, width, params,
                                            max_displacement_px(frame_rate, calibration, params), tracker_state)
        return HeadRangeResult(first_frame, track, flags, stop, lines if keep_lines else None,
                               self.head_scalars(pending))

    # ------------------------------------------------------------------ frame-level operators (seam B3)
    _PX_TYPES = {torch.uint8: FF_PX_U8, torch.uint16: FF_PX_U16, torch.float64: FF_PX_F64}

    def _px_type(self, t: torch.Tensor, name: str) -> int:
        self._check_dev(t, name)
        if t.dtype not in self._PX_TYPES:
            raise TypeError(f"{name}: frame-level operators take uint8, uint16 or float64 frames, got {t.dtype}")
        return self._PX_TYPES[t.dtype]

    def frame_op(self, op: str, frames: Sequence[torch.Tensor], scalar: float) -> torch.Tensor:
        """Element-wise float64 operators of scripts/process_videos.py on decoded device frames:
        ``"subtract_background"`` (:670-674, one frame), ``"difference"`` (:677-701, current, prior),
        ``"three_difference"`` (:704-740, prev, curr, next).  Returns a float64 tensor of the frames' shape."""
        arity = {"subtract_background": 1, "difference": 2, "three_difference": 3}
        if op not in arity:
            raise ValueError(f"unknown frame operator {op!r}")
        if len(frames) != arity[op]:
            raise ValueError(f"{op} takes {arity[op]} frame(s)")
        px = self._px_type(frames[0], "frame")
        for f in frames[1:]:
            if self._px_type(f, "frame") != px or f.shape != frames[0].shape:
                raise ValueError("frames must share dtype and shape")
        n = frames[0].numel()
        out = torch.empty(frames[0].shape, dtype=torch.float64, device=self.device)
        if n == 0:
            return out
        st = self._stream()
        with torch.cuda.device(self.device):
            if op == "subtract_background":
                rc = self._lib.ff_frame_subtract_background(frames[0].data_ptr(), px, n, float(scalar), out.data_ptr(), st)
            elif op == "difference":
                rc = self._lib.ff_frame_difference(frames[0].data_ptr(), frames[1].data_ptr(), px, n, float(scalar),
                                                   out.data_ptr(), st)
            else:
                rc = self._lib.ff_frame_three_difference(frames[0].data_ptr(), frames[1].data_ptr(),
                                                         frames[2].data_ptr(), px, n, float(scalar), out.data_ptr(), st)
        _cabi.check(rc, f"ff_frame_{op}")
        self.launches += 1
        return out

    def frame_count_above(self, frame: torch.Tensor, threshold: float) -> int:
        """``np.sum(frame > threshold)`` of ``is_empty_frame`` (:759) on a decoded device frame."""
        px = self._px_type(frame, "frame")
        if frame.numel() == 0:
            return 0
        count = torch.empty(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_frame_count_above(frame.data_ptr(), px, frame.numel(), float(threshold),
                                                       count.data_ptr(), self._stream()), "ff_frame_count_above")
        self.launches += 1
        return int(count.item())

    HEAD_IMAGES = ("frame_subtracted", "frame_diff", "noise_removed", "blurred", "sobel_output", "gradient_output")

    def head_images(self, frames: torch.Tensor, n_frames: int, height: int, width: int, bits: int, background: int,
                    *, frame_diff_threshold: float = 5.0, morphology_kernel_size: int = 3,
                    gaussian_sigma: float = 1.5, halo: Optional[torch.Tensor] = None,
                    halo_background: Optional[int] = None, skip: Optional[torch.Tensor] = None,
                    want: Sequence[str] = HEAD_IMAGES) -> Dict[str, torch.Tensor]:
        """Full-frame intermediates of ``FlameDetector.detect`` (scripts/process_videos.py:380-413) for
        ``n_frames`` device-resident frames (``ff_head_images``): float64 ``[n,H,W]`` tensors named like
        the fields of ``FlameDetectionResult`` (:197-217), plus ``"state"`` uint8[n] (0 skipped, 1 all
        valid, 2 no prior frame: only ``frame_subtracted`` is meaningful) and ``"stack"``, the one
        tensor ``[len(want), n, H, W]`` they are views of."""
        from .head import gaussian_weights
        self._check_dev(frames, "frames")
        fb = frame_nbytes(height, width, bits)
        if frames.numel() * frames.element_size() < n_frames * fb:
            raise ValueError(f"frames must hold at least {n_frames * fb} bytes")
        unknown = [w for w in want if w not in self.HEAD_IMAGES]
        if unknown:
            raise ValueError(f"unknown image(s) {unknown}; options: {', '.join(self.HEAD_IMAGES)}")
        if frame_diff_threshold < 0:
            raise ValueError("frame_diff_threshold must be >= 0")
        if background < 0 or (halo_background is not None and halo_background < 0):
            raise ValueError("background scalars must be >= 0")
        if width < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least 2 elements are required.")
        if halo is not None:
            self._check_dev(halo, "halo")
        if skip is not None:
            self._check_dev(skip, "skip")
            if skip.dtype != torch.uint8 or skip.numel() != n_frames:
                raise ValueError("skip must be uint8[n_frames]")
        weights = np.ascontiguousarray(gaussian_weights(gaussian_sigma), dtype=np.float64)
        # one allocation for all requested images (``out["stack"]``, [len(want), n, H, W]): a caller that wants
        # them on the host needs a single device-to-host copy
        stack = torch.empty((len(want), n_frames, height, width), dtype=torch.float64, device=self.device)
        out = {name: stack[i] for i, name in enumerate(want)}
        state = torch.empty(n_frames, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_head_images(
                frames.data_ptr(), _ptr(halo), n_frames, height, width, bits, int(background),
                int(background if halo_background is None else halo_background),
                _clamp_i32(math.ceil(frame_diff_threshold)), int(morphology_kernel_size),
                weights.ctypes.data_as(C.POINTER(C.c_double)), (weights.size - 1) // 2, _ptr(skip),
                *(_ptr(out.get(name)) for name in self.HEAD_IMAGES), state.data_ptr(), self._stream()),
                "ff_head_images")
        self.launches += 1
        out["state"] = state
        out["stack"] = stack
        return out

    def head_track_lines(self, lines: torch.Tensor, flags: torch.Tensor, first_frame: int, width: int, params,
                         max_displacement: int, tracker_state=(-1, -1)):
        """``ff_head_track`` on caller-provided centre-row lines ``float64[n,2,W]`` (Sobel, gradient) and
        flags ``uint8[n]``: search window (:317-348), candidates (:420-465), exit stop.  Returns
        ``(track int32[n,5], stop int32[3])`` device tensors."""
        self._check_dev(lines, "lines")
        self._check_dev(flags, "flags")
        n = flags.numel()
        if lines.dtype != torch.float64 or lines.numel() != n * 2 * width or flags.dtype != torch.uint8:
            raise ValueError("lines must be float64[n,2,W] and flags uint8[n]")
        track = torch.empty((n, 5), dtype=torch.int32, device=self.device)
        stop = torch.empty(3, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_head_track(
                lines.data_ptr(), flags.data_ptr(), n, first_frame, width, params.edge_margin_px, max_displacement,
                params.search_window_px, float(params.min_gradient_strength),
                float(params.sobel_threshold_fraction), params.exit_margin_px, int(tracker_state[0]),
                int(tracker_state[1]), track.data_ptr(), stop.data_ptr(), self._track_scratch(n).data_ptr(),
                self._stream()), "ff_head_track")
        self.launches += 4 if n > 16 else 1      # full-width, speculative walk, fix-up, resolve | one plain walk
        return track, stop

    def _track_scratch(self, n_frames: int) -> torch.Tensor:
        """Scratch of ``ff_head_track`` (second result table + per-segment states of the chained
        speculation); fully written before it is read, so it needs no initialisation."""
        n_elems = C.c_int64(0)
        _cabi.check(self._lib.ff_head_track_scratch_len(n_frames, C.byref(n_elems)), "ff_head_track_scratch_len")
        return torch.empty(n_elems.value, dtype=torch.int32, device=self.device)

    def truncate(self, pos: torch.Tensor, first_frame: int, first_exit: torch.Tensor) -> None:
        """Mark frames at/after the (global) first exit frame as dropped (README.md:145-149)."""
        self._check_dev(pos, "pos")
        self._check_dev(first_exit, "first_exit")
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_truncate(pos.data_ptr(), pos.numel(), first_frame, first_exit.data_ptr(),
                                              self._stream()), "ff_truncate")
        self.launches += 1

    def merge_ranges(self, gathered: torch.Tensor, world: int, cap_frames: int, total: int,
                     pos_out: torch.Tensor, counts_out: Optional[torch.Tensor],
                     first_exit_out: torch.Tensor) -> None:
        """Finish a range-sharded clip from the all-gathered range blocks (``ff_merge_ranges``):
        global exit frame = min of the block headers, truncation against it, de-padding."""
        self._check_dev(gathered, "gathered")
        self._check_dev(pos_out, "pos_out")
        self._check_dev(first_exit_out, "first_exit_out")
        if gathered.dtype != torch.int32 or gathered.numel() < world * (4 + 2 * cap_frames):
            raise ValueError("gathered must hold world range blocks of int32")
        if pos_out.dtype != torch.int32 or pos_out.numel() < total:
            raise ValueError("pos_out must be int32[total]")
        if counts_out is not None:
            self._check_dev(counts_out, "counts_out")
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.ff_merge_ranges(gathered.data_ptr(), world, cap_frames, total, pos_out.data_ptr(),
                                                  _ptr(counts_out), first_exit_out.data_ptr(), self._stream()),
                        "ff_merge_ranges")
        self.launches += 1

    # ------------------------------------------------------------------ host-resident clips
    def _ctx(self) -> C.c_void_p:
        if self._host_ctx is None:
            ctx = C.c_void_p()
            _cabi.check(self._lib.ff_host_ctx_create(self.device.index, self._host_chunk_bytes, C.byref(ctx)),
                        "ff_host_ctx_create")
            if self._copy_threads is not None:
                _cabi.check(self._lib.ff_host_ctx_set_copy_threads(ctx, int(self._copy_threads)),
                            "ff_host_ctx_set_copy_threads")
            self._host_ctx = ctx
        return self._host_ctx

    def upload(self, host: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
        """Host bytes -> a uint8 device tensor (``ff_host_upload``): in place DMA for pinned memory,
        threaded pinned bounce buffers for pageable memory such as the memory-mapped .mraw file."""
        ptr, nbytes, _keep = _host_buffer(host)
        out = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if nbytes:
            torch.cuda.current_stream(self.device).synchronize()     # the copy runs on the context's own stream
            _cabi.check(self._lib.ff_host_upload(self._ctx(), ptr, out.data_ptr(), nbytes), "ff_host_upload")
        return out

    def process_host(self, frames: Union[np.ndarray, torch.Tensor], n_frames: int, height: int, width: int,
                     bits: int, params: DetectionParams, scalars: ClipScalars, *, first_frame: int = 0,
                     halo: Union[np.ndarray, torch.Tensor, None] = None,
                     skip: Optional[np.ndarray] = None, block=None, hooks: Optional[RangeHooks] = None,
                     to_host: bool = True) -> HostResult:
        """End-to-end form: frames live in host memory (pinned tensor or mmapped file); chunks are copied
        H2D double-buffered against the range kernel; blocks until the results are complete.

        ``block`` (a ``sharding.RangeBlock``) + ``hooks``: one rank's range of a range-sharded clip - the
        results are written into the block on the device and published to the peers, exit frames are
        shared between the ranks while they stream, and nothing at or behind the smallest exit frame ANY
        rank has seen is uploaded (``HostResult.bytes_uploaded``).  ``to_host=False`` then skips the
        copies of the rank-local arrays to the host."""
        fb = frame_nbytes(height, width, bits)
        src_ptr, src_bytes, _keep = _host_buffer(frames)
        if src_bytes < n_frames * fb:
            raise ValueError(f"frames hold {src_bytes} bytes, need {n_frames * fb}")
        halo_ptr = None
        if halo is not None:
            halo_ptr, hb, _keep_h = _host_buffer(halo)
            if hb < fb:
                raise ValueError("halo must hold one whole frame")
        skip_ptr = None
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
            if skip.size != n_frames:
                raise ValueError("skip must have n_frames entries")
            skip_ptr = skip.ctypes.data
        if params.method == "gradient" and width < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least 2 elements are required.")
        if block is None and not to_host:
            raise ValueError("to_host=False needs a block for the results")
        kb = derive_kernel_bounds(scalars, params, height * width)
        self._ctx()
        pos = np.empty(n_frames, dtype=np.int32) if to_host else None
        counts = np.empty(n_frames, dtype=np.int32) if to_host else None
        done, fexit, moved = C.c_int64(0), C.c_int32(FF_NO_EXIT), C.c_int64(0)
        args = HostArgs(
            frames_host=src_ptr, halo_host=halo_ptr, n_frames=n_frames, first_frame=first_frame, height=height,
            width=width, bits=bits, bg=int(scalars.background), empty_thr=kb.empty_thr,
            method=FF_METHOD[params.method], use_frame_diff=int(params.use_frame_diff), diff_thr=kb.diff_thr,
            threshold_floor=kb.threshold_floor, grad2_bound=kb.grad2_bound, min_run_px=params.min_run_px,
            exit_margin_px=params.exit_margin_px, min_signal_count=kb.min_signal_count, skip_host=skip_ptr,
            pos_out_host=None if pos is None else pos.ctypes.data,
            count_out_host=None if counts is None else counts.ctypes.data,
            pos_block_dev=None if block is None else block.pos.data_ptr(),
            count_block_dev=None if block is None else block.counts.data_ptr(),
            first_exit_block_dev=None if block is None else block.first_exit.data_ptr(),
            hooks=C.pointer(hooks) if hooks is not None else None,
            frames_done_out=C.pointer(done), first_exit_out=C.pointer(fexit), bytes_uploaded_out=C.pointer(moved))
        _cabi.check(self._lib.ff_process_host_range(self._host_ctx, C.byref(args)), "ff_process_host_range")
        chunk_frames = max(1, min(n_frames, self._host_chunk_bytes // fb))
        n_chunks = -(-int(done.value) // chunk_frames)
        self.launches += 1 + n_chunks + 1      # prep, one range kernel per chunk, truncate / publish
        return HostResult(first_frame, pos, counts, int(fexit.value), int(done.value), int(moved.value))


def _host_buffer(buf: Union[np.ndarray, torch.Tensor]):
    """(pointer, nbytes, keep-alive) of a contiguous host buffer."""
    if isinstance(buf, torch.Tensor):
        if buf.device.type != "cpu" or not buf.is_contiguous():
            raise ValueError("host frames must be a contiguous CPU tensor")
        return buf.data_ptr(), buf.numel() * buf.element_size(), buf
    arr = np.asarray(buf)
    if not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("host frames must be C-contiguous")
    return arr.ctypes.data, arr.nbytes, arr


_engines = {}


def get_engine(device: Union[int, str, torch.device, None] = None) -> FlameFrontEngine:
    """Process-wide engine per device (created on first use; raises without CUDA/the library)."""
    _cabi.load()
    if not torch.cuda.is_available():
        raise RuntimeError(
            "the flame-front path needs a CUDA device: libflamefront has no CPU implementation")
    if device is None:
        idx = torch.cuda.current_device()
    else:
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = FlameFrontEngine(idx)
    return _engines[idx]
