"""NumPy restatement of the reference's per-frame flame-front path.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  The product never imports it.

All ``file:line`` citations are into ``/root/reference``.  Every function below follows
the NumPy expression the reference uses (same dtype promotion, same order of operations)
so that integer-valued float64 intermediates are bit-identical to the reference's.

Parity status
-------------
* decode (``unpack12`` / ``frames_from_bytes``): the arithmetic lives in the third-party
  ``pyMRAW`` package (PyPI, upstream ladisk/pyMRAW, version NOT pinned by the reference -
  ``README.md:29`` is the only mention) which is absent from ``/root/reference`` and from
  this image.  Restated from its published layout.  **parity unpinned**.
* ``subtract_scalar_background``, ``is_empty_frame``, ``frame_difference``,
  ``background_scalar``, ``centerline_stats``, ``frame_time_*``, ``position_m``,
  ``FileCalibration`` matching, ``distribute_indices``: pinned against the reference's own
  functions executed in this container (``oracle/make_golden.py`` ->
  ``tests/golden/``), and against the README sample rows (``README.md:93-96``).
* ``detect_gradient``: follows HEAD's "Method A" (``scripts/process_videos.py:413,
  427-430``) over the full width.  Pinned through ``np.gradient``/``np.argmin`` only.
* ``detect_threshold`` / ``detect_half_maximum``: the reference has NO code for these
  (only prose at ``README.md:134-138``).  The spec frozen here is SURVEY.md section 8(c).
  **parity unpinned**.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

METHODS = ("threshold", "gradient", "half_maximum")


# --------------------------------------------------------------------------------------
# decode  (pyMRAW.load_video, called at src/photron/video.py:332)
# --------------------------------------------------------------------------------------
def unpack12(buf: np.ndarray) -> np.ndarray:
    """Packed 12-bit -> uint16.  Three bytes ``b0 b1 b2`` hold two pixels:
    ``p0 = (b0 << 4) | (b1 >> 4)``, ``p1 = ((b1 & 15) << 8) | b2`` (pyMRAW's 12-bit
    reader; the whole file is treated as ONE flat byte stream, then reshaped)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8).ravel()
    if buf.size % 3 != 0:
        raise ValueError("packed 12-bit stream length must be a multiple of 3 bytes")
    trip = buf.reshape(-1, 3).astype(np.uint16)
    out = np.empty(trip.shape[0] * 2, dtype=np.uint16)
    out[0::2] = (trip[:, 0] << 4) | (trip[:, 1] >> 4)
    out[1::2] = ((trip[:, 1] & 15) << 8) | trip[:, 2]
    return out


def pack12(pixels: np.ndarray) -> np.ndarray:
    """Inverse of :func:`unpack12` (used by the synthetic writer's checks)."""
    px = np.ascontiguousarray(pixels, dtype=np.uint16).ravel()
    if px.size % 2 != 0:
        raise ValueError("need an even number of pixels to pack 12-bit")
    if px.size and int(px.max()) > 0xFFF:
        raise ValueError("pixel value exceeds 12 bits")
    p0 = px[0::2]
    p1 = px[1::2]
    out = np.empty((p0.size, 3), dtype=np.uint8)
    out[:, 0] = p0 >> 4
    out[:, 1] = ((p0 & 15) << 4) | (p1 >> 8)
    out[:, 2] = p1 & 255
    return out.ravel()


def frames_from_bytes(raw: np.ndarray, n: int, h: int, w: int, bits: int) -> np.ndarray:
    """``images[N,H,W]`` as pyMRAW returns them: uint8 for 8-bit, little-endian uint16 for
    16-bit, unpacked uint16 for packed 12-bit (selector is the CIH 'Color Bit')."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).ravel()
    if bits == 8:
        return raw[: n * h * w].reshape(n, h, w)
    if bits == 16:
        return raw[: n * h * w * 2].view("<u2").reshape(n, h, w)
    if bits == 12:
        nbytes = n * h * w * 3 // 2
        return unpack12(raw[:nbytes]).reshape(n, h, w)
    raise ValueError(f"unsupported bit depth {bits}")


# --------------------------------------------------------------------------------------
# per-frame primitives
# --------------------------------------------------------------------------------------
def subtract_scalar_background(image: np.ndarray, background_scalar: float) -> np.ndarray:
    """scripts/process_videos.py:670-674."""
    sub = image.astype(np.float64) - background_scalar
    sub[sub < 0] = 0
    return sub


def is_empty_frame(frame: np.ndarray, noise_threshold: float = 50.0,
                   min_signal_fraction: float = 0.001) -> bool:
    """scripts/process_videos.py:743-763."""
    above = np.sum(frame > noise_threshold)
    return bool(above / frame.size < min_signal_fraction)


def nonempty_count(frame: np.ndarray, noise_threshold: float) -> int:
    """The integer that ``is_empty_frame`` divides (``:759``)."""
    return int(np.sum(frame > noise_threshold))


def frame_difference(current: np.ndarray, prior: np.ndarray, threshold: float) -> np.ndarray:
    """scripts/process_videos.py:397-399 (and ``subtract_prior_frame`` ``:677-701``)."""
    diff = current.astype(np.float64) - prior.astype(np.float64)
    diff[diff < threshold] = 0
    return diff


def three_frame_difference(frame_prev: np.ndarray, frame_curr: np.ndarray, frame_next: np.ndarray,
                           threshold: float = 0.0) -> np.ndarray:
    """scripts/process_videos.py:704-740."""
    prev, curr, nxt = (f.astype(np.float64) for f in (frame_prev, frame_curr, frame_next))
    motion = np.minimum(np.abs(curr - prev), np.abs(nxt - curr))
    motion[motion < threshold] = 0
    return motion


def background_scalar(frame0: np.ndarray) -> float:
    """scripts/process_videos.py:1357-1358."""
    return float(np.max(frame0))


def centerline_stats(frame0: np.ndarray) -> Tuple[float, float, float, float]:
    """scripts/process_videos.py:1361-1370 -> (mean, std, max, flame_threshold)."""
    row = frame0.shape[0] // 2
    line = frame0[row, :].astype(np.float64)
    mean = np.mean(line)
    std = np.std(line)
    mx = np.max(line)
    thr = max(mean + 5 * std, mx * 2.0)
    return float(mean), float(std), float(mx), float(thr)


def empty_noise_threshold(bg: float) -> float:
    """scripts/process_videos.py:1458."""
    return max(10.0, bg * 0.5)


# --------------------------------------------------------------------------------------
# detection methods on a centre-row profile
# --------------------------------------------------------------------------------------
def detect_gradient(profile: np.ndarray, min_gradient_strength: float = 10.0) -> Optional[int]:
    """HEAD Method A over the full width (scripts/process_videos.py:413,427-430):
    ``g = np.gradient(p)``; first arg-min; valid iff ``g[i] < -min_gradient_strength``."""
    p = np.asarray(profile, dtype=np.float64)
    g = np.gradient(p)
    if np.min(g) < -min_gradient_strength:
        return int(np.argmin(g))
    return None


def detect_threshold(profile: np.ndarray, threshold: float, min_run_px: int = 1) -> Optional[int]:
    """README.md:134-135 ("rightmost edge of contiguous high-intensity regions").
    Frozen spec: ``m = p > T`` (strict, like ``:759``); answer = right edge of
    the rightmost run of consecutive True whose length >= ``min_run_px``."""
    m = np.asarray(profile, dtype=np.float64) > threshold
    w = m.size
    i = w - 1
    while i >= 0:
        if not m[i]:
            i -= 1
            continue
        end = i
        while i >= 0 and m[i]:
            i -= 1
        if end - i >= min_run_px:
            return int(end)
    return None


def detect_half_maximum(profile: np.ndarray) -> Optional[int]:
    """README.md:137-138 ("drops to 50% of peak value on the falling edge").
    Frozen spec: ``k = argmax(p)`` (first); None if ``p[k] <= 0``;
    ``h = 0.5*p[k]``; answer = smallest ``i > k`` with ``p[i] < h``; None if there is none."""
    p = np.asarray(profile, dtype=np.float64)
    k = int(np.argmax(p))
    peak = p[k]
    if peak <= 0:
        return None
    half = 0.5 * peak
    below = np.nonzero(p[k + 1:] < half)[0]
    if below.size == 0:
        return None
    return int(k + 1 + below[0])


def detect(profile: np.ndarray, method: str, *, threshold: float = 0.0,
           min_gradient_strength: float = 10.0, min_run_px: int = 1) -> Optional[int]:
    if method == "threshold":
        return detect_threshold(profile, threshold, min_run_px)
    if method == "gradient":
        return detect_gradient(profile, min_gradient_strength)
    if method == "half_maximum":
        return detect_half_maximum(profile)
    raise ValueError(f"unknown detection_method {method!r}")


# --------------------------------------------------------------------------------------
# timing / calibration scalars
# --------------------------------------------------------------------------------------
def frame_time_absolute(idx: int, start_frame: int, skip_frame: int, rate: int) -> float:
    """src/photron/video.py:235-241."""
    if rate <= 0:
        return 0.0
    return (start_frame + (idx * skip_frame)) / rate


def frame_time_relative(idx: int, trigger_frame: int, rate: int) -> float:
    """src/photron/video.py:218-220."""
    if rate <= 0:
        return 0.0
    return (idx - trigger_frame) / rate


def position_m(px: int, calibration: float, offset: float) -> float:
    """scripts/process_videos.py:1512 (unfused multiply then add)."""
    return px * calibration + offset


# --------------------------------------------------------------------------------------
# the frame loop  (scripts/process_videos.py:1441-1516, README.md:143-149 for exit)
# --------------------------------------------------------------------------------------
@dataclass
class ClipParams:
    method: str = "half_maximum"
    use_frame_diff: bool = True            # VideoSourceConfig.use_frame_diff  (:112)
    frame_diff_threshold: float = 5.0      # FlameDetectorConfig.frame_diff_threshold (:169)
    min_gradient_strength: float = 10.0    # (:174)
    min_run_px: int = 1
    exit_margin_px: int = 10               # README.md:146 ; HEAD uses 15 (:193)
    min_signal_fraction: float = 0.0005    # (:1459)
    skip_frames: Sequence[int] = ()
    keep_profiles: bool = False
    keep_diffs: bool = False


@dataclass
class ClipResult:
    background: float
    centerline_mean: float
    centerline_std: float
    centerline_max: float
    flame_threshold: float
    noise_threshold: float
    pos_px: np.ndarray                     # int32[N]; -1 = None / skipped / empty
    nonempty: np.ndarray                   # int64[N]; count of pixels above noise threshold
    empty: np.ndarray                      # bool[N]
    first_exit: int                        # first frame with pos >= W - margin, or N
    records: List[Tuple[int, int]] = field(default_factory=list)   # (frame, px), truncated
    profiles: Optional[np.ndarray] = None  # float64[N,W]
    diffs: Optional[np.ndarray] = None     # float64[N,H,W]


def process_clip(frames: np.ndarray, params: ClipParams, *, frame0: Optional[np.ndarray] = None,
                 first_index: int = 0, prior_frame: Optional[np.ndarray] = None) -> ClipResult:
    """Run the serial reference loop over ``frames[N,H,W]``.

    ``frame0`` (default ``frames[0]``) supplies the per-clip background statistics;
    ``first_index`` is the clip-global index of ``frames[0]`` and ``prior_frame`` the raw
    frame preceding it (both used only to evaluate a contiguous sub-range exactly as the
    serial loop would - the multi-GPU sharding tests rely on this).
    """
    n, h, w = frames.shape
    f0 = frames[0] if frame0 is None else frame0
    bg = background_scalar(f0)
    c_mean, c_std, c_max, flame_thr = centerline_stats(f0)
    noise_thr = empty_noise_threshold(bg)
    row = h // 2
    skip = set(int(s) for s in params.skip_frames)

    pos = np.full(n, -1, dtype=np.int32)
    nonempty = np.zeros(n, dtype=np.int64)
    empty = np.zeros(n, dtype=bool)
    profiles = np.zeros((n, w), dtype=np.float64) if params.keep_profiles else None
    diffs = np.zeros((n, h, w), dtype=np.float64) if params.keep_diffs else None

    prior = None
    if prior_frame is not None:
        prior = subtract_scalar_background(prior_frame, bg)

    for i in range(n):
        gidx = first_index + i
        if gidx in skip:                                   # :1443-1445 (prior untouched)
            continue
        sub = subtract_scalar_background(frames[i], bg)    # :1455
        cnt = nonempty_count(sub, noise_thr)
        nonempty[i] = cnt
        is_empty = bool(cnt / sub.size < params.min_signal_fraction)   # :761-763
        empty[i] = is_empty
        diff = None
        # the reference never builds the difference image of a skipped-empty frame (:1459-1463
        # `continue`s before detect); it is only materialised here when the caller keeps it
        want_diff = params.keep_diffs or (params.use_frame_diff and (not is_empty or params.keep_profiles))
        if prior is not None and want_diff:
            diff = frame_difference(sub, prior, params.frame_diff_threshold)   # :397-399
            if diffs is not None:
                diffs[i] = diff
        if params.use_frame_diff:
            profile = None if diff is None else diff[row, :]
        else:
            profile = sub[row, :]
        if profile is not None and profiles is not None:
            profiles[i] = profile
        if not is_empty and profile is not None:           # :1459-1463 skip, :397 no prior
            p = detect(profile, params.method, threshold=flame_thr,
                       min_gradient_strength=params.min_gradient_strength,
                       min_run_px=params.min_run_px)
            if p is not None:
                pos[i] = p
        prior = sub                                        # :469 and :1462

    exits = np.nonzero(pos >= w - params.exit_margin_px)[0]            # :1488-1494
    first_exit = int(exits[0]) if exits.size else n
    records = [(first_index + i, int(pos[i])) for i in range(first_exit) if pos[i] >= 0]
    return ClipResult(bg, c_mean, c_std, c_max, flame_thr, noise_thr, pos, nonempty, empty,
                      first_exit, records, profiles, diffs)
