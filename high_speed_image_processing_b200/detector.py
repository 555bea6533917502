"""Frame-level operators of ``scripts/process_videos.py`` on the GPU (seam B3).

The driver loop of the reference is replaced as a whole by ``process_videos.process_video`` (one
pass over the packed clip).  A script that drives the reference frame by frame instead - its own
loop around ``subtract_scalar_background`` / ``is_empty_frame`` / ``FlameDetector.detect``
(:1441-1516) - finds the same names here, with the same arguments, results and error behaviour;
every computation on pixels runs in the sm_100a kernels behind the C-ABI (``ff_frame_*``,
``ff_head_images``, ``ff_head_track``; csrc/ff_frameops.cu, csrc/ff_head.cu).  There is no CPU
implementation: without the library or a CUDA device these raise.

What stays on the host is what the reference does in Python scalars: the velocity / DDT bookkeeping
(``head.VelocityBook``) and the optional spline estimator (``scipy.interpolate.UnivariateSpline``,
the reference's own call at :296-301; it never influences a position, :448,:464-465).
"""
from __future__ import annotations

import csv
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from .engine import FlameFrontEngine, get_engine
from .head import VelocityBook

_NATIVE = (np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.float64))


# --------------------------------------------------------------------------------------
# element-wise frame functions (:670-763)
# --------------------------------------------------------------------------------------
def _as_device_frames(engine: FlameFrontEngine, images) -> List[torch.Tensor]:
    """Decoded frames as contiguous device tensors of ONE dtype the kernels take (uint8, uint16 or
    float64).  Other dtypes, and operands of different dtypes, are widened to float64 first, which
    is exact - the reference's first step is ``astype(np.float64)`` (:672, :695-696, :729-731)."""
    arrs = [im.detach() if isinstance(im, torch.Tensor) else np.asarray(im) for im in images]
    names = {str(a.dtype).replace("torch.", "") for a in arrs}
    widen = len(names) > 1 or next(iter(names)) not in ("uint8", "uint16", "float64")
    out = []
    for a in arrs:
        if isinstance(a, torch.Tensor):
            t = a.to(engine.device)
            if widen and t.dtype != torch.float64:
                t = t.to(torch.float64)
        else:
            if widen and a.dtype != np.float64:
                a = a.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)
        out.append(t.contiguous())
    return out


def _like_input(result: torch.Tensor, image):
    return result if isinstance(image, torch.Tensor) else result.cpu().numpy()


def subtract_scalar_background(image, background_scalar: float, *, engine: Optional[FlameFrontEngine] = None):
    """Subtract scalar background; set negative values to zero (scripts/process_videos.py:670-674).
    NumPy array in, float64 NumPy array out (a CUDA tensor in gives a CUDA tensor out)."""
    eng = engine or get_engine()
    return _like_input(eng.frame_op("subtract_background", _as_device_frames(eng, [image]), background_scalar), image)


def subtract_prior_frame(current_frame, prior_frame, threshold: float = 0.0, *,
                         engine: Optional[FlameFrontEngine] = None):
    """``current - prior`` with differences below ``threshold`` zeroed (:677-701)."""
    eng = engine or get_engine()
    return _like_input(eng.frame_op("difference", _as_device_frames(eng, [current_frame, prior_frame]), threshold),
                       current_frame)


def three_frame_difference(frame_prev, frame_curr, frame_next, threshold: float = 0.0, *,
                           engine: Optional[FlameFrontEngine] = None):
    """``minimum(|curr - prev|, |next - curr|)`` with values below ``threshold`` zeroed (:704-740)."""
    eng = engine or get_engine()
    frames = _as_device_frames(eng, [frame_prev, frame_curr, frame_next])
    return _like_input(eng.frame_op("three_difference", frames, threshold), frame_curr)


def is_empty_frame(frame, noise_threshold: float = 50.0, min_signal_fraction: float = 0.001, *,
                   engine: Optional[FlameFrontEngine] = None) -> bool:
    """True if the fraction of pixels above ``noise_threshold`` is below ``min_signal_fraction``
    (:743-763).  The count runs on the GPU; the division and the comparison are the reference's
    float64 expressions (an empty array divides by zero there, and here)."""
    eng = engine or get_engine()
    if not isinstance(frame, torch.Tensor):
        dt = np.asarray(frame).dtype
        if dt.kind == "f" and dt.itemsize < 8:
            # NumPy compares a float32/float16 array with a Python float in the array's precision
            noise_threshold = float(dt.type(noise_threshold))
    dev = _as_device_frames(eng, [frame])[0]
    above_noise = eng.frame_count_above(dev, noise_threshold)
    total_pixels = dev.numel()
    signal_fraction = np.int64(above_noise) / total_pixels
    return bool(signal_fraction < min_signal_fraction)


def write_results(output_dict: dict, path: str) -> str:
    """Space-delimited text file with a header row of the keys (:766-780; host-side I/O)."""
    names = list(output_dict.keys())
    n_rows = len(list(output_dict.values())[0])
    with open(path, "w", newline="") as f:
        writer = csv.writer(f, delimiter=" ", skipinitialspace=True)
        writer.writerow(names)
        for i in range(n_rows):
            writer.writerow([output_dict[key][i] for key in names])
    return path


# --------------------------------------------------------------------------------------
# FlameDetector (:163-663)
# --------------------------------------------------------------------------------------
@dataclass
class FlameDetectorConfig:
    """Configuration for flame front detection - the reference's fields and defaults (:164-193)."""
    frame_diff_threshold: float = 5.0
    morphology_kernel_size: int = 3
    gaussian_sigma: float = 1.5
    min_gradient_strength: float = 10.0
    edge_margin_px: int = 10
    sobel_threshold_fraction: float = 0.1
    max_velocity_change_m_s: float = 200.0
    ddt_velocity_jump_m_s: float = 1250.0
    use_spline_estimator: bool = True
    spline_smoothing: float = 0.5
    min_points_for_spline: int = 5
    search_window_px: int = 100
    exit_margin_px: int = 15


@dataclass
class FlameDetectionResult:
    """Results from a single frame's flame detection (:196-217)."""
    frame_idx: int
    time_s: float
    frame_subtracted: Optional[np.ndarray]
    frame_diff: Optional[np.ndarray]
    noise_removed: Optional[np.ndarray]
    blurred: Optional[np.ndarray]
    sobel_output: Optional[np.ndarray]
    gradient_output: Optional[np.ndarray]
    pos_min_gradient: Optional[int]
    pos_rightmost_sobel: Optional[int]
    pos_spline_predicted: Optional[int]
    search_bounds: Optional[Tuple[int, int]]
    final_position: Optional[int]


class FlameDetector:
    """Flame front detection with velocity-constrained tracking - ``FlameDetector`` of the reference
    (scripts/process_videos.py:220-663), one GPU pass per ``detect`` call.

    ``intermediates`` chooses what the image fields of the results hold: ``"host"`` (default, like
    the reference: float64 NumPy arrays), ``"device"`` (CUDA tensors, no download) or ``"none"``
    (``frame_subtracted`` ... ``gradient_output`` are ``None``; positions only).
    ``keep_results=False`` stops the detector from retaining every result (the reference keeps all
    of them for plotting, :536)."""

    def __init__(self, config: FlameDetectorConfig, frame_rate: float, calibration_m_per_px: float, *,
                 engine: Optional[FlameFrontEngine] = None, intermediates: str = "host",
                 keep_results: bool = True):
        if intermediates not in ("host", "device", "none"):
            raise ValueError("intermediates must be 'host', 'device' or 'none'")
        self.config = config
        self.frame_rate = frame_rate
        self.calibration = calibration_m_per_px
        self._engine = engine or get_engine()
        self._intermediates = intermediates
        self._keep_results = keep_results
        self._book = VelocityBook(frame_rate, calibration_m_per_px, config.ddt_velocity_jump_m_s)
        self._prior_dev: Optional[torch.Tensor] = None     # uint16 [H,W] on the device
        self._prior_bg = 0                                 # background scalar _prior_dev still carries
        self._spline = None
        self._detection_results: List[FlameDetectionResult] = []
        self._max_displacement_px = self._compute_max_displacement()

    # ---- state the reference exposes ---------------------------------------------------------------
    @property
    def _position_history(self) -> List[Tuple[int, Optional[int]]]:
        return self._book.history

    @property
    def _velocity_history(self) -> List[Tuple[int, float, Optional[float], Optional[float]]]:
        return [tuple(e) for e in self._book.velocities]

    @property
    def _ddt_frame_idx(self) -> Optional[int]:
        return self._book.ddt_frame

    @property
    def _prior_frame(self) -> Optional[np.ndarray]:
        """Background-subtracted prior frame (float64), as the reference keeps it (:469)."""
        if self._prior_dev is None:
            return None
        return self._engine.frame_op("subtract_background", [self._prior_dev], float(self._prior_bg)).cpu().numpy()

    @_prior_frame.setter
    def _prior_frame(self, value) -> None:
        """The reference's driver assigns the background-subtracted frame here for frames it skips
        as empty (:1462).  The values must be what a camera frame minus a background can be:
        integers in [0, 65535]."""
        if value is None:
            self._prior_dev = None
            return
        arr = np.asarray(value.detach().cpu() if isinstance(value, torch.Tensor) else value)
        if arr.ndim != 2:
            raise ValueError("prior frame must be a 2-D image")
        with np.errstate(invalid="ignore"):
            as_u16 = arr.astype(np.uint16)
        if not np.array_equal(as_u16, arr):
            raise ValueError("prior frame must hold integer values in [0, 65535] "
                             "(a background-subtracted camera frame)")
        self._prior_dev = torch.from_numpy(np.ascontiguousarray(as_u16)).to(self._engine.device)
        self._prior_bg = 0

    def _compute_max_displacement(self) -> int:
        """Maximum allowed pixel displacement between frames (:270-276)."""
        if self.frame_rate <= 0 or self.calibration <= 0:
            return 1000
        dt = 1.0 / self.frame_rate
        max_displacement_m = self.config.max_velocity_change_m_s * dt
        return int(np.ceil(max_displacement_m / self.calibration)) + 1

    def reset(self) -> None:
        """Reset tracking state for a new video (:278-285)."""
        self._book.reset()
        self._detection_results.clear()
        self._prior_dev = None
        self._prior_bg = 0
        self._spline = None

    # ---- spline estimator (:287-315): host-side, SciPy's own routine ----------------------------------
    def _update_spline(self) -> None:
        valid = [(f, p) for f, p in self._book.history if p is not None]
        if len(valid) < self.config.min_points_for_spline:
            self._spline = None
            return
        frames = np.array([f for f, _ in valid])
        positions = np.array([p for _, p in valid])
        try:
            from scipy.interpolate import UnivariateSpline
            self._spline = UnivariateSpline(frames, positions, s=self.config.spline_smoothing * len(frames),
                                            k=min(3, len(frames) - 1))
        except Exception:
            self._spline = None

    def predict_with_spline(self, frame_idx: int) -> Optional[int]:
        """Predict position using the spline estimator (:306-315)."""
        if self._spline is None:
            return None
        try:
            return max(0, int(self._spline(frame_idx)))
        except Exception:
            return None

    def get_search_bounds(self, frame_idx: int, width: int) -> Tuple[int, int]:
        """Velocity-constrained search bounds for this frame (:317-348) - the host statement of what
        the tracker kernel evaluates; ``detect`` reports the kernel's."""
        margin = self.config.edge_margin_px
        last_frame_idx, last_position = self._book.last_detection()
        if last_position < 0:
            return (margin, width - margin)
        max_displacement = self._max_displacement_px * max(1, frame_idx - last_frame_idx)
        return (last_position,
                min(width - margin, last_position + max_displacement + self.config.search_window_px))

    # ---- detection ------------------------------------------------------------------------------------
    def detect(self, frame, frame_idx: int, background_scalar: float) -> FlameDetectionResult:
        """Main detection entry point with the full pipeline (:350-537): background subtraction,
        difference against the prior frame, k x k opening, Gaussian blur, Sobel and gradient
        (``ff_head_images``), then the windowed candidate search on the centre row (``ff_head_track``).

        ``frame`` is a raw camera frame ``[H,W]`` of dtype uint8 or uint16 (what ``video[i]`` returns);
        ``background_scalar`` must be a non-negative integer value, as ``float(np.max(video[0]))`` is."""
        eng = self._engine
        cfg = self.config
        if isinstance(frame, torch.Tensor):
            raw = frame.detach()
            if raw.dtype not in (torch.uint8, torch.uint16):
                raise TypeError(f"detect() takes uint8 or uint16 camera frames, got {raw.dtype}")
            if raw.dtype == torch.uint8:       # one 16-bit format for the frame and its prior
                raw = torch.from_numpy(raw.cpu().numpy().astype(np.uint16))
            raw = raw.to(eng.device)
        else:
            arr = np.asarray(frame)
            if arr.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
                raise TypeError(f"detect() takes uint8 or uint16 camera frames, got {arr.dtype}")
            raw = torch.from_numpy(np.ascontiguousarray(arr.astype(np.uint16, copy=False))).to(eng.device)
        if raw.ndim != 2:
            raise ValueError(f"detect() takes one 2-D frame, got shape {tuple(raw.shape)}")
        raw = raw.contiguous()
        bg = int(background_scalar)
        if bg != background_scalar or bg < 0:
            raise ValueError("background_scalar must be a non-negative integer value "
                             "(the maximum of a camera frame, :1357-1358)")
        if cfg.morphology_kernel_size < 1 or cfg.morphology_kernel_size % 2 == 0 or cfg.morphology_kernel_size > 7:
            raise ValueError("morphology_kernel_size must be odd and at most 7")
        height, width = int(raw.shape[0]), int(raw.shape[1])
        if self._prior_dev is not None and tuple(self._prior_dev.shape) != (height, width):
            raise ValueError(f"operands could not be broadcast together with shapes "
                             f"({height},{width}) {tuple(self._prior_dev.shape)}")
        center_row = height // 2
        time_s = frame_idx / self.frame_rate if self.frame_rate > 0 else 0

        keep = self._intermediates != "none"
        want = eng.HEAD_IMAGES if keep else ("sobel_output", "gradient_output")
        imgs = eng.head_images(raw, 1, height, width, 16, bg,
                               frame_diff_threshold=cfg.frame_diff_threshold,
                               morphology_kernel_size=cfg.morphology_kernel_size,
                               gaussian_sigma=cfg.gaussian_sigma, halo=self._prior_dev,
                               halo_background=self._prior_bg, want=want)
        has_prior = self._prior_dev is not None
        lines = torch.stack((imgs["sobel_output"][0, center_row], imgs["gradient_output"][0, center_row])).contiguous()
        flags = torch.full((1,), 1 if has_prior else 2, dtype=torch.uint8, device=eng.device)
        track, _ = eng.head_track_lines(lines.view(1, 2, width), flags, frame_idx, width, cfg,
                                        self._max_displacement_px, self._book.last_detection())
        final, pos_a, pos_b, s0, s1 = (int(v) for v in track[0].tolist())

        host_stack = imgs["stack"][:, 0].cpu().numpy() if self._intermediates == "host" else None   # one copy

        def image(name: str):
            if not keep or (name != "frame_subtracted" and not has_prior):
                return None
            return imgs[name][0] if host_stack is None else host_stack[want.index(name)]

        pos_spline_predicted = self.predict_with_spline(frame_idx) if cfg.use_spline_estimator else None
        final_position = final if final >= 0 else None

        # ---- update state (:467-516) ----
        self._book.update(frame_idx, final_position)
        self._prior_dev = raw
        self._prior_bg = bg
        self._update_spline()                  # also with use_spline_estimator off, as at :472

        result = FlameDetectionResult(
            frame_idx=frame_idx, time_s=time_s,
            frame_subtracted=image("frame_subtracted"), frame_diff=image("frame_diff"),
            noise_removed=image("noise_removed"), blurred=image("blurred"),
            sobel_output=image("sobel_output"), gradient_output=image("gradient_output"),
            pos_min_gradient=pos_a if pos_a >= 0 else None,
            pos_rightmost_sobel=pos_b if pos_b >= 0 else None,
            pos_spline_predicted=pos_spline_predicted,
            search_bounds=(s0, s1), final_position=final_position)
        if self._keep_results:
            self._detection_results.append(result)
        return result

    def _validate_position(self, candidate_position: int, frame_idx: int) -> Optional[int]:
        """Validate a position against the tracking constraints (:539-570; unused by ``detect``,
        as in the reference)."""
        last_frame_idx, last_position = self._book.last_detection()
        if last_position < 0:
            return candidate_position
        if candidate_position < last_position:
            return None
        frames_elapsed = frame_idx - last_frame_idx
        if frames_elapsed > 0:
            max_displacement = self._max_displacement_px * frames_elapsed
            if candidate_position - last_position > max_displacement:
                return last_position + max_displacement
        return candidate_position

    def get_spline_curve(self, frame_range: Optional[Tuple[int, int]] = None):
        """(frames, positions) of the fitted spline for plotting, or None (:572-600)."""
        if self._spline is None:
            return None
        valid = [(f, p) for f, p in self._book.history if p is not None]
        if not valid:
            return None
        if frame_range is None:
            f_min, f_max = min(f for f, _ in valid), max(f for f, _ in valid)
        else:
            f_min, f_max = frame_range
        frames = np.linspace(f_min, f_max, 100)
        try:
            return frames, self._spline(frames)
        except Exception:
            return None

    # ---- read-outs (:602-663) ---------------------------------------------------------------------------
    @property
    def position_history(self) -> List[Tuple[int, Optional[int]]]:
        return self._book.history

    @property
    def last_position(self) -> Optional[int]:
        pos = self._book.last_detection()[1]
        return pos if pos >= 0 else None

    @property
    def last_velocity(self) -> Optional[float]:
        """Last computed velocity (first-order backward) in m/s."""
        return self._book.velocities[-1][1] if self._book.velocities else None

    @property
    def last_velocities(self) -> Tuple[Optional[float], Optional[float], Optional[float]]:
        if self._book.velocities:
            e = self._book.velocities[-1]
            return (e[1], e[2], e[3])
        return (None, None, None)

    @property
    def ddt_frame(self) -> Optional[int]:
        return self._book.ddt_frame

    @property
    def ddt_detected(self) -> bool:
        return self._book.ddt_frame is not None

    def get_velocity_history(self) -> List[Tuple[int, float, Optional[float], Optional[float]]]:
        return [tuple(e) for e in self._book.velocities]

    def get_pre_ddt_velocities(self) -> List[Tuple[int, float, Optional[float], Optional[float]]]:
        ddt = self._book.ddt_frame
        return [tuple(e) for e in self._book.velocities if ddt is None or e[0] < ddt]

    def get_post_ddt_velocities(self) -> List[Tuple[int, float, Optional[float], Optional[float]]]:
        ddt = self._book.ddt_frame
        if ddt is None:
            return []
        return [tuple(e) for e in self._book.velocities if e[0] >= ddt]

    def clear_last_central_difference(self) -> None:
        """Clear the central difference of the second-to-last velocity entry: it used the position
        of the frame on which the flame left the domain (:654-663)."""
        self._book.clear_last_central()
