"""HEAD-parity detector: host side (SURVEY.md section 8 f1 + f2).

The image pipeline and the windowed position search run on the GPU (``ff_head_lines`` /
``ff_head_track``, csrc/ff_head.cu).  What is left for the host is O(detections) scalar
bookkeeping that the reference does in float64 Python: the Gaussian taps, the per-frame
displacement bound, the three velocity estimates, DDT detection and the two stop rules
(scripts/process_videos.py:270-276, :474-516, :1486-1509, :654-663) - written with the
reference's own expressions so every float is bit-identical.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np


@dataclass
class HeadParams:
    """``FlameDetectorConfig`` (scripts/process_videos.py:164-193); same names and defaults.
    The spline estimator fields are omitted: it never influences a position (:448, :464-465)."""
    frame_diff_threshold: float = 5.0
    morphology_kernel_size: int = 3
    gaussian_sigma: float = 1.5
    min_gradient_strength: float = 10.0
    edge_margin_px: int = 10
    sobel_threshold_fraction: float = 0.1
    max_velocity_change_m_s: float = 200.0
    ddt_velocity_jump_m_s: float = 1250.0
    search_window_px: int = 100
    exit_margin_px: int = 15
    min_signal_fraction: float = 0.0005      # is_empty_frame call at :1459

    def __post_init__(self) -> None:
        if self.morphology_kernel_size not in (1, 3, 5, 7):
            raise ValueError("morphology_kernel_size must be odd and at most 7 (1, 3, 5 or 7)")
        if self.frame_diff_threshold < 0:
            raise ValueError("frame_diff_threshold must be >= 0")


def gaussian_weights(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The taps scipy.ndimage.gaussian_filter uses: radius int(truncate*sigma+0.5), normalised
    exp(-x^2/(2 sigma^2)) - the same NumPy expressions, so the same float64 values."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


def max_displacement_px(frame_rate: float, calibration: float, params: HeadParams) -> int:
    """scripts/process_videos.py:270-276."""
    if frame_rate <= 0 or calibration <= 0:
        return 1000
    dt = 1.0 / frame_rate
    max_displacement_m = params.max_velocity_change_m_s * dt
    return int(np.ceil(max_displacement_m / calibration)) + 1


VelocityEntry = List  # [frame_idx, v_backward1, v_backward2, v_central]


@dataclass
class HeadTrackSummary:
    rows: List[Tuple[int, float, int, float, bool]] = field(default_factory=list)
    per_frame: List[dict] = field(default_factory=list)
    velocity_history: List[VelocityEntry] = field(default_factory=list)
    ddt_frame: Optional[int] = None
    stop: Optional[Tuple[str, int]] = None       # ("exit" | "velocity_drop", frame)


class VelocityBook:
    """The detector's position / velocity / DDT bookkeeping (scripts/process_videos.py:467-516),
    one ``update`` per detect() call.  Entries are ``[frame_idx, v_backward1, v_backward2, v_central]``:
    first-order backward, second-order backward, and the second-order central difference that is
    filled into the PREVIOUS entry once the next position is known."""

    def __init__(self, frame_rate: float, calibration: float, ddt_velocity_jump_m_s: float):
        self.frame_rate = frame_rate
        self.calibration = calibration
        self.ddt_jump = ddt_velocity_jump_m_s
        self.history: List[Tuple[int, Optional[int]]] = []
        self.velocities: List[VelocityEntry] = []
        self.ddt_frame: Optional[int] = None

    def reset(self) -> None:
        self.history.clear()
        self.velocities.clear()
        self.ddt_frame = None

    def last_detection(self) -> Tuple[int, int]:
        """(frame, position) of the latest frame with a position, or (-1, -1)."""
        for f_idx, pos in reversed(self.history):
            if pos is not None:
                return f_idx, pos
        return -1, -1

    def update(self, frame_idx: int, final_position: Optional[int]) -> None:
        history, vel, calibration = self.history, self.velocities, self.calibration
        history.append((frame_idx, final_position))
        if final_position is None or len(history) < 2:
            return
        curr_frame, curr_pos = history[-1]
        prev_frame, prev_pos = history[-2]
        if prev_pos is None or not self.frame_rate > 0:
            return
        dt = (curr_frame - prev_frame) / self.frame_rate
        if not dt > 0:
            return
        v_backward1 = (curr_pos - prev_pos) * calibration / dt
        v_backward2 = None
        if len(history) >= 3:
            _, prev2_pos = history[-3]
            if prev2_pos is not None:
                v_backward2 = (3 * curr_pos - 4 * prev_pos + prev2_pos) * calibration / (2 * dt)
                v_central = (curr_pos - prev2_pos) * calibration / (2 * dt)
                if len(vel) >= 1:
                    old = vel[-1]
                    vel[-1] = [old[0], old[1], old[2], v_central]
        vel.append([frame_idx, v_backward1, v_backward2, None])
        if self.ddt_frame is None and len(vel) >= 2:
            if v_backward1 - vel[-2][1] > self.ddt_jump:
                self.ddt_frame = frame_idx

    def clear_last_central(self) -> None:
        _clear_last_central(self.velocities)


def finish_head_track(track: np.ndarray, flags: np.ndarray, first_frame: int, width: int, frame_rate: float,
                      calibration: float, offset: float, time_of: Callable[[int], float],
                      params: HeadParams) -> HeadTrackSummary:
    """Replay the reference's per-frame bookkeeping over the GPU tracker's output.

    ``track`` int32[n,5] = (final, pos_min_gradient, pos_rightmost_sobel, search_start, search_end),
    ``flags`` uint8[n] (0 = frame never reached the detector)."""
    out = HeadTrackSummary()
    book = VelocityBook(frame_rate, calibration, params.ddt_velocity_jump_m_s)
    vel = out.velocity_history = book.velocities
    active = np.nonzero(flags)[0]
    for i, (final, pos_a, pos_b, s0, s1) in zip(active.tolist(), track[active].tolist()):   # Python ints, one conversion
        frame_idx = first_frame + i
        if s0 < 0 and s1 < 0 and final < 0 and pos_a < 0:
            break                                   # the device tracker stopped before this frame
        final_position = final if final >= 0 else None
        out.per_frame.append({"frame": frame_idx, "final": final_position,
                              "min_gradient": pos_a if pos_a >= 0 else None,
                              "rightmost_sobel": pos_b if pos_b >= 0 else None, "search": [s0, s1]})
        book.update(frame_idx, final_position)          # velocities and DDT (:479-516)
        out.ddt_frame = book.ddt_frame
        velocity = vel[-1][1] if vel else None
        # ---- stop rules (:1486-1509); the stopping frame is not recorded ----
        if final_position is not None and final_position >= width - params.exit_margin_px:
            book.clear_last_central()
            out.stop = ("exit", frame_idx)
            break
        if velocity is not None and len(vel) >= 2:
            prev_v1 = vel[-2][1]
            if prev_v1 is not None and prev_v1 > 100:
                if (prev_v1 - velocity) / prev_v1 > 0.5:
                    book.clear_last_central()
                    out.stop = ("velocity_drop", frame_idx)
                    break
        if final_position is not None:
            pos_m = final_position * calibration + offset                       # :1512
            is_post_ddt = out.ddt_frame is not None and frame_idx >= out.ddt_frame
            out.rows.append((frame_idx, time_of(frame_idx), final_position, pos_m, is_post_ddt))
    return out


def _clear_last_central(vel: List[VelocityEntry]) -> None:
    """:654-663 - the central difference of the previous entry used the rejected position."""
    if len(vel) >= 2:
        e = vel[-2]
        vel[-2] = [e[0], e[1], e[2], None]
