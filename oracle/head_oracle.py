"""Oracle for the HEAD detector: what the reference's
``FlameDetector.detect`` actually executes (scripts/process_videos.py:317-516).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Two layers:

* ``detect_lines_scipy`` - the reference's own sequence of SciPy calls
  (``grey_opening`` 3x3 -> ``gaussian_filter`` sigma 1.5 -> ``sobel(axis=1)`` and
  ``np.gradient(axis=1)``, :403-413) on the full difference image; centre row returned.
* ``detect_lines_restated`` - the same arithmetic written out operation by operation on the
  19-row centre band only, in the exact floating-point order of SciPy's ``correlate1d``
  (``NI_Correlate1D``: centre tap first, then symmetric pairs from the outermost inwards,
  ``tmp += (in[-j] + in[+j]) * w[j]``; no fused multiply-add).  This is the specification the
  CUDA band kernel implements; ``tests/test_oracle_head.py`` asserts it is bit-identical to
  the SciPy layer, and ``oracle/make_golden.py`` pins both to the reference's detector.

``HeadTracker`` restates the sequential part: search bounds (:317-348), candidate selection
(:424-465), velocities and DDT (:474-516), exit and velocity-drop stops (:1486-1509).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import flame_oracle as fo


# --------------------------------------------------------------------------------------
# layer 1: the reference's SciPy calls
# --------------------------------------------------------------------------------------
def detect_lines_scipy(frame_diff: np.ndarray, kernel_size: int = 3, sigma: float = 1.5):
    """(sobel centre row, gradient centre row) exactly as :403-418."""
    from scipy.ndimage import gaussian_filter, grey_opening, sobel
    noise_removed = grey_opening(frame_diff, size=(kernel_size, kernel_size))
    blurred = gaussian_filter(noise_removed, sigma=sigma)
    sobel_output = sobel(blurred, axis=1)
    gradient_output = np.gradient(blurred, axis=1)
    row = frame_diff.shape[0] // 2
    return sobel_output[row, :], gradient_output[row, :]


def detect_images_scipy(frame_sub: np.ndarray, prior_sub: Optional[np.ndarray], threshold: float = 5.0,
                        kernel_size: int = 3, sigma: float = 1.5) -> dict:
    """Every full-frame intermediate of FlameDetector.detect (:380-413), named like the fields of
    FlameDetectionResult (:197-217); the ones after ``frame_subtracted`` are None without a prior."""
    from scipy.ndimage import gaussian_filter, grey_opening, sobel
    out = {"frame_subtracted": frame_sub, "frame_diff": None, "noise_removed": None, "blurred": None,
           "sobel_output": None, "gradient_output": None}
    if prior_sub is None:
        return out
    out["frame_diff"] = fo.frame_difference(frame_sub, prior_sub, threshold)
    out["noise_removed"] = grey_opening(out["frame_diff"], size=(kernel_size, kernel_size))
    out["blurred"] = gaussian_filter(out["noise_removed"], sigma=sigma)
    out["sobel_output"] = sobel(out["blurred"], axis=1)
    out["gradient_output"] = np.gradient(out["blurred"], axis=1)
    return out


class FrameDetectorOracle:
    """FlameDetector driven frame by frame (:350-537) without the driver loop around it: the
    restated search bounds / candidate selection / velocity bookkeeping on the SciPy images."""

    def __init__(self, frame_rate: float, calibration: float, cfg: Optional["HeadConfig"] = None,
                 kernel_size: int = 3):
        self.cfg = cfg or HeadConfig()
        self.kernel_size = kernel_size
        self.frame_rate, self.calibration = frame_rate, calibration
        self.maxdisp = max_displacement_px(frame_rate, calibration, self.cfg)
        self.history: List[Tuple[int, Optional[int]]] = []
        self.velocities: List[list] = []
        self.ddt_frame: Optional[int] = None
        self.prior: Optional[np.ndarray] = None

    def detect(self, frame: np.ndarray, frame_idx: int, background: float) -> dict:
        cfg = self.cfg
        w = frame.shape[1]
        sub = fo.subtract_scalar_background(frame, background)
        last = next(((f, p) for f, p in reversed(self.history) if p is not None), None)
        if last is None:
            s0, s1 = cfg.edge_margin_px, w - cfg.edge_margin_px
        else:
            lf, lp = last
            s0, s1 = lp, min(w - cfg.edge_margin_px, lp + self.maxdisp * max(1, frame_idx - lf) + cfg.search_window_px)
        imgs = detect_images_scipy(sub, self.prior, cfg.frame_diff_threshold, self.kernel_size, cfg.gaussian_sigma)
        pa = pb = final = None
        if self.prior is not None:
            row = frame.shape[0] // 2
            pa, pb, final = select_position(imgs["sobel_output"][row, :], imgs["gradient_output"][row, :], s0, s1, cfg)
        self.history.append((frame_idx, final))
        self.prior = sub
        self.ddt_frame = velocities_update(self.history, self.velocities, frame_idx, final, self.frame_rate,
                                           self.calibration, cfg, self.ddt_frame)
        return {"images": imgs, "final": final, "min_gradient": pa, "rightmost_sobel": pb, "search": (s0, s1)}


# --------------------------------------------------------------------------------------
# layer 2: operation-by-operation restatement on the centre band
# --------------------------------------------------------------------------------------
def reflect(i: int, n: int) -> int:
    """scipy.ndimage mode='reflect' (d c b a | a b c d | d c b a), any distance."""
    period = 2 * n
    i %= period
    return i if i < n else period - 1 - i


def gaussian_weights(sigma: float = 1.5, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) (symmetric, so the [::-1]
    applied by gaussian_filter1d is a no-op)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


def _rows(a: np.ndarray, rows: List[int]) -> np.ndarray:
    h = a.shape[0]
    return np.stack([a[reflect(r, h)] for r in rows])


def _minmax3_cols(a: np.ndarray, op) -> np.ndarray:
    """3-tap min/max along axis 1 with reflect boundaries."""
    left = np.concatenate([a[:, :1], a[:, :-1]], axis=1)     # reflect: index -1 -> 0
    right = np.concatenate([a[:, 1:], a[:, -1:]], axis=1)    # index n -> n-1
    return op(op(left, a), right)


def _corr_sym(get, weights: np.ndarray) -> np.ndarray:
    """NI_Correlate1D symmetric branch: centre first, then pairs outermost -> innermost."""
    r = (len(weights) - 1) // 2
    tmp = get(0) * weights[r]
    for j in range(-r, 0):
        tmp = tmp + (get(j) + get(-j)) * weights[r + j]
    return tmp


def detect_lines_restated(frame_diff: np.ndarray, sigma: float = 1.5):
    """Centre-row Sobel and gradient lines from the 19-row band (kernel size 3 fixed)."""
    h, w = frame_diff.shape
    c = h // 2
    wts = gaussian_weights(sigma)
    r = (len(wts) - 1) // 2                      # 6

    # rows needed, outermost stage first (all with reflect in the row direction)
    # sobel row c needs blurred rows c-1..c+1 -> gaussian(axis 0) rows c-1-r..c+1+r of the opening
    nr_rows = list(range(c - 1 - r, c + 1 + r + 1))                 # 15 rows of noise_removed
    er_rows = list(range(nr_rows[0] - 1, nr_rows[-1] + 2))          # 17 rows of the erosion
    # erosion rows need diff rows +-1 more (19 rows); each stage reflects on the TRUE image height
    def eroded_row(rr: int) -> np.ndarray:
        rr = reflect(rr, h)
        band = _rows(frame_diff, [rr - 1, rr, rr + 1])
        return _minmax3_cols(band, np.minimum).min(axis=0)
    er = {rr: eroded_row(rr) for rr in er_rows}

    def opened_row(rr: int) -> np.ndarray:
        rr0 = reflect(rr, h)
        band = np.stack([er_get(rr0 - 1), er_get(rr0), er_get(rr0 + 1)])
        return _minmax3_cols(band, np.maximum).max(axis=0)

    def er_get(rr: int) -> np.ndarray:
        key = reflect(rr, h)
        if key not in er:
            er[key] = eroded_row(key)
        return er[key]

    nr = {}

    def nr_get(rr: int) -> np.ndarray:
        key = reflect(rr, h)
        if key not in nr:
            nr[key] = opened_row(key)
        return nr[key]

    # gaussian along axis 0 (rows) for rows c-1, c, c+1, then along axis 1 (columns)
    def blurred_row(rr: int) -> np.ndarray:
        rr0 = reflect(rr, h)
        g0 = _corr_sym(lambda j: nr_get(rr0 + j), wts)                       # axis 0
        def col(j):                                                          # g0 shifted by j, reflect
            idx = np.array([reflect(x + j, w) for x in range(w)])
            return g0[idx]
        return _corr_sym(col, wts)                                           # axis 1

    b_m, b_c, b_p = blurred_row(c - 1), blurred_row(c), blurred_row(c + 1)

    # sobel(axis=1): correlate1d([-1,0,1], axis 1) (antisymmetric branch) then [1,2,1] along axis 0
    def d1(line: np.ndarray) -> np.ndarray:
        left = np.concatenate([line[:1], line[:-1]])
        right = np.concatenate([line[1:], line[-1:]])
        tmp = line * 0.0                                  # centre tap: in[0] * w[centre] with w = 0
        return tmp + (left - right) * -1.0                # tmp += (in[-1] - in[+1]) * w[-1], w[-1] = -1
    s_m, s_c, s_p = d1(b_m), d1(b_c), d1(b_p)
    sob = s_c * 2.0                                       # symmetric [1,2,1]: centre first
    sob = sob + (s_m + s_p) * 1.0

    # np.gradient(axis=1), unit spacing, edge_order 1
    grad = np.empty(w, dtype=np.float64)
    grad[1:-1] = (b_c[2:] - b_c[:-2]) / 2.0
    grad[0] = (b_c[1] - b_c[0]) / 1.0
    grad[-1] = (b_c[-1] - b_c[-2]) / 1.0
    return sob, grad


# --------------------------------------------------------------------------------------
# the sequential tracker
# --------------------------------------------------------------------------------------
@dataclass
class HeadConfig:
    """FlameDetectorConfig defaults (scripts/process_videos.py:164-193)."""
    frame_diff_threshold: float = 5.0
    gaussian_sigma: float = 1.5
    min_gradient_strength: float = 10.0
    edge_margin_px: int = 10
    sobel_threshold_fraction: float = 0.1
    max_velocity_change_m_s: float = 200.0
    ddt_velocity_jump_m_s: float = 1250.0
    search_window_px: int = 100
    exit_margin_px: int = 15
    min_signal_fraction: float = 0.0005


def max_displacement_px(frame_rate: float, calibration: float, cfg: HeadConfig) -> int:
    """:270-276."""
    if frame_rate <= 0 or calibration <= 0:
        return 1000
    dt = 1.0 / frame_rate
    return int(np.ceil(cfg.max_velocity_change_m_s * dt / calibration)) + 1


def select_position(sobel_line, gradient_line, search_start: int, search_end: int, cfg: HeadConfig):
    """:420-465 -> (pos_min_gradient, pos_rightmost_sobel, final)."""
    pos_a = pos_b = None
    ss = sobel_line[search_start:search_end]
    sg = gradient_line[search_start:search_end]
    if len(ss) > 0 and len(sg) > 0:
        if np.min(sg) < -cfg.min_gradient_strength:
            pos_a = search_start + int(np.argmin(sg))
        smax = np.max(np.abs(ss))
        if smax > cfg.min_gradient_strength:
            above = np.abs(ss) > smax * cfg.sobel_threshold_fraction
            if np.any(above):
                pos_b = search_start + int(np.max(np.where(above)[0]))
    cands = [p for p in (pos_a, pos_b) if p is not None]
    return pos_a, pos_b, (max(cands) if cands else None)


@dataclass
class HeadResult:
    background: float
    per_frame: List[dict] = field(default_factory=list)       # one entry per detect() call
    rows: List[list] = field(default_factory=list)            # [frame, time_s, px, pos_m, post_ddt]
    velocity_history: List[list] = field(default_factory=list)
    ddt_frame: Optional[int] = None
    stop: Optional[Tuple[str, int]] = None
    empty: int = 0


def velocities_update(history, vel_history, frame_idx, final, frame_rate, calibration, cfg, ddt_frame):
    """:479-516.  Mutates vel_history; returns the (possibly new) ddt frame."""
    if final is not None and len(history) >= 2:
        curr_frame, curr_pos = history[-1]
        prev_frame, prev_pos = history[-2]
        if prev_pos is not None and frame_rate > 0:
            dt = (curr_frame - prev_frame) / frame_rate
            if dt > 0:
                v1 = (curr_pos - prev_pos) * calibration / dt
                v2 = vc = None
                if len(history) >= 3:
                    _, prev2_pos = history[-3]
                    if prev2_pos is not None:
                        v2 = (3 * curr_pos - 4 * prev_pos + prev2_pos) * calibration / (2 * dt)
                        vc = (curr_pos - prev2_pos) * calibration / (2 * dt)
                        if len(vel_history) >= 1:
                            o = vel_history[-1]
                            vel_history[-1] = [o[0], o[1], o[2], vc]
                vel_history.append([frame_idx, v1, v2, None])
                if ddt_frame is None and len(vel_history) >= 2:
                    if v1 - vel_history[-2][1] > cfg.ddt_velocity_jump_m_s:
                        ddt_frame = frame_idx
    return ddt_frame


def run_head(frames: np.ndarray, frame_rate: float, calibration: float, offset: float, time_of,
             cfg: Optional[HeadConfig] = None, lines=detect_lines_scipy) -> HeadResult:
    """The reference loop :1441-1516 with the HEAD detector, plotting removed."""
    cfg = cfg or HeadConfig()
    n, h, w = frames.shape
    bg = fo.background_scalar(frames[0])
    noise_thr = fo.empty_noise_threshold(bg)
    maxdisp = max_displacement_px(frame_rate, calibration, cfg)
    res = HeadResult(background=bg)
    history: List[Tuple[int, Optional[int]]] = []
    prior = None
    for i in range(n):
        sub = fo.subtract_scalar_background(frames[i], bg)
        if fo.is_empty_frame(sub, noise_thr, cfg.min_signal_fraction):
            res.empty += 1
            prior = sub
            continue
        # search bounds (:317-348)
        last = next(((f, p) for f, p in reversed(history) if p is not None), None)
        if last is None:
            s0, s1 = cfg.edge_margin_px, w - cfg.edge_margin_px
        else:
            lf, lp = last
            s0 = lp
            s1 = min(w - cfg.edge_margin_px, lp + maxdisp * max(1, i - lf) + cfg.search_window_px)
        pa = pb = final = None
        if prior is not None:
            d = fo.frame_difference(sub, prior, cfg.frame_diff_threshold)
            sl, gl = lines(d)
            pa, pb, final = select_position(sl, gl, s0, s1, cfg)
        history.append((i, final))
        prior = sub
        res.ddt_frame = velocities_update(history, res.velocity_history, i, final, frame_rate, calibration, cfg,
                                          res.ddt_frame)
        res.per_frame.append({"frame": i, "final": final, "min_gradient": pa, "rightmost_sobel": pb,
                              "search": [s0, s1]})
        velocity = res.velocity_history[-1][1] if res.velocity_history else None
        if final is not None and final >= w - cfg.exit_margin_px:          # :1488-1494
            if len(res.velocity_history) >= 2:
                e = res.velocity_history[-2]
                res.velocity_history[-2] = [e[0], e[1], e[2], None]
            res.stop = ("exit", i)
            break
        if velocity is not None and len(res.velocity_history) >= 2:        # :1499-1509
            pv1 = res.velocity_history[-2][1]
            if pv1 is not None and pv1 > 100 and (pv1 - velocity) / pv1 > 0.5:
                e = res.velocity_history[-2]
                res.velocity_history[-2] = [e[0], e[1], e[2], None]
                res.stop = ("velocity_drop", i)
                break
        if final is not None:
            post = res.ddt_frame is not None and i >= res.ddt_frame
            res.rows.append([i, time_of(i), int(final), final * calibration + offset, bool(post)])
    return res
