"""Diagnostic figures of the reference driver, off the timed path.

``save_frame_image`` (scripts/process_videos.py:783-1107), ``generate_stacked_sequence`` (:1110-1186)
and ``generate_stacked_sequence_single_column`` (:1189-1270) under their reference names and
signatures.  Everything that is computed comes from the GPU: the per-frame images are the fields of
a ``FlameDetectionResult`` (``detector.FlameDetector.detect`` -> ``ff_head_images``), the stacked
sequences use ``ff_frame_subtract_background`` / ``ff_frame_difference``
(``stacked_sequence_arrays``).  Only the drawing is Matplotlib, imported when a figure is asked for;
without it these functions raise ``ImportError`` (Matplotlib is an optional dependency, as the
figures are an optional output: ``process_video_source(..., diagnostics=True)``).
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .detector import FlameDetectionResult, FlameDetector, FlameDetectorConfig, is_empty_frame, \
    subtract_prior_frame, subtract_scalar_background


def _pyplot():
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
    except ImportError as exc:
        raise ImportError("diagnostic figures need matplotlib (optional dependency); the flame-front "
                          "results themselves do not") from exc
    return plt


# --------------------------------------------------------------------------------------
# data for the figures (GPU)
# --------------------------------------------------------------------------------------
def stacked_sequence_arrays(video, frame_indices: Sequence[int], background_scalar: float,
                            engine=None) -> Tuple[np.ndarray, np.ndarray]:
    """What the stacked-sequence figures show (:1146-1157, :1222-1232): per listed frame the
    background-subtracted image and the difference of the RAW frame against the previously listed
    raw frame (threshold 0; zeros for the first).  Returns two float64 arrays ``[n,H,W]``."""
    height, width = video.frame_shape
    sub = np.zeros((len(frame_indices), height, width), dtype=np.float64)
    diff = np.zeros_like(sub)
    prior = None
    for i, frame_idx in enumerate(frame_indices):
        frame = video[frame_idx]
        sub[i] = subtract_scalar_background(frame, background_scalar, engine=engine)
        if prior is not None:
            diff[i] = subtract_prior_frame(frame, prior, threshold=0.0, engine=engine)
        prior = frame
    return sub, diff


def display_limit(image: Optional[np.ndarray], signed: bool = False) -> float:
    """Colour-scale limit of an image panel: its 99th percentile (of |x| for signed images), or 1
    for an image without signal (:877,:890,:903,:916,:929)."""
    if image is None:
        return 1.0
    if signed:
        return float(np.percentile(np.abs(image), 99)) if np.any(image != 0) else 1.0
    return float(np.percentile(image, 99)) if np.any(image > 0) else 1.0


# --------------------------------------------------------------------------------------
# per-frame figure (:783-1107)
# --------------------------------------------------------------------------------------
_IMAGE_PANELS = (          # (field, title, colour map, signed)
    ("frame_diff", "2. Frame Diff (current - prior)", "hot", False),
    ("noise_removed", "3. Noise Removed (morphological opening)", "hot", False),
    ("blurred", "4. Gaussian Blur", "hot", False),
    ("sobel_output", "5. Sobel Filter (horizontal)", "RdBu", True),
    ("gradient_output", "6. Gradient Filter (np.gradient)", "RdBu", True),
)


def _mark_positions(ax, result: FlameDetectionResult, final: bool = True, width: float = 1.5, labels: bool = False,
                    which=("search", "gradient", "sobel")) -> None:
    lab = (lambda text: text) if labels else (lambda text: None)
    if result.search_bounds and "search" in which:
        lo, hi = result.search_bounds
        ax.axvline(x=lo, color="lime", linestyle="--", linewidth=width, alpha=0.8, label=lab(f"Search: {lo}-{hi}"))
        ax.axvline(x=hi, color="lime", linestyle=":", linewidth=width, alpha=0.8)
    if result.pos_min_gradient is not None and "gradient" in which:
        ax.axvline(x=result.pos_min_gradient, color="purple", linewidth=2, alpha=0.7,
                   label=lab(f"Min Grad: {result.pos_min_gradient}"))
    if result.pos_rightmost_sobel is not None and "sobel" in which:
        ax.axvline(x=result.pos_rightmost_sobel, color="orange", linewidth=2, alpha=0.7,
                   label=lab(f"R-Sobel: {result.pos_rightmost_sobel}"))
    if final and result.final_position is not None:
        ax.axvline(x=result.final_position, color="red", linewidth=3, alpha=0.9,
                   label=lab(f"FINAL: {result.final_position}"))


def _host(a):
    """Image fields may be CUDA tensors (``FlameDetector(intermediates="device")``)."""
    return a if a is None or isinstance(a, np.ndarray) else a.detach().cpu().numpy()


def save_frame_image(frame: np.ndarray, result: FlameDetectionResult, output_path: Path, source_name: str,
                     detector: Optional[FlameDetector] = None) -> None:
    """All processing steps of one frame stacked vertically: six images, three centre-row profiles,
    the result overlay, position history with the spline estimator, velocity comparison - written to
    ``<output_path>/<source_name>-Frame-<frame_idx:06d>.png`` (:1104)."""
    plt = _pyplot()
    height, width = frame.shape[:2]
    center_row = height // 2
    x_pixels = np.arange(width)
    image_h, plot_h = 1.5, 2.5
    ratios = [image_h] * 6 + [plot_h] * 3 + [image_h, plot_h, plot_h]
    fig = plt.figure(figsize=(14, sum(ratios)))
    grid = fig.add_gridspec(12, 1, height_ratios=ratios, hspace=0.3)
    axes = [fig.add_subplot(grid[i, 0]) for i in range(12)]
    images = {name: _host(getattr(result, name)) for name in
              ("frame_subtracted", "frame_diff", "noise_removed", "blurred", "sobel_output", "gradient_output")}
    velocity = f" | v={detector.last_velocity:.1f} m/s" if detector is not None and detector.last_velocity is not None else ""

    # 1: background-subtracted frame
    ax = axes[0]
    ax.imshow(images["frame_subtracted"], cmap="gray", aspect="auto")
    ax.axhline(y=center_row, color="cyan", linestyle="--", linewidth=0.5, alpha=0.5)
    _mark_positions(ax, result)
    ax.set_title(f"1. BG Subtracted - Frame {result.frame_idx} | t={result.time_s * 1e6:.1f} µs{velocity}", fontsize=10)
    ax.set_ylabel("Y")
    # 2-6: the detector's intermediate images
    for ax, (name, title, cmap, signed) in zip(axes[1:6], _IMAGE_PANELS):
        img = images[name]
        if img is None:
            ax.text(0.5, 0.5, "No prior frame" if name == "frame_diff" else "N/A", ha="center", va="center",
                    transform=ax.transAxes, fontsize=12)
            ax.set_facecolor("lightgray")
        else:
            vmax = display_limit(img, signed)
            ax.imshow(img, cmap=cmap, aspect="auto", vmin=-vmax if signed else 0, vmax=vmax)
            ax.axhline(y=center_row, color="black" if signed else "cyan", linestyle="--", linewidth=0.5, alpha=0.5)
            _mark_positions(ax, result)
        ax.set_title(title, fontsize=10)
        ax.set_ylabel("Y")
    # 7-9: centre-row profiles
    profiles = (("frame_diff", "7. Frame Diff Centerline", "Intensity", "r-", ("search", "gradient", "sobel")),
                ("sobel_output", "8. Sobel Centerline", "Sobel Value", "b-", ("search", "sobel")),
                ("gradient_output", "9. Gradient Centerline (min = leading edge)", "Gradient Value", "purple",
                 ("search", "gradient")))
    for ax, (name, title, ylabel, style, which) in zip(axes[6:9], profiles):
        img = images[name]
        if img is not None:
            line = img[center_row, :]
            ax.plot(x_pixels, line, style, linewidth=1.5 if name == "frame_diff" else 1)
            if name == "frame_diff":
                ax.fill_between(x_pixels, 0, line, alpha=0.3, color="red")
            else:
                ax.axhline(y=0, color="gray", linewidth=0.5)
        _mark_positions(ax, result, width=2, labels=True, which=which)
        ax.set_xlim(0, width)
        ax.set_ylabel(ylabel)
        ax.set_title(title, fontsize=10)
        ax.legend(loc="upper right", fontsize=8, ncol=3 if name == "frame_diff" else 1)
        ax.grid(True, alpha=0.3)
    # 10: result overlay
    ax = axes[9]
    ax.imshow(images["frame_subtracted"], cmap="gray", aspect="auto")
    ax.axhline(y=center_row, color="cyan", linestyle="--", linewidth=0.5, alpha=0.5)
    _mark_positions(ax, result, final=False, width=2, which=("search",))
    for pos, marker, colour, size, label in ((result.pos_min_gradient, "p", "purple", 6, "Min Grad"),
                                             (result.pos_rightmost_sobel, "s", "orange", 6, "R-Sobel"),
                                             (result.pos_spline_predicted, "^", "cyan", 6, "Spline")):
        if pos is not None:
            ax.plot(pos, center_row, marker, color=colour, markersize=size, label=f"{label}: {pos}")
    if result.final_position is not None:
        ax.plot(result.final_position, center_row, "o", color="red", markersize=8, markeredgecolor="yellow",
                markeredgewidth=1, label=f"FINAL: {result.final_position}")
    ax.legend(loc="upper right", fontsize=8, ncol=2)
    outcome = f"FINAL: x={result.final_position} px" if result.final_position else "No detection"
    ax.set_title(f"10. Result: {outcome}{velocity}", fontsize=10)
    ax.set_ylabel("Y")
    # 11: position history + spline estimator
    ax = axes[10]
    detected = [(f, p) for f, p in detector.position_history if p is not None] if detector is not None else []
    if detector is not None and len(detector.position_history) > 0:
        if detected:
            ax.scatter([f for f, _ in detected], [p for _, p in detected], c="blue", s=20, alpha=0.7,
                       label="Detected positions", zorder=3)
            curve = detector.get_spline_curve()
            if curve is not None:
                ax.plot(curve[0], curve[1], "g-", linewidth=2, label="Spline estimator", zorder=2)
            ax.axvline(x=result.frame_idx, color="red", linestyle="--", linewidth=1.5, alpha=0.7)
            if result.final_position is not None:
                ax.scatter([result.frame_idx], [result.final_position], c="red", s=60, marker="*", zorder=5,
                           label=f"Current: {result.final_position}")
            if result.pos_spline_predicted is not None:
                ax.scatter([result.frame_idx], [result.pos_spline_predicted], c="cyan", s=40, marker="^", zorder=4,
                           label=f"Spline pred: {result.pos_spline_predicted}")
            ax.legend(loc="upper left", fontsize=8)
    else:
        ax.text(0.5, 0.5, "No history yet", ha="center", va="center", transform=ax.transAxes, fontsize=12)
    ax.set_ylabel("Position (pixels)")
    ax.set_title("11. Position History + Spline Estimator", fontsize=10)
    ax.grid(True, alpha=0.3)
    # 12: the three velocity estimates
    ax = axes[11]
    history = detector.get_velocity_history() if detector is not None else []
    if history:
        series = (("b-", 1, "1st-order backward", 1.5), ("g--", 2, "2nd-order backward", 1.5),
                  ("r:", 3, "2nd-order central", 2))
        for style, column, label, lw in series:
            pts = [(e[0], e[column]) for e in history if e[column] is not None]
            if pts:
                ax.plot([f for f, _ in pts], [v for _, v in pts], style, linewidth=lw, alpha=0.8, label=label)
        ax.axhline(y=0, color="gray", linewidth=0.5)
        if detector.ddt_detected:
            ax.axvline(x=detector.ddt_frame, color="magenta", linestyle="--", linewidth=2,
                       label=f"DDT @ frame {detector.ddt_frame}")
        v1 = detector.last_velocities[0]
        ax.scatter([result.frame_idx], [v1] if v1 else [], c="blue", s=40, marker="*", zorder=5)
        ax.legend(loc="upper left", fontsize=7)
    else:
        ax.text(0.5, 0.5, "No velocity data yet", ha="center", va="center", transform=ax.transAxes, fontsize=12)
    ax.set_xlabel("Frame Index")
    ax.set_ylabel("Velocity (m/s)")
    ddt = f" | DDT @ {detector.ddt_frame}" if detector is not None and detector.ddt_detected else ""
    ax.set_title(f"12. Velocity Comparison{ddt}", fontsize=10)
    ax.grid(True, alpha=0.3)

    plt.savefig(Path(output_path) / f"{source_name}-Frame-{result.frame_idx:06d}.png", dpi=120, bbox_inches="tight")
    plt.close(fig)


# --------------------------------------------------------------------------------------
# stacked sequences (:1110-1270)
# --------------------------------------------------------------------------------------
def generate_stacked_sequence(video, frame_indices: List[int], background_scalar: float, output_path: Path,
                              title: str = "", show_frame_diff: bool = True, figsize_width: float = 10.0,
                              engine=None) -> None:
    """Frames stacked vertically, one row per listed frame: background-subtracted (and, in a second
    column, the raw frame difference) - the paper-style figure of :1110-1186."""
    plt = _pyplot()
    sub, diff = stacked_sequence_arrays(video, frame_indices, background_scalar, engine)
    height, width = video.frame_shape
    n_cols = 2 if show_frame_diff else 1
    panel_height = (figsize_width / n_cols) / (width / height)
    fig, axes = plt.subplots(len(frame_indices), n_cols, figsize=(figsize_width, panel_height * len(frame_indices)),
                             squeeze=False)
    for i in range(len(frame_indices)):
        for col, stack in enumerate((sub, diff)[:n_cols]):
            ax = axes[i, col]
            ax.imshow(stack[i], cmap="gray", aspect="equal", vmin=0)
            ax.set_xticks([])
            ax.set_yticks([])
        axes[i, 0].set_ylabel(f"{i + 1}", rotation=0, labelpad=20, fontsize=10, fontweight="bold", color="white")
    plt.subplots_adjust(wspace=0.02, hspace=0)
    if title:
        fig.suptitle(title, fontsize=12, fontweight="bold", color="white")
    plt.savefig(output_path, dpi=300, bbox_inches="tight", facecolor="black", edgecolor="none")
    plt.close(fig)
    print(f"Saved stacked sequence: {output_path}")


def generate_stacked_sequence_single_column(video, frame_indices: List[int], background_scalar: float,
                                            output_path: Path, use_frame_diff: bool = False, title: str = "",
                                            figsize_width: float = 6.0, engine=None) -> None:
    """The compact variant (:1189-1270): all listed frames in ONE image, numbered at the left."""
    plt = _pyplot()
    sub, diff = stacked_sequence_arrays(video, frame_indices, background_scalar, engine)
    height, width = video.frame_shape
    stacked = (diff if use_frame_diff else sub).reshape(len(frame_indices) * height, width)
    fig, ax = plt.subplots(figsize=(figsize_width, figsize_width / (width / stacked.shape[0])))
    ax.imshow(stacked, cmap="gray", aspect="equal", vmin=0)
    for i in range(len(frame_indices)):
        ax.text(-width * 0.02, i * height + height // 2, f"{i + 1}", color="white", fontsize=8, fontweight="bold",
                ha="right", va="center")
        if i > 0:
            ax.axhline(y=i * height - 0.5, color="white", linewidth=0.5, alpha=0.5)
    ax.set_xlim(-width * 0.05, width)
    ax.set_xticks([])
    ax.set_yticks([])
    ax.set_facecolor("black")
    if title:
        ax.set_title(title, color="white", fontsize=10, fontweight="bold")
    plt.savefig(output_path, dpi=300, bbox_inches="tight", facecolor="black", edgecolor="none")
    plt.close(fig)
    print(f"Saved stacked sequence: {output_path}")


# --------------------------------------------------------------------------------------
# what the reference driver renders for one recording (:1385-1418, :1441-1494)
# --------------------------------------------------------------------------------------
def render_video_diagnostics(video, source_name: str, stem: str, calibration: float, frames_output_dir,
                             skip_frames: Sequence[int] = (), detector_config: Optional[FlameDetectorConfig] = None,
                             engine=None) -> int:
    """The two stacked-sequence figures plus one 12-panel figure per frame that reaches the detector,
    with the reference's file names.  Replays the clip frame by frame through the GPU
    ``FlameDetector`` (the whole-clip kernels do not keep full-frame intermediates).  Returns the
    number of per-frame figures written."""
    out_dir = Path(frames_output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    total = len(video)
    background_scalar = float(np.max(video[0]))                                   # :1357-1358
    n_display = min(15, total)                                                    # :1388-1390
    display = list(range(0, total, max(1, total // n_display)))[:n_display]
    generate_stacked_sequence(video, display, background_scalar, out_dir / f"{stem}-stacked-sequence.png",
                              title=stem, show_frame_diff=True, figsize_width=12.0, engine=engine)
    generate_stacked_sequence_single_column(video, display, background_scalar, out_dir / f"{stem}-stacked-single.png",
                                            use_frame_diff=False, title=stem, figsize_width=8.0, engine=engine)
    cfg = detector_config or FlameDetectorConfig(gaussian_sigma=1.5, morphology_kernel_size=3,
                                                 max_velocity_change_m_s=200.0)      # :1372-1377
    det = FlameDetector(cfg, video.frame_rate, calibration, engine=engine, keep_results=False)
    written = 0
    for frame_idx in range(total):
        if frame_idx in skip_frames:
            continue
        frame = video[frame_idx]
        sub = subtract_scalar_background(frame, background_scalar, engine=engine)
        if is_empty_frame(sub, noise_threshold=max(10.0, background_scalar * 0.5), min_signal_fraction=0.0005,
                          engine=engine):
            det._prior_frame = sub
            continue
        result = det.detect(frame=frame, frame_idx=frame_idx, background_scalar=background_scalar)
        save_frame_image(frame, result, out_dir, source_name, det)
        written += 1
        velocity, history = det.last_velocity, det.get_velocity_history()
        if result.final_position is not None and result.final_position >= video.width - cfg.exit_margin_px:
            break                                                                 # :1488-1494
        if velocity is not None and len(history) >= 2 and history[-2][1] > 100 and \
                (history[-2][1] - velocity) / history[-2][1] > 0.5:
            break                                                                 # :1499-1509
    return written
