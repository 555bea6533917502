"""Exact integer bounds the host derives for the kernels (engine.derive_kernel_bounds)."""
import numpy as np
import pytest

from high_speed_image_processing_b200.engine import (ClipScalars, DetectionParams, derive_kernel_bounds,
                                                     frame_nbytes, min_signal_count)
from oracle import flame_oracle as fo


@pytest.mark.parametrize("n_px", [1, 7, 1999, 2000, 2001, 32768, 131072, 262144, 1048576, 999983])
@pytest.mark.parametrize("frac", [0.0005, 0.001, 0.0, 1.0, 0.3333])
def test_min_signal_count_reproduces_is_empty(n_px, frac):
    c = min_signal_count(n_px, frac)
    for k in {0, max(0, c - 2), max(0, c - 1), c, min(n_px, c + 1), n_px}:
        frame = np.zeros(n_px)
        frame[:k] = 100.0
        assert fo.is_empty_frame(frame, 50.0, frac) == (k < c), (n_px, frac, k, c)


def test_bounds_are_exact_for_integer_data():
    rng = np.random.default_rng(0)
    sc = ClipScalars(background=57.0, centerline_mean=40.1, centerline_std=3.9, centerline_max=53.0,
                     flame_threshold=106.5, noise_threshold=28.5)
    kb = derive_kernel_bounds(sc, DetectionParams(frame_diff_threshold=4.5, min_gradient_strength=10.25), 1000)
    v = rng.integers(-200, 4096, size=5000)
    assert np.array_equal(v > sc.noise_threshold, v > kb.empty_thr)
    assert np.array_equal(v > sc.flame_threshold, v > kb.threshold_floor)
    assert np.array_equal(v < 4.5, v < kb.diff_thr)
    assert np.array_equal(v / 2.0 < -10.25, v < kb.grad2_bound)
    kb2 = derive_kernel_bounds(sc, DetectionParams(), 1000)          # integral thresholds: strictness kept
    assert np.array_equal(v < 5.0, v < kb2.diff_thr) and np.array_equal(v / 2.0 < -10.0, v < kb2.grad2_bound)


def test_clip_scalars_follow_the_reference_expressions(golden):
    g = golden["primitives"]
    z = np.load("tests/golden/clip_small.npz")
    c = golden["clip_small"]
    frames = fo.frames_from_bytes(z["packed"], c["n_frames"], c["height"], c["width"], 12)
    sc = ClipScalars.from_frame0_stats(int(frames[0].max()), frames[0][c["height"] // 2])
    assert (sc.background, sc.centerline_mean, sc.centerline_std, sc.centerline_max, sc.flame_threshold,
            sc.noise_threshold) == (g["background"], g["centerline_mean"], g["centerline_std"], g["centerline_max"],
                                    g["flame_threshold"], g["noise_threshold"])


def test_param_validation():
    with pytest.raises(ValueError):
        DetectionParams(method="sobel")
    with pytest.raises(ValueError):
        DetectionParams(min_run_px=0)
    with pytest.raises(ValueError):
        frame_nbytes(3, 3, 12)
    with pytest.raises(ValueError):
        frame_nbytes(4, 4, 10)
    assert frame_nbytes(128, 1024, 12) == 196608
