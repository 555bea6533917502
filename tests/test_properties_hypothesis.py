"""Property tests (hypothesis) of the host logic and of the oracle itself - CPU only.

* the 12-bit packing the whole path is built on is a bijection;
* the integer bounds handed to the kernels are equivalent to the reference's float comparisons
  for EVERY threshold, not just the defaults;
* the range partition is the reference's contiguous `distribute_indices`;
* the oracle is range-split invariant (the property the multi-GPU decomposition relies on);
* the range-block merge reproduces the serial truncation for arbitrary exits and world sizes.
"""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED
from high_speed_image_processing_b200.engine import (ClipScalars, DetectionParams, derive_kernel_bounds,
                                                     min_signal_count)
from high_speed_image_processing_b200.process_videos import FileCalibration
from high_speed_image_processing_b200.sharding import assign_videos, contiguous_range
from oracle import flame_oracle as fo

FAST = settings(max_examples=60, deadline=None)


@FAST
@given(st.lists(st.integers(0, 4095), min_size=2, max_size=400).filter(lambda v: len(v) % 2 == 0))
def test_pack12_unpack12_roundtrip(values):
    px = np.array(values, dtype=np.uint16)
    packed = fo.pack12(px)
    assert packed.dtype == np.uint8 and packed.size == px.size * 3 // 2
    assert np.array_equal(fo.unpack12(packed), px)
    # layout: b0 b1 b2 -> (b0<<4)|(b1>>4), ((b1&15)<<8)|b2
    b = packed.astype(np.int64).reshape(-1, 3)
    assert np.array_equal((b[:, 0] << 4) | (b[:, 1] >> 4), px[0::2])
    assert np.array_equal(((b[:, 1] & 15) << 8) | b[:, 2], px[1::2])


@FAST
@given(st.integers(1, 1 << 21), st.floats(0.0, 1.0, allow_nan=False))
def test_min_signal_count_is_the_exact_float_boundary(n_px, frac):
    c = min_signal_count(n_px, frac)
    assert 0 <= c <= n_px + 1
    if c <= n_px:
        assert not (c / n_px < frac)
    if c > 0:
        assert (c - 1) / n_px < frac


@FAST
@given(st.floats(-50.0, 5000.0, allow_nan=False), st.floats(0.0, 500.0, allow_nan=False),
       st.floats(0.0, 2000.0, allow_nan=False), st.floats(0.0, 9000.0, allow_nan=False))
def test_integer_bounds_match_float_comparisons(noise_thr, diff_thr, min_grad, flame_thr):
    sc = ClipScalars(background=60.0, centerline_mean=40.0, centerline_std=4.0, centerline_max=55.0,
                     flame_threshold=flame_thr, noise_threshold=noise_thr)
    kb = derive_kernel_bounds(sc, DetectionParams(frame_diff_threshold=diff_thr, min_gradient_strength=min_grad), 1000)
    v = np.arange(-4200, 8400, dtype=np.int64)
    assert np.array_equal(v > noise_thr, v > kb.empty_thr)
    assert np.array_equal(v > flame_thr, v > kb.threshold_floor)
    assert np.array_equal(v < diff_thr, v < kb.diff_thr)
    assert np.array_equal(v / 2.0 < -min_grad, v < kb.grad2_bound)       # np.gradient halves, compared doubled


@FAST
@given(st.integers(0, 5000), st.integers(1, 64))
def test_contiguous_range_partitions_the_frames(total, size):
    spans = [contiguous_range(total, r, size) for r in range(size)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
        assert b0 == a1 and b0 >= a0
    lengths = [b - a for a, b in spans]
    assert max(lengths) - min(lengths) <= 1 and lengths == sorted(lengths, reverse=True)   # remainder to the first ranks


@FAST
@given(st.lists(st.integers(1, 10 ** 6), min_size=1, max_size=40), st.integers(1, 8))
def test_assign_videos_covers_every_video_once(weights, size):
    parts = [assign_videos(len(weights), r, size, weights) for r in range(size)]
    assert sorted(sum(parts, [])) == list(range(len(weights)))
    loads = [sum(weights[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(weights)                       # greedy longest-first bound
    assert [assign_videos(len(weights), r, size) for r in range(size)] == [list(range(r, len(weights), size))
                                                                           for r in range(size)]


@FAST
@given(st.integers(0, 99999), st.integers(0, 999), st.integers(0, 999))
def test_file_calibration_range_uses_the_last_integer(last, lo, hi):
    rule = FileCalibration(calibration=1.0, files=[f"run-{lo}-:run-{hi}-"])
    name = f"run-7-_C001H001S{last:04d}.cihx"
    assert rule.matches(name) == (lo <= last <= hi)                      # bug-compatible with the reference (:94-99)


@settings(max_examples=12, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from(["threshold", "gradient", "half_maximum"]),
       st.integers(1, 5), st.data())
def test_oracle_is_range_split_invariant(seed, method, n_cuts, data):
    """Serial run == concatenation of sub-ranges with a one-frame halo (+ min of the exits): the
    decomposition rule of SURVEY 8e, checked on the oracle so that the GPU tests may rely on it."""
    from high_speed_image_processing_b200 import synthetic as syn
    spec = syn.SyntheticSpec(width=64, height=6, n_frames=40, style="mini" if method != "half_maximum" else "nova",
                             t_enter=4.0, velocity=2.0, curvature_px=1.0, seed=seed % 10007)
    frames = syn.render_frames(spec)
    n = len(frames)
    serial = fo.process_clip(frames, fo.ClipParams(method=method))
    cuts = sorted(set(data.draw(st.lists(st.integers(1, n - 1), min_size=n_cuts, max_size=n_cuts))))
    edges = [0] + cuts + [n]
    pos, exits = [], []
    for a, b in zip(edges, edges[1:]):
        part = fo.process_clip(frames[a:b], fo.ClipParams(method=method), frame0=frames[0], first_index=a,
                               prior_frame=frames[a - 1] if a else None)
        pos.append(part.pos_px)
        exits.append(a + part.first_exit if part.first_exit < b - a else n)
    joined = np.concatenate(pos)
    fe = min(exits)
    assert fe == serial.first_exit
    assert np.array_equal(joined[:fe], serial.pos_px[:fe])


@FAST
@given(st.integers(1, 9), st.integers(0, 300), st.data())
def test_range_block_merge_equals_serial_truncation(world, total, data):
    """The arithmetic of merge_ranges_kernel (csrc/ff_exchange.cu), restated in NumPy: owner /
    offset of every frame, min of the block headers, truncation."""
    cap = max(1, -(-total // world)) + data.draw(st.integers(0, 3))
    pos = np.array(data.draw(st.lists(st.integers(-1, 500), min_size=total, max_size=total)), dtype=np.int64)
    exits = []
    for r in range(world):
        a, b = contiguous_range(total, r, world)
        exits.append(data.draw(st.one_of(st.just(FF_NO_EXIT), st.integers(a, max(a, b - 1)))) if b > a else FF_NO_EXIT)
    fe = min(exits)
    base, extra = divmod(total, world)
    out = np.empty(total, dtype=np.int64)
    for i in range(total):                            # owner_of() in the kernel
        boundary = extra * (base + 1)
        if i < boundary:
            r, off = divmod(i, base + 1)
        else:
            r, off = extra + (i - boundary) // base, (i - boundary) % base
        a, b = contiguous_range(total, r, world)
        assert a + off == i and off < cap and a <= i < b
        out[i] = FF_POS_DROPPED if i >= fe else pos[a + off]
    want = pos.copy()
    want[min(fe, total):] = FF_POS_DROPPED
    assert np.array_equal(out, want)


def test_float_time_and_position_expressions():
    assert fo.frame_time_absolute(39, 500, 1, 160000) == (500 + 39 * 1) / 160000
    assert f"{fo.frame_time_absolute(39, 500, 1, 160000):.9f}" == "0.003368750"          # README.md:95
    assert f"{fo.position_m(6, 0.000833333, 1.347567):.9f}" == "1.352566998"
    assert math.isclose(fo.position_m(14, 0.000833333, 1.347567), 1.359233662, rel_tol=1e-9)


# ---- velocity / DDT bookkeeping of the frame-level detector -----------------------------------------
@settings(max_examples=120, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 4), st.one_of(st.none(), st.integers(0, 1023))), min_size=1, max_size=40),
       st.sampled_from([0.0, 1000.0, 160000.0]), st.sampled_from([0.000833333, 0.002, 0.0]),
       st.sampled_from([50.0, 1250.0]))
def test_velocity_book_equals_the_oracle_update(steps, frame_rate, calibration, ddt_jump):
    """head.VelocityBook (used by FlameDetector.detect and by the whole-clip HEAD path) against the
    oracle's restatement of scripts/process_videos.py:474-516 on arbitrary detection sequences with
    gaps, missing positions, zero frame rate / calibration."""
    from high_speed_image_processing_b200.head import VelocityBook
    from oracle import head_oracle as ho
    cfg = ho.HeadConfig(ddt_velocity_jump_m_s=ddt_jump)
    book = VelocityBook(frame_rate, calibration, ddt_jump)
    history, vel, ddt = [], [], None
    frame = 0
    for gap, pos in steps:
        frame += gap
        book.update(frame, pos)
        history.append((frame, pos))
        ddt = ho.velocities_update(history, vel, frame, pos, frame_rate, calibration, cfg, ddt)
        assert book.velocities == vel and book.ddt_frame == ddt and book.history == history
    last = next(((f, p) for f, p in reversed(history) if p is not None), (-1, -1))
    assert book.last_detection() == last
