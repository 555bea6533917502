"""The oracle against the vectors produced by running the reference (oracle/make_golden.py)
and against the README's known-answer rows.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import flame_oracle as fo


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def frames(clip_small, golden):
    c = golden["clip_small"]
    return fo.frames_from_bytes(clip_small["packed"], c["n_frames"], c["height"], c["width"], c["bits"])


def test_unpack12_known_bytes():
    # b0 b1 b2 = 0xAB 0xCD 0xEF -> p0 = 0xABC, p1 = 0xDEF
    assert fo.unpack12(np.array([0xAB, 0xCD, 0xEF], dtype=np.uint8)).tolist() == [0xABC, 0xDEF]
    assert fo.unpack12(np.array([0xFF, 0xF0, 0x00, 0x00, 0x0F, 0xFF], dtype=np.uint8)).tolist() == [4095, 0, 0, 4095]


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    px = rng.integers(0, 4096, size=4098, dtype=np.uint16)
    assert np.array_equal(fo.unpack12(fo.pack12(px)), px)
    with pytest.raises(ValueError):
        fo.unpack12(np.zeros(4, dtype=np.uint8))


def test_frames_from_bytes_8_16():
    rng = np.random.default_rng(1)
    a8 = rng.integers(0, 256, size=(3, 4, 6), dtype=np.uint8)
    assert np.array_equal(fo.frames_from_bytes(a8.reshape(-1), 3, 4, 6, 8), a8)
    a16 = rng.integers(0, 65536, size=(3, 4, 6), dtype=np.uint16)
    assert np.array_equal(fo.frames_from_bytes(a16.astype("<u2").view(np.uint8).reshape(-1), 3, 4, 6, 16), a16)


def test_primitives_match_reference(frames, golden):
    g = golden["primitives"]
    assert fo.background_scalar(frames[0]) == g["background"]
    mean, std, mx, thr = fo.centerline_stats(frames[0])
    assert (mean, std, mx, thr) == (g["centerline_mean"], g["centerline_std"], g["centerline_max"],
                                    g["flame_threshold"])
    assert fo.empty_noise_threshold(g["background"]) == g["noise_threshold"]
    prior = None
    for i in range(len(frames)):
        sub = fo.subtract_scalar_background(frames[i], g["background"])
        assert _sha(sub) == g["sub_sha1"][i]
        assert fo.is_empty_frame(sub, g["noise_threshold"], 0.0005) == g["empty"][i]
        assert fo.nonempty_count(sub, g["noise_threshold"]) == g["nonempty_count"][i]
        if prior is not None:
            assert _sha(fo.frame_difference(sub, prior, 5.0)) == g["diff_sha1"][i]
        prior = sub


def test_clip_loop_matches_reference_primitives(frames, golden, clip_small_profiles):
    g = golden["primitives"]
    res = fo.process_clip(frames, fo.ClipParams(method="gradient", keep_profiles=True))
    assert list(res.empty) == g["empty"]
    assert list(res.nonempty) == g["nonempty_count"]
    assert np.array_equal(res.profiles[1:], clip_small_profiles[1:])
    # gradient == HEAD Method A primitives evaluated by the reference run
    want = golden["gradient_on_profiles"]
    for i in range(1, len(frames)):
        assert fo.detect_gradient(clip_small_profiles[i], 10.0) == want[i]
        if not res.empty[i]:
            assert (int(res.pos_px[i]) if res.pos_px[i] >= 0 else None) == want[i]


def test_head_replay_diff_images(frames, golden):
    """FlameDetector.detect's own frame_diff (scripts/process_videos.py:397-399) pins a9/a10,
    including the prior-frame carry across skipped-empty frames (:1462)."""
    g = golden["head_replay"]
    bg = g["background"]
    subs = [fo.subtract_scalar_background(f, bg) for f in frames]
    for rec in g["per_frame"]:
        i = rec["frame"]
        if rec["diff_sha1"] is not None:
            assert _sha(fo.frame_difference(subs[i], subs[i - 1], 5.0)) == rec["diff_sha1"]


def test_time_and_position_formulas(golden):
    c = golden["clip_small"]
    for i, t in enumerate(golden["video"]["absolute_time"]):
        assert fo.frame_time_absolute(i, c["start_frame"], 1, c["record_rate"]) == t
    for i, t in enumerate(golden["video"]["time_trigger10"]):
        assert fo.frame_time_relative(i, 10, 160000) == t
    assert fo.frame_time_absolute(3, 0, 1, 0) == 0.0
    for row in golden["head_replay"]["results"]:
        f, t, px, pm, _ = row
        assert fo.position_m(px, 0.000833333, 1.347567) == pm
        assert fo.frame_time_absolute(f, c["start_frame"], 1, c["record_rate"]) == t


def test_readme_known_answer_rows(golden):
    """README.md:93-96 - the only known-answer data in the reference repository."""
    for row in golden["readme_rows"]:
        t = fo.frame_time_absolute(row["frame"], 500, 1, 160000)
        assert f"{t:.9f}" == row["time_s"]
        assert f"{fo.position_m(row['px'], 0.000833333, 1.347567):.9f}" == row["pos_m"]


# ---- frozen spec of the two prose-only methods and edge cases -----------------
def test_threshold_spec():
    p = np.array([0, 9, 9, 0, 9, 0, 0], dtype=float)
    assert fo.detect_threshold(p, 5.0) == 4
    assert fo.detect_threshold(p, 5.0, min_run_px=2) == 2
    assert fo.detect_threshold(p, 5.0, min_run_px=3) is None
    assert fo.detect_threshold(p, 9.0) is None            # strict >
    assert fo.detect_threshold(np.full(5, 7.0), 1.0) == 4  # run reaching the right edge
    assert fo.detect_threshold(np.full(5, 7.0), 1.0, min_run_px=5) == 4


def test_half_maximum_spec():
    assert fo.detect_half_maximum(np.array([0, 10, 10, 6, 5, 4, 0], dtype=float)) == 5   # first argmax, p < 5
    assert fo.detect_half_maximum(np.array([0, 0, 0], dtype=float)) is None              # peak <= 0
    assert fo.detect_half_maximum(np.array([1, 2, 3], dtype=float)) is None              # peak at right edge
    assert fo.detect_half_maximum(np.array([8, 4, 4, 3], dtype=float)) == 3              # strict <


def test_gradient_spec():
    p = np.array([0, 0, 50, 50, 0, 0], dtype=float)
    assert fo.detect_gradient(p, 10.0) == 3                # first of the two equal minima (-25)
    assert fo.detect_gradient(np.array([0, 0, 10, 0], dtype=float), 10.0) is None   # -10 is not < -10
    assert fo.detect_gradient(np.array([30, 0], dtype=float), 10.0) == 0            # one-sided ends


def test_exit_truncation_and_skip_frames(frames):
    base = fo.process_clip(frames, fo.ClipParams(method="threshold", exit_margin_px=10))
    assert base.first_exit < len(frames)
    assert all(f < base.first_exit for f, _ in base.records)
    assert base.pos_px[base.first_exit] >= frames.shape[2] - 10
    # a skipped frame is not processed and does not become anybody's prior frame
    k = base.records[5][0]
    sk = fo.process_clip(frames, fo.ClipParams(method="threshold", skip_frames=[k]))
    assert sk.pos_px[k] == -1
    bg = base.background
    d = fo.frame_difference(fo.subtract_scalar_background(frames[k + 1], bg),
                            fo.subtract_scalar_background(frames[k - 1], bg), 5.0)
    want = fo.detect_threshold(d[frames.shape[1] // 2], base.flame_threshold)
    assert sk.pos_px[k + 1] == (want if want is not None else -1)


def test_subrange_equals_serial(frames):
    """Evaluating a contiguous range with its one-frame halo reproduces the serial loop."""
    full = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    a, b = 20, 47
    part = fo.process_clip(frames[a:b], fo.ClipParams(method="half_maximum"), frame0=frames[0], first_index=a,
                           prior_frame=frames[a - 1])
    assert np.array_equal(part.pos_px, full.pos_px[a:b])
    assert np.array_equal(part.nonempty, full.nonempty[a:b])
