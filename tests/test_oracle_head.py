"""HEAD-parity detector, CPU side: the operation-by-operation restatement equals SciPy, the
oracle loop equals the reference's replay, and the product's host bookkeeping + writers
reproduce the files the reference's own driver wrote (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

from high_speed_image_processing_b200.head import (HeadParams, finish_head_track, gaussian_weights,
                                                    max_displacement_px)
from high_speed_image_processing_b200.process_videos import VideoResult, write_head_outputs
from oracle import flame_oracle as fo
from oracle import head_oracle as ho


@pytest.fixture(scope="module")
def frames(clip_small, golden):
    c = golden["clip_small"]
    return fo.frames_from_bytes(clip_small["packed"], c["n_frames"], c["height"], c["width"], c["bits"])


@pytest.mark.parametrize("h,w", [(16, 128), (64, 300), (7, 40), (3, 33), (19, 64), (1, 16), (2, 20), (40, 5)])
def test_restated_lines_equal_scipy(h, w):
    rng = np.random.default_rng(h * 100 + w)
    for trial in range(3):
        d = rng.integers(0, 3000, size=(h, w)).astype(np.float64)
        d[d < (1500 if trial else 5)] = 0
        s1, g1 = ho.detect_lines_scipy(d)
        s2, g2 = ho.detect_lines_restated(d)
        assert np.array_equal(s1, s2) and np.array_equal(g1, g2)


def test_gaussian_weights_are_scipys():
    from scipy.ndimage import _filters
    for sigma in (1.0, 1.5, 2.0):
        radius = int(4.0 * sigma + 0.5)
        assert np.array_equal(gaussian_weights(sigma), _filters._gaussian_kernel1d(sigma, 0, radius))
        assert np.array_equal(ho.gaussian_weights(sigma), gaussian_weights(sigma))


@pytest.mark.parametrize("lines", [ho.detect_lines_scipy, ho.detect_lines_restated])
def test_oracle_loop_matches_reference_replay(frames, golden, lines):
    hr = golden["head_replay"]
    res = ho.run_head(frames, 160000, 0.000833333, 1.347567,
                      lambda i: fo.frame_time_absolute(i, 500, 1, 160000), lines=lines)
    assert res.per_frame == [{k: v for k, v in p.items() if k != "diff_sha1"} for p in hr["per_frame"]]
    assert res.rows == hr["results"]
    assert res.velocity_history == hr["velocity_history"]
    assert res.ddt_frame == hr["ddt_frame"] and list(res.stop) == hr["stop"] and res.empty == hr["empty"]


def _track_from_golden(hr, n):
    track = np.full((n, 5), -1, dtype=np.int32)
    flags = np.zeros(n, dtype=np.uint8)
    for p in hr["per_frame"]:
        v = lambda x: -1 if x is None else x
        track[p["frame"]] = [v(p["final"]), v(p["min_gradient"]), v(p["rightmost_sobel"]), *p["search"]]
        flags[p["frame"]] = 1
    return track, flags


def test_host_bookkeeping_matches_reference(golden):
    hr = golden["head_replay"]
    n = golden["clip_small"]["n_frames"]
    track, flags = _track_from_golden(hr, n)
    flags[hr["stop"][1] + 1:] = 1                      # frames after the stop must be ignored
    s = finish_head_track(track, flags, 0, golden["clip_small"]["width"], 160000, 0.000833333, 1.347567,
                          lambda i: fo.frame_time_absolute(i, 500, 1, 160000), HeadParams())
    assert [list(r) for r in s.rows] == hr["results"]
    assert s.velocity_history == hr["velocity_history"]
    assert s.ddt_frame == hr["ddt_frame"] and list(s.stop) == hr["stop"]
    assert max_displacement_px(160000, 0.000833333, HeadParams()) == ho.max_displacement_px(160000, 0.000833333,
                                                                                             ho.HeadConfig())
    assert max_displacement_px(0, 1.0, HeadParams()) == 1000


def test_velocity_files_equal_the_reference_drivers(golden, tmp_path):
    hr = golden["head_replay"]
    n = golden["clip_small"]["n_frames"]
    track, flags = _track_from_golden(hr, n)
    s = finish_head_track(track, flags, 0, golden["clip_small"]["width"], 160000, 0.000833333, 1.347567,
                          lambda i: fo.frame_time_absolute(i, 500, 1, 160000), HeadParams())
    res = VideoResult(None, np.zeros(n, np.int32), np.zeros(n, np.int32), None, list(s.rows), 0,
                      s.velocity_history, s.ddt_frame, s.stop)
    written = write_head_outputs(res, tmp_path, "run-3-")
    for path in written:
        name = path.rsplit("/", 1)[1]
        assert open(path).read() == golden["driver_outputs"][name], name
    assert sorted(p.rsplit("/", 1)[1] for p in written) == sorted(k for k in golden["driver_outputs"]
                                                                 if k.startswith("run-3--"))


def test_ddt_and_velocity_drop_rules():
    """Synthetic position tracks exercising :511-516 and :1499-1509."""
    n, w = 40, 4000
    hp = HeadParams()
    flags = np.ones(n, dtype=np.uint8)
    track = np.full((n, 5), -1, dtype=np.int32)
    pos = 100
    for i in range(n):
        pos += 2 if i < 20 else 40                      # velocity jump at frame 20
        track[i] = [pos, pos, pos, 0, w]
    s = finish_head_track(track, flags, 0, w, 10000.0, 0.01, 0.0, lambda i: i / 10000.0, hp)
    ref_like = []
    assert s.ddt_frame == 20                            # (40-2)*0.01*1e4 = 3800 m/s jump > 1250
    assert all(r[4] == (r[0] >= 20) for r in s.rows)
    # velocity drop > 50 % stops the run without recording the frame
    track2 = track.copy()
    track2[30:, 0] = track2[29, 0] + 1
    s2 = finish_head_track(track2, flags, 0, w, 10000.0, 0.01, 0.0, lambda i: i / 10000.0, hp)
    assert s2.stop == ("velocity_drop", 30) and s2.rows[-1][0] == 29
    assert s2.velocity_history[-2][3] is None           # central difference cleared (:654-663)
    assert HeadParams(morphology_kernel_size=5).morphology_kernel_size == 5       # odd sizes up to 7 run on the GPU
    for bad in (0, 2, 4, 9):
        with pytest.raises(ValueError):
            HeadParams(morphology_kernel_size=bad)
