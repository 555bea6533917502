"""Frame-level API on the GPU (seam B3): FlameDetector.detect, the element-wise frame functions and
ff_head_images, bit for bit against the vectors recorded from the reference (tests/golden) and
against the oracle's SciPy statement on shapes the golden clip does not cover."""
import hashlib

import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import process_videos as pv
from high_speed_image_processing_b200 import synthetic as syn
from oracle import flame_oracle as fo
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu

IMAGES = ("frame_subtracted", "frame_diff", "noise_removed", "blurred", "sobel_output", "gradient_output")


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def api(golden):
    return golden["detector_api"]


@pytest.fixture(scope="module")
def frames(api):
    from conftest import GOLDEN
    f = np.load(GOLDEN / "detector_frames.npz")["frames"]
    assert _sha(f) == api["frames_sha1"]
    return f


def test_detector_reproduces_reference_calls(engine, api, frames):
    """Every field of every FlameDetectionResult the reference returned, for three configurations
    (defaults with the spline estimator; 5x5 opening / sigma 1 / frame gaps; 1x1 opening / sigma 2)."""
    for run in api["runs"]:
        d = pv.FlameDetector(pv.FlameDetectorConfig(**run["cfg"]), api["frame_rate"], run["calibration"], engine=engine)
        assert d._max_displacement_px == run["max_displacement_px"]
        for call in run["calls"]:
            idx = call["frame"]
            bounds = d.get_search_bounds(idx, frames.shape[2])
            r = d.detect(frame=frames[idx], frame_idx=idx, background_scalar=run["background"])
            where = (run["name"], idx)
            assert (r.frame_idx, r.time_s) == (idx, call["time_s"]), where
            assert (r.final_position, r.pos_min_gradient, r.pos_rightmost_sobel) == \
                (call["final"], call["min_gradient"], call["rightmost_sobel"]), where
            assert list(r.search_bounds) == call["search"] == list(bounds), where
            assert r.pos_spline_predicted == call["spline"], where
            for name in IMAGES:
                img = getattr(r, name)
                assert (img is None) == (call["sha1"][name] is None), where + (name,)
                if img is not None:
                    assert img.dtype == np.float64 and img.shape == frames.shape[1:]
                    assert _sha(img) == call["sha1"][name], where + (name,)
        assert [list(e) for e in d.get_velocity_history()] == run["velocity_history"]
        assert d.ddt_frame == run["ddt_frame"] and d.ddt_detected == (run["ddt_frame"] is not None)
        assert d.last_position == run["last_position"] and list(d.last_velocities) == run["last_velocities"]
        assert d.last_velocity == run["last_velocities"][0]
        assert len(d.get_pre_ddt_velocities()) == run["pre_ddt"] and len(d.get_post_ddt_velocities()) == run["post_ddt"]
        assert _sha(d._prior_frame) == run["prior_sha1"]
        assert len(d._detection_results) == len(run["calls"]) and len(d.position_history) == len(run["calls"])
        if run["cfg"].get("use_spline_estimator", True):
            curve = d.get_spline_curve()
            assert curve is not None and curve[0].shape == curve[1].shape == (100,)
        d.reset()
        assert d.position_history == [] and d._prior_frame is None and d.last_velocity is None
        assert d.get_search_bounds(3, 136) == (d.config.edge_margin_px, 136 - d.config.edge_margin_px)


@pytest.mark.parametrize("h,w,dtype,k,sigma", [(16, 128, np.uint16, 3, 1.5), (5, 40, np.uint16, 3, 1.5),
                                                (1, 70, np.uint16, 3, 1.5), (33, 130, np.uint8, 3, 1.5),
                                                (48, 257, np.uint16, 5, 0.8), (17, 66, np.uint16, 7, 2.0),
                                                (128, 1024, np.uint16, 3, 1.5), (3, 2, np.uint16, 3, 1.5)])
def test_detector_images_bit_exact_vs_scipy(engine, h, w, dtype, k, sigma):
    """Tile edges, frames smaller than the stencil halo (multiple reflections), 8-bit frames, other
    opening sizes and Gaussian radii: every intermediate equals the reference's SciPy sequence."""
    bits = 8 if dtype == np.uint8 else 16
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=9, bits=bits, style="mini", t_enter=1.0,
                             velocity=max(1.0, w / 10.0), curvature_px=1.0, seed=31 * h + w)
    frames = syn.render_frames(spec)
    rng = np.random.default_rng(h * w)
    frames[4] = rng.integers(0, 256 if bits == 8 else 65536, size=(h, w)).astype(dtype)   # full-range noise frame
    bg = float(np.max(frames[0]))
    cfg = pv.FlameDetectorConfig(morphology_kernel_size=k, gaussian_sigma=sigma, edge_margin_px=min(10, w // 4))
    d = pv.FlameDetector(cfg, 50000.0, 0.001, engine=engine)
    hcfg = ho.HeadConfig(gaussian_sigma=sigma, edge_margin_px=cfg.edge_margin_px)
    o = ho.FrameDetectorOracle(50000.0, 0.001, hcfg, kernel_size=k)
    for i in range(len(frames)):
        r = d.detect(frames[i], i, bg)
        want = o.detect(frames[i], i, bg)
        for name in IMAGES:
            a, b = getattr(r, name), want["images"][name]
            assert (a is None) == (b is None), (i, name)
            if a is not None:
                assert np.array_equal(a, b), (i, name, float(np.abs(a - b).max()))
        assert (r.final_position, r.pos_min_gradient, r.pos_rightmost_sobel, tuple(r.search_bounds)) == \
            (want["final"], want["min_gradient"], want["rightmost_sobel"], tuple(want["search"])), i
    assert [list(e) for e in d.get_velocity_history()] == o.velocities


def test_detector_follows_the_reference_driver_loop(engine, golden, clip_small):
    """The reference's own loop (:1441-1516) written against this package's names - including the
    assignment to ``_prior_frame`` for frames skipped as empty (:1462) - gives the recorded replay."""
    c, want = golden["clip_small"], golden["head_replay"]
    video = fo.frames_from_bytes(clip_small["packed"], c["n_frames"], c["height"], c["width"], c["bits"])
    cal, off = 0.000833333, 1.347567
    background_scalar = float(np.max(video[0]))
    det = pv.FlameDetector(pv.FlameDetectorConfig(gaussian_sigma=1.5, morphology_kernel_size=3,
                                                  max_velocity_change_m_s=200.0),
                           frame_rate=c["record_rate"], calibration_m_per_px=cal, engine=engine,
                           intermediates="none")
    per_frame, results, empty, stop = [], [], 0, None
    for frame_idx in range(len(video)):
        frame = video[frame_idx]
        sub = pv.subtract_scalar_background(frame, background_scalar, engine=engine)
        if pv.is_empty_frame(sub, noise_threshold=max(10.0, background_scalar * 0.5), min_signal_fraction=0.0005,
                             engine=engine):
            empty += 1
            det._prior_frame = sub.copy()
            continue
        r = det.detect(frame=frame, frame_idx=frame_idx, background_scalar=background_scalar)
        assert r.frame_diff is None and r.frame_subtracted is None          # intermediates="none"
        per_frame.append({"frame": frame_idx, "final": r.final_position, "min_gradient": r.pos_min_gradient,
                          "rightmost_sobel": r.pos_rightmost_sobel, "search": list(r.search_bounds)})
        pos, velocity = r.final_position, det.last_velocity
        if pos is not None and pos >= c["width"] - 15:
            det.clear_last_central_difference()
            stop = ["exit", frame_idx]
            break
        vh = det.get_velocity_history()
        if velocity is not None and len(vh) >= 2 and vh[-2][1] > 100 and (vh[-2][1] - velocity) / vh[-2][1] > 0.5:
            det.clear_last_central_difference()
            stop = ["velocity_drop", frame_idx]
            break
        if pos is not None:
            results.append([frame_idx, golden["video"]["absolute_time"][frame_idx], int(pos), pos * cal + off,
                            bool(det.ddt_detected and frame_idx >= det.ddt_frame)])
    assert empty == want["empty"] and stop == want["stop"]
    assert per_frame == [{k: v for k, v in e.items() if k != "diff_sha1"} for e in want["per_frame"]]
    assert results == want["results"]
    assert [list(e) for e in det.get_velocity_history()] == want["velocity_history"]
    assert det.ddt_frame == want["ddt_frame"]


def test_detector_options_and_errors(engine, frames):
    bg = float(np.max(frames[0]))
    dev_det = pv.FlameDetector(pv.FlameDetectorConfig(), 1e5, 0.001, engine=engine, intermediates="device",
                               keep_results=False)
    host_det = pv.FlameDetector(pv.FlameDetectorConfig(), 1e5, 0.001, engine=engine)
    for i in (3, 4):
        a = dev_det.detect(torch.from_numpy(frames[i]).to(engine.device), i, bg)      # CUDA tensor in
        b = host_det.detect(frames[i], i, bg)
        assert a.final_position == b.final_position
        assert isinstance(a.frame_subtracted, torch.Tensor) and a.frame_subtracted.is_cuda
        assert np.array_equal(a.frame_subtracted.cpu().numpy(), b.frame_subtracted)
    assert np.array_equal(a.gradient_output.cpu().numpy(), b.gradient_output)
    assert dev_det._detection_results == [] and len(host_det._detection_results) == 2

    d = pv.FlameDetector(pv.FlameDetectorConfig(), 1e5, 0.001, engine=engine)
    with pytest.raises(TypeError):
        d.detect(frames[1].astype(np.float64), 1, bg)
    with pytest.raises(ValueError):
        d.detect(frames[:2], 1, bg)                               # not one 2-D frame
    with pytest.raises(ValueError):
        d.detect(frames[1], 1, 40.5)                              # non-integer background
    with pytest.raises(ValueError):
        d.detect(frames[1], 1, -1.0)
    d.detect(frames[1], 1, bg)
    with pytest.raises(ValueError):
        d.detect(frames[2][:, :64], 2, bg)                        # shape differs from the prior frame
    with pytest.raises(ValueError):
        d._prior_frame = np.full((40, 136), 0.5)                  # not a background-subtracted camera frame
    with pytest.raises(ValueError):
        pv.FlameDetector(pv.FlameDetectorConfig(morphology_kernel_size=4), 1e5, 0.001, engine=engine).detect(frames[1], 1, bg)
    with pytest.raises(ValueError):
        pv.FlameDetector(pv.FlameDetectorConfig(), 1e5, 0.001, engine=engine, intermediates="disk")
    # frame_rate / calibration unknown: no displacement constraint, time 0 (:270-276, :377)
    z = pv.FlameDetector(pv.FlameDetectorConfig(), 0, 0.001, engine=engine)
    assert z._max_displacement_px == 1000 and z.detect(frames[1], 7, bg).time_s == 0
    # a background that changes between calls: the prior keeps the value it was subtracted with (:469)
    c1 = pv.FlameDetector(pv.FlameDetectorConfig(), 1e5, 0.001, engine=engine)
    c1.detect(frames[5], 5, 30.0)
    got = c1.detect(frames[6], 6, 45.0)
    want = fo.frame_difference(fo.subtract_scalar_background(frames[6], 45.0),
                               fo.subtract_scalar_background(frames[5], 30.0), 5.0)
    assert np.array_equal(got.frame_diff, want)


def test_frame_functions_match_reference_vectors(engine, api, frames):
    ops = api["ops"]
    f5, f6, f7 = frames[5], frames[6], frames[7]
    sub6 = pv.subtract_scalar_background(f6, 41.5, engine=engine)
    assert isinstance(sub6, np.ndarray) and sub6.dtype == np.float64 and _sha(sub6) == ops["sub_bg_41.5"]
    assert _sha(pv.subtract_scalar_background(f6, float(np.max(frames[0])), engine=engine)) == ops["sub_bg_max0"]
    assert _sha(pv.subtract_prior_frame(f6, f5, threshold=0.0, engine=engine)) == ops["prior_raw_thr0"]
    assert _sha(pv.subtract_prior_frame(f6, f5, threshold=7.5, engine=engine)) == ops["prior_raw_thr7.5"]
    sub5 = pv.subtract_scalar_background(f5, 41.5, engine=engine)
    assert _sha(pv.subtract_prior_frame(sub6, sub5, threshold=5.0, engine=engine)) == ops["prior_f64"]
    assert _sha(pv.three_frame_difference(f5, f6, f7, engine=engine)) == ops["three_thr0"]
    assert _sha(pv.three_frame_difference(f5, f6, f7, threshold=3.0, engine=engine)) == ops["three_thr3"]
    for key, want in ops["empty"].items():
        thr, frac = (float(v) for v in key.split("/"))
        assert pv.is_empty_frame(sub6, noise_threshold=thr, min_signal_fraction=frac, engine=engine) == want, key
    assert pv.is_empty_frame(f6, engine=engine) == ops["empty_raw_defaults"]
    dev6 = torch.from_numpy(sub6).to(engine.device)
    for thr, want in ops["count_above"].items():
        assert engine.frame_count_above(dev6, float(thr)) == want


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int32, np.float32, np.float64])
def test_frame_functions_vs_oracle_any_dtype(engine, dtype):
    rng = np.random.default_rng(7)
    shape = (37, 53)
    if np.dtype(dtype).kind == "f":
        a, b, c = (rng.normal(100.0, 60.0, size=shape).astype(dtype) for _ in range(3))
        b[3, 4] = np.nan
        a[0, 0], c[0, 1] = np.inf, -np.inf
    else:
        hi = 255 if dtype == np.uint8 else 4000
        a, b, c = (rng.integers(0, hi + 1, size=shape).astype(dtype) for _ in range(3))
    with np.errstate(invalid="ignore"):
        for bgv in (0.0, 33.25, 1e6):
            assert np.array_equal(pv.subtract_scalar_background(a, bgv, engine=engine),
                                  fo.subtract_scalar_background(a, bgv), equal_nan=True)
        for thr in (0.0, 12.5, -3.0):
            assert np.array_equal(pv.subtract_prior_frame(a, b, thr, engine=engine), fo.frame_difference(a, b, thr),
                                  equal_nan=True)
            assert np.array_equal(pv.three_frame_difference(a, b, c, thr, engine=engine),
                                  fo.three_frame_difference(a, b, c, thr), equal_nan=True)
        for thr, frac in ((50.0, 0.001), (100.1, 0.5), (1e9, 1e-9), (-1.0, 1.0)):
            assert pv.is_empty_frame(a, thr, frac, engine=engine) == fo.is_empty_frame(a, thr, frac)
    # mixed dtypes widen to float64 like astype(np.float64) does
    assert np.array_equal(pv.subtract_prior_frame(a, b.astype(np.float64), 1.0, engine=engine),
                          fo.frame_difference(a, b.astype(np.float64), 1.0), equal_nan=True)
    # CUDA tensor in -> CUDA tensor out
    if dtype in (np.uint8, np.uint16, np.float64):
        t = pv.subtract_scalar_background(torch.from_numpy(a).to(engine.device), 20.0, engine=engine)
        assert isinstance(t, torch.Tensor) and t.is_cuda
        assert np.array_equal(t.cpu().numpy(), fo.subtract_scalar_background(a, 20.0), equal_nan=True)
    with pytest.raises(ValueError):
        pv.subtract_prior_frame(a, b[:, :10], engine=engine)


@pytest.mark.parametrize("bits", [12, 16, 8])
def test_head_images_on_a_packed_range(engine, bits):
    """ff_head_images over frames stored as in the .mraw file: halo frame, skip_frames entries
    (prior = latest non-skipped frame, :1443-1445), a sub-selection of the outputs."""
    spec = syn.SyntheticSpec(width=160, height=24, n_frames=14, bits=bits, style="nova", t_enter=1.0, velocity=9.0,
                             tail_length=25.0, seed=900 + bits)
    frames = syn.render_frames(spec)
    bg = int(np.max(frames[0]))
    packed = torch.from_numpy(syn.pack_frames(frames, bits)).to(engine.device)
    fb = frames[0].size * bits // 8
    skip = np.zeros(13, dtype=np.uint8)
    skip[[4, 5, 9]] = 1
    out = engine.head_images(packed[fb:], 13, 24, 160, bits, bg, halo=packed[:fb],
                             skip=torch.from_numpy(skip).to(engine.device))
    state = out["state"].cpu().numpy()
    prior = fo.subtract_scalar_background(frames[0], bg)
    for i in range(13):
        if skip[i]:
            assert state[i] == 0
            for name in IMAGES:
                assert not out[name][i].any()
            continue
        sub = fo.subtract_scalar_background(frames[i + 1], bg)
        want = ho.detect_images_scipy(sub, prior)
        assert state[i] == 1
        for name in IMAGES:
            assert np.array_equal(out[name][i].cpu().numpy(), want[name]), (i, name)
        prior = sub
    # without a halo the first frame has no prior: state 2, only frame_subtracted is meaningful
    few = engine.head_images(packed, 3, 24, 160, bits, bg, want=("frame_subtracted", "sobel_output"))
    assert sorted(few) == ["frame_subtracted", "sobel_output", "stack", "state"]
    assert few["state"].cpu().tolist() == [2, 1, 1]
    assert np.array_equal(few["frame_subtracted"][0].cpu().numpy(), fo.subtract_scalar_background(frames[0], bg))
    assert not few["sobel_output"][0].any()
    want2 = ho.detect_images_scipy(fo.subtract_scalar_background(frames[2], bg), fo.subtract_scalar_background(frames[1], bg))
    assert np.array_equal(few["sobel_output"][2].cpu().numpy(), want2["sobel_output"])
    with pytest.raises(ValueError):
        engine.head_images(packed, 3, 24, 160, bits, bg, want=("nope",))
    with pytest.raises(ValueError):
        engine.head_images(packed, 3, 24, 160, bits, bg, morphology_kernel_size=9)
    with pytest.raises(ValueError):
        engine.head_images(packed, 3, 24, 160, bits, bg, frame_diff_threshold=-1.0)
