"""Diagnostic figures.  Matplotlib is not installed here, so a recording stand-in
checks what the figure functions hand to it: which arrays, which markers, which file names."""
import numpy as np
import pytest

from high_speed_image_processing_b200 import diagnostics as dg
from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200.detector import FlameDetectionResult
from oracle import flame_oracle as fo
from oracle import head_oracle as ho


def calls(log, name):
    return [(a, k) for n, a, k in log if n == name or n.endswith("." + name)]


class FakeDetector:
    last_velocity = 412.5
    last_velocities = (412.5, 400.0, None)
    ddt_detected = True
    ddt_frame = 7
    position_history = [(5, None), (6, 40), (7, 52)]

    def get_spline_curve(self):
        return np.linspace(6, 7, 100), np.linspace(40, 52, 100)

    def get_velocity_history(self):
        return [(6, 300.0, None, 350.0), (7, 412.5, 400.0, None)]


def _result(frames, i, with_prior=True):
    bg = float(np.max(frames[0]))
    sub = fo.subtract_scalar_background(frames[i], bg)
    prior = fo.subtract_scalar_background(frames[i - 1], bg) if with_prior else None
    im = ho.detect_images_scipy(sub, prior)
    return FlameDetectionResult(frame_idx=i, time_s=i / 1e5, pos_min_gradient=50 if with_prior else None,
                                pos_rightmost_sobel=52 if with_prior else None, pos_spline_predicted=51,
                                search_bounds=(40, 120), final_position=52 if with_prior else None, **im)


def test_without_matplotlib_the_figures_raise_import_error(tmp_path):
    frames = syn.render_frames(syn.SyntheticSpec(width=128, height=16, n_frames=8, t_enter=1.0, velocity=8.0, seed=3))
    with pytest.raises(ImportError, match="matplotlib"):
        dg.save_frame_image(frames[7], _result(frames, 7), tmp_path, "Nova")


def test_save_frame_image_panels_and_file_name(tmp_path, fake_pyplot):
    frames = syn.render_frames(syn.SyntheticSpec(width=128, height=16, n_frames=8, t_enter=1.0, velocity=8.0, seed=3))
    res = _result(frames, 7)
    dg.save_frame_image(frames[7], res, tmp_path, "Nova", FakeDetector())
    log = fake_pyplot
    shown = calls(log, "imshow")
    assert len(shown) == 7                                        # six images + the result overlay
    for (args, kw), want in zip(shown, [res.frame_subtracted, res.frame_diff, res.noise_removed, res.blurred,
                                        res.sobel_output, res.gradient_output, res.frame_subtracted]):
        assert args[0] is want
    assert shown[1][1]["vmax"] == float(np.percentile(res.frame_diff, 99)) and shown[1][1]["vmin"] == 0
    lim = float(np.percentile(np.abs(res.sobel_output), 99))
    assert shown[4][1]["vmin"] == -lim and shown[4][1]["vmax"] == lim and shown[4][1]["cmap"] == "RdBu"
    assert calls(log, "savefig")[0][0][0] == tmp_path / "Nova-Frame-000007.png"
    assert len(calls(log, "add_subplot")) == 12 and len(calls(log, "close")) == 1
    titles = [a[0] for a, _ in calls(log, "set_title")]
    assert titles[0].startswith("1. BG Subtracted - Frame 7 | t=70.0 µs | v=412.5 m/s")
    assert titles[9] == "10. Result: FINAL: x=52 px | v=412.5 m/s" and titles[11] == "12. Velocity Comparison | DDT @ 7"
    profile = [a for a, _ in calls(log, "plot") if len(a) >= 2 and isinstance(a[1], np.ndarray) and a[1].shape == (128,)]
    assert any(np.array_equal(a[1], res.frame_diff[8]) for a in profile)
    assert any(np.array_equal(a[1], res.sobel_output[8]) for a in profile)
    assert any(np.array_equal(a[1], res.gradient_output[8]) for a in profile)
    finals = [k for a, k in calls(log, "axvline") if k.get("x") == 52 and k.get("color") == "red"]
    assert len(finals) >= 9                                       # six images + three profiles


def test_save_frame_image_without_prior_or_detector(tmp_path, fake_pyplot):
    frames = syn.render_frames(syn.SyntheticSpec(width=128, height=16, n_frames=4, t_enter=1.0, velocity=8.0, seed=3))
    res = _result(frames, 2, with_prior=False)
    dg.save_frame_image(frames[2], res, tmp_path, "Mini", None)
    log = fake_pyplot
    assert len(calls(log, "imshow")) == 2                         # only the subtracted frame, twice
    texts = [a[2] for a, _ in calls(log, "text")]
    assert texts == ["No prior frame", "N/A", "N/A", "N/A", "N/A", "No history yet", "No velocity data yet"]
    assert [a[0] for a, _ in calls(log, "set_title")][9] == "10. Result: No detection"
    assert calls(log, "savefig")[0][0][0] == tmp_path / "Mini-Frame-000002.png"


def test_display_limit():
    assert dg.display_limit(None) == 1.0 and dg.display_limit(np.zeros((3, 3))) == 1.0
    a = np.arange(100.0).reshape(10, 10)
    assert dg.display_limit(a) == float(np.percentile(a, 99))
    assert dg.display_limit(-a, signed=True) == float(np.percentile(a, 99))
    assert dg.display_limit(-a) == 1.0


@pytest.mark.gpu
def test_stacked_sequences_from_the_gpu(engine, clip_small_on_disk, clip_small, golden, tmp_path, fake_pyplot):
    from high_speed_image_processing_b200.photron import open_video
    c = golden["clip_small"]
    frames = fo.frames_from_bytes(clip_small["packed"], c["n_frames"], c["height"], c["width"], c["bits"])
    idx = [0, 9, 18, 40, 41]
    bg = float(np.max(frames[0]))
    with open_video(str(clip_small_on_disk)) as video:
        sub, diff = dg.stacked_sequence_arrays(video, idx, bg, engine)
        for i, f in enumerate(idx):
            assert np.array_equal(sub[i], fo.subtract_scalar_background(frames[f], bg))
            want = np.zeros(frames[f].shape) if i == 0 else fo.frame_difference(frames[f], frames[idx[i - 1]], 0.0)
            assert np.array_equal(diff[i], want)
        dg.generate_stacked_sequence(video, idx, bg, tmp_path / "seq.png", title="run", figsize_width=12.0, engine=engine)
        log = fake_pyplot
        shown = calls(log, "imshow")
        assert len(shown) == 10
        assert np.array_equal(shown[2][0][0], sub[1]) and np.array_equal(shown[3][0][0], diff[1])
        assert calls(log, "savefig")[0][0][0] == tmp_path / "seq.png"
        del log[:]
        dg.generate_stacked_sequence_single_column(video, idx, bg, tmp_path / "one.png", use_frame_diff=True,
                                                   engine=engine)
        shown = calls(log, "imshow")
        assert len(shown) == 1 and np.array_equal(shown[0][0][0], diff.reshape(-1, c["width"]))
        assert len(calls(log, "text")) == 5 and len(calls(log, "axhline")) == 4


@pytest.mark.gpu
def test_driver_renders_one_figure_per_detect_call(engine, tmp_path, clip_small, golden, fake_pyplot):
    from high_speed_image_processing_b200.process_videos import VideoSourceConfig, process_video_source
    vdir = tmp_path / "videos"
    vdir.mkdir()
    (vdir / "run-3-.cihx").write_bytes(clip_small["cihx"])
    (vdir / "run-3-.mraw").write_bytes(clip_small["packed"].tobytes())
    cfg = VideoSourceConfig(name="Nova")
    cfg.detection_method = "head"
    cfg.calibration = 0.000833333
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    process_video_source(cfg, None, engine=engine, verbose=False, diagnostics=True)
    saved = [a[0] for a, _ in calls(fake_pyplot, "savefig")]
    frames_dir = tmp_path / "out" / "run-3--frames"
    assert saved[0] == frames_dir / "run-3--stacked-sequence.png" and saved[1] == frames_dir / "run-3--stacked-single.png"
    want = [frames_dir / f"Nova-Frame-{p['frame']:06d}.png" for p in golden["head_replay"]["per_frame"]]
    assert saved[2:] == want and frames_dir.is_dir()
