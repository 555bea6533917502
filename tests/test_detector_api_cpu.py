"""Frame-level API (seam B3): what can be checked without a GPU - the oracle against
the vectors recorded from the reference's FlameDetector / frame functions, the host-side pieces
of the drop-in (configuration, result type, write_results, velocity bookkeeping), and that the
GPU-backed functions refuse to run without the CUDA path instead of falling back."""
import hashlib

import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import detector as det
from high_speed_image_processing_b200 import process_videos as pv
from high_speed_image_processing_b200.head import VelocityBook
from oracle import flame_oracle as fo
from oracle import head_oracle as ho

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
    assert _sha(f) == api["frames_sha1"] and list(f.shape) == api["shape"]
    return f


def test_oracle_detector_reproduces_reference_calls(api, frames):
    for run in api["runs"]:
        cfg = ho.HeadConfig(**{k: v for k, v in run["cfg"].items() if k in ho.HeadConfig.__dataclass_fields__})
        orc = ho.FrameDetectorOracle(api["frame_rate"], run["calibration"], cfg,
                                     kernel_size=run["cfg"].get("morphology_kernel_size", 3))
        for call in run["calls"]:
            o = orc.detect(frames[call["frame"]], call["frame"], run["background"])
            assert (o["final"], o["min_gradient"], o["rightmost_sobel"], list(o["search"])) == \
                (call["final"], call["min_gradient"], call["rightmost_sobel"], call["search"]), (run["name"], call["frame"])
            for name in IMAGES:
                img = o["images"][name]
                assert (None if img is None else _sha(img)) == call["sha1"][name], (run["name"], call["frame"], name)
        assert orc.velocities == run["velocity_history"] and orc.ddt_frame == run["ddt_frame"]


def test_oracle_frame_functions_reproduce_reference(api, frames):
    ops = api["ops"]
    f5, f6, f7 = frames[5], frames[6], frames[7]
    sub6 = fo.subtract_scalar_background(f6, 41.5)
    assert _sha(sub6) == ops["sub_bg_41.5"]
    assert _sha(fo.subtract_scalar_background(f6, float(np.max(frames[0])))) == ops["sub_bg_max0"]
    assert _sha(fo.frame_difference(f6, f5, 0.0)) == ops["prior_raw_thr0"]
    assert _sha(fo.frame_difference(f6, f5, 7.5)) == ops["prior_raw_thr7.5"]
    assert _sha(fo.frame_difference(sub6, fo.subtract_scalar_background(f5, 41.5), 5.0)) == ops["prior_f64"]
    assert _sha(fo.three_frame_difference(f5, f6, f7)) == ops["three_thr0"]
    assert _sha(fo.three_frame_difference(f5, f6, f7, 3.0)) == ops["three_thr3"]
    for key, want in ops["empty"].items():
        thr, frac = (float(v) for v in key.split("/"))
        assert fo.is_empty_frame(sub6, thr, frac) == want, key
    assert fo.is_empty_frame(f6) == ops["empty_raw_defaults"]
    for thr, want in ops["count_above"].items():
        assert fo.nonempty_count(sub6, float(thr)) == want


def test_config_and_result_types_match_reference(api):
    cfg = pv.FlameDetectorConfig()
    assert {k: getattr(cfg, k) for k in cfg.__dataclass_fields__} == api["config_defaults"]
    assert list(pv.FlameDetectionResult.__dataclass_fields__) == api["result_fields"]
    for name in ("FlameDetector", "subtract_scalar_background", "subtract_prior_frame", "three_frame_difference",
                 "is_empty_frame", "write_results"):
        assert getattr(pv, name) is getattr(det, name)


def test_write_results_bytes(api, tmp_path):
    w = api["ops"]["write_results"]
    path = tmp_path / "out.txt"
    assert pv.write_results({k: v for k, v in w["columns"]}, str(path)) == str(path)
    assert path.read_bytes().decode("latin-1") == w["bytes"]


def test_velocity_book_replays_reference_histories(api):
    for run in api["runs"]:
        book = VelocityBook(api["frame_rate"], run["calibration"],
                            run["cfg"].get("ddt_velocity_jump_m_s", 1250.0))
        for call in run["calls"]:
            book.update(call["frame"], call["final"])
        assert book.velocities == run["velocity_history"]
        assert book.ddt_frame == run["ddt_frame"]
        assert book.last_detection()[1] == run["last_position"]
    book = VelocityBook(1000.0, 0.01, 5.0)
    for f, p in [(0, 10), (1, 11), (2, 12), (3, 2000), (4, None), (6, 2010)]:
        book.update(f, p)
    assert book.ddt_frame == 3                      # (2000-12)*0.01*1000 - 10 > 5
    assert book.velocities[1][3] == (2000 - 11) * 0.01 / (2 * 0.001)   # central difference filled in afterwards
    assert [e[0] for e in book.velocities] == [1, 2, 3]               # no entry across the missing position
    book.clear_last_central()
    assert book.velocities[-2][3] is None
    book.reset()
    assert book.history == [] and book.velocities == [] and book.ddt_frame is None and book.last_detection() == (-1, -1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without CUDA")
def test_frame_functions_have_no_cpu_fallback():
    frame = np.zeros((4, 8), dtype=np.uint16)
    for call in (lambda: pv.subtract_scalar_background(frame, 1.0), lambda: pv.is_empty_frame(frame),
                 lambda: pv.subtract_prior_frame(frame, frame), lambda: pv.three_frame_difference(frame, frame, frame),
                 lambda: pv.FlameDetector(pv.FlameDetectorConfig(), 1000.0, 0.001)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
