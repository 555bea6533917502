"""Host-side mirror of the reference API (photron package, config dataclasses) against the
golden vectors produced by the reference's own classes.  CPU only - no frame compute."""
import numpy as np
import pytest

import high_speed_image_processing_b200 as hsip
from high_speed_image_processing_b200 import mraw, synthetic as syn
from high_speed_image_processing_b200.photron import (MPIVideoProcessor, PhotonVideo, SpatialCalibration, TimingInfo,
                                                      VideoCollection, open_collection, open_video, parse_cihx_xml)
from high_speed_image_processing_b200.photron.parallel import split_indices
from high_speed_image_processing_b200.process_videos import (FileCalibration, VideoSourceConfig, build_rows,
                                                             write_position_file)


class FakeComm:
    def __init__(self, rank, size):
        self._r, self._s = rank, size

    def Get_rank(self):
        return self._r

    def Get_size(self):
        return self._s


def test_public_names_match_reference_package():
    for name in ["PhotonVideo", "VideoCollection", "MetadataConfig", "MPIVideoProcessor", "SpatialCalibration",
                 "TimingInfo", "open_video", "open_collection"]:          # src/__init__.py:47-61
        assert hasattr(hsip, name)


def test_photon_video_metadata_and_timing(clip_small_on_disk, golden):
    g = golden["video"]
    video = open_video(str(clip_small_on_disk), calibration=SpatialCalibration(scale=0.000833333, units="m"))
    assert len(video) == g["len"]
    assert list(video.frame_shape) == g["frame_shape"]
    assert video.frame_rate == g["frame_rate"] and video.fps == g["frame_rate"]
    assert str(video.dtype) == g["dtype"]
    assert video.duration == g["duration"]
    assert video.has_absolute_timing == g["has_absolute_timing"]
    assert [video.get_absolute_time(i) for i in range(len(video))] == g["absolute_time"]   # bit-exact
    assert str(video.get_datetime(5)) == g["datetime_5"]
    assert [video.timing.time_to_frame(t) for t in (0.0, 1e-4, 3.3e-4)] == g["time_to_frame"]
    meta = {k: (str(v) if k == "recording_datetime" else v) for k, v in video.cihx_metadata.items()}
    assert meta == g["cihx_metadata"]
    assert video.storage_bits == 12 and video.bit_depth == 12
    assert video.width == 128 and video.height == 16
    assert video.pixels_to_physical(6) == 6 * 0.000833333
    video.set_trigger_frame(10)
    assert [video.get_time(i) for i in range(20)] == g["time_trigger10"]
    assert video.raw_frames(2, 4).size == 2 * 128 * 16 * 3 // 2
    video.close()
    with pytest.raises(ValueError):
        video.raw_frames(0, 1)


def test_photon_video_errors(clip_small_on_disk, tmp_path):
    with pytest.raises(FileNotFoundError):
        PhotonVideo(str(tmp_path / "missing.cihx"))
    video = open_video(str(clip_small_on_disk))
    with pytest.raises(IndexError):
        video[len(video)]
    with pytest.raises(IndexError):
        video[-len(video) - 1]
    with pytest.raises(TypeError):
        video["0"]
    with pytest.raises(ValueError):
        video.pixels_to_physical(1.0)        # no calibration set (reference video.py:694-695)
    assert TimingInfo(frame_rate=0).frame_to_time(5) == 0.0
    assert TimingInfo(frame_rate=0).frame_to_absolute_time(5) == 0.0
    assert TimingInfo(frame_rate=0).time_to_frame(1.0) == 0


def test_16bit_and_8bit_frame_access_is_a_plain_copy(tmp_path):
    for bits in (16, 8):
        spec = syn.SyntheticSpec(width=64, height=8, n_frames=6, bits=bits, seed=9)
        frames = syn.render_frames(spec)
        path = syn.write_clip(tmp_path, f"clip{bits}", spec, frames=frames)
        video = open_video(str(path))
        assert video.dtype == frames.dtype
        assert np.array_equal(video[3], frames[3])
        assert np.array_equal(video[-1], frames[-1])
        assert np.array_equal(video[1:4], frames[1:4])
        assert np.array_equal(video[::2], frames[::2])
        assert sum(1 for _ in video) == 6


def test_12bit_frame_access_needs_the_gpu(clip_small_on_disk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    video = open_video(str(clip_small_on_disk))
    with pytest.raises(RuntimeError):        # no CPU decoder on the product path
        video[0]


def test_cih_text_header(tmp_path):
    spec = syn.SyntheticSpec(width=64, height=8, n_frames=4, bits=16, seed=3, start_frame=-20)
    path = syn.write_clip(tmp_path, "legacy", spec, header="cih")
    video = open_video(str(path))
    assert len(video) == 4 and video.frame_shape == (8, 64)
    assert video.frame_rate == spec.record_rate          # from 'Record Rate(fps)' (video.py:360)
    assert video.timing.start_frame == -20               # from 'Start Frame' (video.py:367)
    assert not video.has_absolute_timing
    assert video.shutter_speed == pytest.approx(2.5e-6)


def test_header_validation(tmp_path):
    spec = syn.SyntheticSpec(width=64, height=8, n_frames=2, bits=16)
    bad = syn.cihx_bytes(spec).replace(b"<bit>16</bit>", b"<bit>10</bit>")
    (tmp_path / "bad.cihx").write_bytes(bad)
    (tmp_path / "bad.mraw").write_bytes(b"\0" * 4096)
    with pytest.raises(ValueError):
        mraw.load_video(tmp_path / "bad.cihx")
    (tmp_path / "nomraw.cihx").write_bytes(syn.cihx_bytes(spec))
    with pytest.raises(FileNotFoundError):
        mraw.load_video(tmp_path / "nomraw.cihx")
    (tmp_path / "short.cihx").write_bytes(syn.cihx_bytes(spec))
    (tmp_path / "short.mraw").write_bytes(b"\0" * 10)
    with pytest.raises(ValueError):
        mraw.load_video(tmp_path / "short.cihx")
    # parse_cihx_xml never raises (video.py:146-148)
    (tmp_path / "garbage.cihx").write_bytes(b"\x00\x01<cih><frameInfo><totalFrame>x</totalFrame></frameInfo></cih>")
    assert parse_cihx_xml(tmp_path / "garbage.cihx")["record_rate"] == 0
    assert parse_cihx_xml(tmp_path / "does-not-exist.cihx")["skip_frame"] == 1


def test_file_calibration_matches_reference(golden):
    rules = [FileCalibration(calibration=0.000833333, position_offset=1.0159, files=["run-1-"]),
             FileCalibration(calibration=0.000833333, position_offset=1.197565, files=["run-2-"]),
             FileCalibration(calibration=0.000833333, position_offset=1.347567, files=["run-3-:run-10-"])]
    cfg = VideoSourceConfig(name="Nova", calibration=1.0, position_offset=0.0, file_calibrations=rules)
    for name, want in golden["calibration_lookup"].items():
        assert list(cfg.get_calibration_for_file(name)) == want, name
    fc = FileCalibration(calibration=1.0, files=["Run-001:Run-005", "special", "A:B"])
    for name, want in golden["file_calibration_matches"].items():
        assert fc.matches(name) == want, name


def test_video_source_config_surface(golden):
    cfg = VideoSourceConfig(name="Nova")
    assert (cfg.enabled, cfg.calibration, cfg.position_offset, cfg.trigger_frame) == (False, 1.0, 0.0, None)
    assert cfg.use_frame_diff and cfg.use_absolute_time and cfg.skip_frames == [] and cfg.file_calibrations == []
    cfg.detection_method = "threshold"                      # README.md:55 attribute-assignment style
    assert cfg.detection_params().method == "threshold"
    cfg.detection_method = "sobel"
    with pytest.raises(ValueError):
        cfg.detection_params()
    cfg.video_path = "/abs/path"
    assert cfg.video_path == golden["abs_path_kept"]
    cfg.output_dir = "rel/out"
    assert cfg.output_dir.endswith("rel/out") and cfg.output_dir.startswith("/")
    cfg.video_path = None
    assert cfg.video_path is None


def test_distribute_indices_matches_reference(golden):
    for key, want in golden["distribute_indices"].items():
        total, size, strategy = key.split("/")
        got = [MPIVideoProcessor(FakeComm(r, int(size))).distribute_indices(int(total), strategy)
               for r in range(int(size))]
        assert got == want, key
    with pytest.raises(ValueError):
        split_indices(5, 0, 1, "zigzag")
    serial = MPIVideoProcessor(None)
    g = golden["serial_processor"]
    assert (serial.rank, serial.size, serial.is_root, serial.is_parallel) == (g["rank"], g["size"], g["is_root"],
                                                                              g["is_parallel"])
    assert serial.gather([1, 2]) == g["gather"] and serial.distribute_indices(5) == g["indices"]
    assert serial.broadcast("x") == "x" and serial.scatter([7]) == 7 and serial.scatter(None) is None
    serial.barrier()
    arr = np.arange(3.0)
    assert serial.reduce_sum(arr) is arr and serial.allreduce_sum(arr) is arr


def test_video_collection_matches_reference(tmp_path, golden):
    g = golden["collection"]
    syn.write_clip(tmp_path, "a_first", syn.SyntheticSpec(width=64, height=8, n_frames=5, bits=16, seed=5))
    syn.write_clip(tmp_path, "b_second", syn.SyntheticSpec(width=64, height=8, n_frames=9, bits=8, seed=6))
    (tmp_path / "c_broken.cihx").write_bytes(b"not a header")        # warning, not an error
    coll = open_collection(str(tmp_path))
    assert len(coll) == g["len"] and coll.total_frames == g["total_frames"]
    for k, want in g["resolve"].items():
        assert list(coll.global_to_local(int(k))) == want
    assert coll.local_to_global(1, 3) == g["local_to_global"]
    assert [p.name for p in coll.filepaths] == g["names"]
    assert [str(v.dtype) for v in coll] == g["dtypes"]
    import hashlib
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(coll.get_global_frame(4)) == g["frame_4_sha1"]
    assert sha(coll.get_global_frame(5)) == g["frame_5_sha1"]
    with pytest.raises(IndexError):
        coll.global_to_local(14)
    with pytest.raises(IndexError):
        coll.local_to_global(2, 0)
    seen = coll.map_frames(lambda frame, v, f: (v, f), frame_indices=[0, 5, 13])
    assert seen == [(0, 0), (1, 0), (1, 8)]
    assert len(coll.map_frames(lambda frame, v, f: 1, video_indices=[0])) == 5
    assert sum(1 for _ in coll.iter_frames()) == 14
    coll.set_calibration_all(2.0).set_trigger_frame_all(3)
    assert coll[0].calibration.scale == 2.0 and coll[1].trigger_frame == 3
    assert "2 videos" in coll.summary()
    proc = MPIVideoProcessor(None)
    assert proc.process_videos(coll, lambda video, idx: len(video)) == [(0, 5), (1, 9)]
    assert [g for g, _ in proc.process_collection(coll, lambda frame, g: 0)] == list(range(14))
    with pytest.raises(ValueError):
        open_collection(123)
    with pytest.raises(FileNotFoundError):
        VideoCollection.from_directory(tmp_path / "nope")
    coll.close_all()


def test_rows_and_result_file(clip_small_on_disk, golden, tmp_path):
    """Time_s / Position_m and the text formatting reproduce the reference's rows exactly."""
    video = open_video(str(clip_small_on_disk))
    ref_rows = golden["head_replay"]["results"]
    pos = np.full(len(video), -1, dtype=np.int32)
    for f, _, px, _, _ in ref_rows:
        pos[f] = px
    rows = build_rows(video, pos, 0.000833333, 1.347567, use_absolute_time=True)
    assert [[r[0], r[1], r[2], r[3], r[4]] for r in rows] == ref_rows          # float64 bit-exact
    out = write_position_file(rows, tmp_path / "x.txt")
    lines = open(out).read().splitlines()
    assert lines[0] == "#Frame Time_s Position_px Position_m"
    f, t, px, pm, _ = ref_rows[0]
    assert lines[1] == f"{f} {t:.9f} {px} {pm:.9f}"
    # README.md:93-96 with the README's own clip parameters
    t39 = TimingInfo(frame_rate=160000, start_frame=500, skip_frame=1).frame_to_absolute_time(39)
    assert f"{t39:.9f} {6 * 0.000833333 + 1.347567:.9f}" == "0.003368750 1.352566998"
