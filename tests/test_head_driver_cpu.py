"""Host logic of the HEAD driver (process_videos._process_video_head / _head_walk_range) on CPU, with
the oracle-backed engine stand-in of tests/test_sharding_gloo.py in place of the CUDA kernels: chunking
(tracker state and halo frame carried from chunk to chunk, halo skipping over skip_frames entries),
early stop of the uploads, and the result assembly."""
import numpy as np
import pytest

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200.photron import open_video
from high_speed_image_processing_b200.process_videos import VideoSourceConfig, _process_video_head
from oracle import flame_oracle as fo
from oracle import head_oracle as ho
from test_sharding_gloo import _OracleHeadEngine


class CountingEngine(_OracleHeadEngine):
    def __init__(self):
        self.uploaded = []

    def upload(self, raw):
        self.uploaded.append(int(np.asarray(raw).size))
        return super().upload(raw)


@pytest.fixture()
def clip(tmp_path):
    spec = syn.SyntheticSpec(width=160, height=12, n_frames=70, style="nova", t_enter=5.0, velocity=3.0,
                             tail_length=30.0, curvature_px=1.0, seed=17, bits=16)
    frames = syn.render_frames(spec)
    syn.write_clip(tmp_path, "run-1-", spec, frames=frames)
    return spec, frames, tmp_path / "run-1-.cihx"


def _run(path, monkeypatch, chunk_mb, skip=()):
    monkeypatch.setenv("FF_HEAD_CHUNK_MB", chunk_mb)
    cfg = VideoSourceConfig(name="t")
    cfg.detection_method = "head"
    cfg.skip_frames = list(skip)
    eng = CountingEngine()
    with open_video(str(path)) as video:
        res = _process_video_head(video, cfg, 0.000833333, 1.347567, eng, None)
    return res, eng


def test_chunked_walk_equals_the_oracle_loop_and_stops_uploading(clip, monkeypatch):
    spec, frames, path = clip
    time_of = lambda i: fo.frame_time_absolute(i, spec.start_frame, spec.skip_frame, spec.record_rate)  # noqa: E731
    want = ho.run_head(frames, spec.record_rate, 0.000833333, 1.347567, time_of)
    assert want.stop and want.stop[0] == "exit" and len(want.rows) > 20
    fb = spec.frame_bytes
    for chunk_mb in ("2048", "0"):
        res, eng = _run(path, monkeypatch, chunk_mb)
        assert [list(r) for r in res.rows] == want.rows and res.velocity_history == want.velocity_history
        assert res.stop == want.stop and res.first_exit == want.stop[1] and res.ddt_frame == want.ddt_frame
        assert res.empty_frames == want.empty
        assert (res.pos_px[want.stop[1]:] == -2).all()
        if chunk_mb == "0":           # 2-frame chunks: nothing after the chunk with the exit frame is uploaded
            assert sum(eng.uploaded) <= (want.stop[1] + 2 + 1) * fb, (sum(eng.uploaded) // fb, want.stop)
            assert max(eng.uploaded) == 2 * fb
        else:
            assert sum(eng.uploaded) == (spec.n_frames + 1) * fb      # frame 0 for the scalars + the clip


def test_chunk_boundaries_next_to_skipped_frames(clip, monkeypatch):
    """The halo of a chunk is the latest NON-skipped frame before it (:1443-1445), wherever the chunk
    boundary falls: 2-frame chunks and one upload must agree for skip lists that straddle boundaries."""
    _, _, path = clip
    for skip in ([9, 10], [10, 11, 12], [7, 8, 9, 10, 11], [1, 2, 3], [20, 22, 24, 26]):
        one, _ = _run(path, monkeypatch, "2048", skip)
        many, eng = _run(path, monkeypatch, "0", skip)
        assert one.rows == many.rows and one.velocity_history == many.velocity_history and one.stop == many.stop
        assert np.array_equal(one.pos_px, many.pos_px) and one.empty_frames == many.empty_frames
        assert all(frame not in [r[0] for r in one.rows] for frame in skip)
