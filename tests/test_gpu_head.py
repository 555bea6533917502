"""HEAD-parity detector on the GPU (ff_head_lines / ff_head_track) against SciPy, the oracle
loop and the files written by the reference's own driver."""
import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200.head import HeadParams, finish_head_track
from high_speed_image_processing_b200.process_videos import FileCalibration, VideoSourceConfig, process_video_source
from oracle import flame_oracle as fo
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


def dev(a, engine):
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


def run_head_gpu(engine, frames, bits, hp, frame_rate=160000, cal=0.000833333, **kw):
    n, h, w = frames.shape
    return engine.process_head(dev(syn.pack_frames(frames, bits), engine), n, h, w, bits, hp, frame_rate, cal,
                               keep_lines=True, **kw)


def oracle_lines(frames, hp):
    """Per-frame (flag, sobel row, gradient row) via the reference's SciPy calls."""
    bg = fo.background_scalar(frames[0])
    noise = fo.empty_noise_threshold(bg)
    out, prior = [], None
    for i in range(len(frames)):
        sub = fo.subtract_scalar_background(frames[i], bg)
        if fo.is_empty_frame(sub, noise, hp.min_signal_fraction):
            out.append((0, None, None))
        elif prior is None:
            out.append((2, None, None))
        else:
            s, g = ho.detect_lines_scipy(fo.frame_difference(sub, prior, hp.frame_diff_threshold),
                                         sigma=hp.gaussian_sigma)
            out.append((1, s, g))
        prior = sub
    return out


@pytest.mark.parametrize("h,w,bits", [(16, 128, 12), (64, 512, 12), (6, 40, 12), (3, 34, 16), (19, 70, 8),
                                       (1, 64, 16), (2, 300, 12), (128, 1024, 12), (24, 257, 12)])
def test_band_lines_bit_exact_vs_scipy(engine, h, w, bits):
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=36, bits=bits, style="mini", t_enter=4.0,
                             velocity=max(1.0, w / 40.0), curvature_px=1.0, seed=h * w)
    frames = syn.render_frames(spec)
    hp = HeadParams()
    res = run_head_gpu(engine, frames, bits, hp)
    flags = res.flags.cpu().numpy()
    lines = res.lines.cpu().numpy()
    want = oracle_lines(frames, hp)
    assert sum(1 for f, _, _ in want if f == 1) > 5, "fixture must exercise the detector"
    for i, (flag, s, g) in enumerate(want):
        assert flags[i] == flag, i
        if flag == 1:
            assert np.array_equal(lines[i, 0], s), (i, np.abs(lines[i, 0] - s).max())
            assert np.array_equal(lines[i, 1], g), (i, np.abs(lines[i, 1] - g).max())


@pytest.mark.parametrize("sigma", [1.0, 1.5, 2.0])
def test_other_sigmas(engine, sigma):
    frames = syn.render_frames(syn.SyntheticSpec(width=200, height=32, n_frames=24, style="nova", t_enter=3.0,
                                                 velocity=6.0, seed=9))
    hp = HeadParams(gaussian_sigma=sigma)
    res = run_head_gpu(engine, frames, 12, hp)
    lines = res.lines.cpu().numpy()
    for i, (flag, s, g) in enumerate(oracle_lines(frames, hp)):
        if flag == 1:
            assert np.array_equal(lines[i, 0], s) and np.array_equal(lines[i, 1], g)


@pytest.mark.parametrize("style,w,h,vel", [("nova", 128, 16, 2.0), ("mini", 512, 64, 5.0), ("mini", 300, 10, 3.0)])
def test_tracker_matches_oracle_loop(engine, style, w, h, vel):
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=int(w / vel) + 30, style=style, t_enter=6.0, velocity=vel,
                             tail_length=40.0, curvature_px=2.0, seed=w + h)
    frames = syn.render_frames(spec)
    hp = HeadParams()
    cal, off, rate = 0.000833333, 1.347567, 160000
    time_of = lambda i: fo.frame_time_absolute(i, 500, 1, rate)
    res = run_head_gpu(engine, frames, 12, hp, frame_rate=rate, cal=cal)
    got = finish_head_track(res.track.cpu().numpy(), res.flags.cpu().numpy(), 0, w, rate, cal, off, time_of, hp)
    want = ho.run_head(frames, rate, cal, off, time_of)
    assert got.per_frame == want.per_frame
    assert [list(r) for r in got.rows] == want.rows
    assert got.velocity_history == want.velocity_history
    assert got.ddt_frame == want.ddt_frame and got.stop == want.stop
    stop = res.stop.cpu().numpy()
    if want.stop and want.stop[0] == "exit":
        assert stop[0] == want.stop[1]


def test_reference_golden_replay(engine, clip_small, golden):
    c = golden["clip_small"]
    hr = golden["head_replay"]
    n, h, w = c["n_frames"], c["height"], c["width"]
    res = engine.process_head(dev(clip_small["packed"], engine), n, h, w, 12, HeadParams(), 160000, 0.000833333)
    got = finish_head_track(res.track.cpu().numpy(), res.flags.cpu().numpy(), 0, w, 160000, 0.000833333, 1.347567,
                            lambda i: fo.frame_time_absolute(i, 500, 1, 160000), HeadParams())
    assert got.per_frame == [{k: v for k, v in p.items() if k != "diff_sha1"} for p in hr["per_frame"]]
    assert [list(r) for r in got.rows] == hr["results"]
    assert got.velocity_history == hr["velocity_history"] and list(got.stop) == hr["stop"]


@pytest.mark.parametrize("chunk_mb", [None, "0"])
def test_driver_writes_the_reference_files_byte_for_byte(tmp_path, clip_small, golden, monkeypatch, chunk_mb):
    """process_video_source(detection_method='head') on the same recordings the reference's own
    driver processed (oracle/make_golden.py): every output file must be identical - with the clip
    uploaded at once and (FF_HEAD_CHUNK_MB=0) in 2-frame chunks, tracker state and halo frame carried
    from chunk to chunk and no upload past the exit."""
    if chunk_mb is not None:
        monkeypatch.setenv("FF_HEAD_CHUNK_MB", chunk_mb)
    vdir = tmp_path / "Nova-Video-Files"
    vdir.mkdir()
    for stem in ("run-3-", "run-5-_C001H001S0001"):
        (vdir / f"{stem}.cihx").write_bytes(clip_small["cihx"])
        (vdir / f"{stem}.mraw").write_bytes(clip_small["packed"].tobytes())
    cfg = VideoSourceConfig(name="Nova")
    cfg.enabled = True
    cfg.detection_method = "head"
    cfg.calibration = golden["driver_config"]["calibration"]
    cfg.position_offset = golden["driver_config"]["position_offset"]
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    cfg.file_calibrations = [FileCalibration(calibration=0.000833333, position_offset=1.347567,
                                             files=["run-3-:run-10-"])]
    process_video_source(cfg, None, verbose=False)
    produced = {p.name: p.read_text() for p in (tmp_path / "out").glob("*.txt")}
    assert produced == golden["driver_outputs"]


def test_head_with_skip_frames_and_no_detection(engine):
    frames = syn.render_frames(syn.SyntheticSpec(width=128, height=16, n_frames=40, style="mini", t_enter=5.0,
                                                 velocity=3.0, seed=4))
    dark = np.full((12, 16, 128), 40, dtype=np.uint16)
    res = run_head_gpu(engine, dark, 12, HeadParams())
    assert res.flags.cpu().numpy().tolist() == [0] * 12
    assert (res.track.cpu().numpy() == -1).all()
    mask = np.zeros(len(frames), dtype=np.uint8)
    mask[[10, 11, 20]] = 1
    res = run_head_gpu(engine, frames, 12, HeadParams(), skip=dev(mask, engine))
    flags = res.flags.cpu().numpy()
    assert flags[10] == 0 and flags[11] == 0 and flags[20] == 0
    # frame 12's prior is frame 9 (the latest non-skipped frame)
    bg = fo.background_scalar(frames[0])
    d = fo.frame_difference(fo.subtract_scalar_background(frames[12], bg),
                            fo.subtract_scalar_background(frames[9], bg), 5.0)
    s, g = ho.detect_lines_scipy(d)
    lines = res.lines.cpu().numpy()
    assert np.array_equal(lines[12, 0], s) and np.array_equal(lines[12, 1], g)


def _track_both_ways(engine, packed, n, h, w, hp, rate, cal, monkeypatch, **kw):
    """(speculative, sequential) tracker outputs on the same lines.  The speculative tracker is run in
    both of its forms - chained guesses checked in parallel (default) and segment-by-segment
    validation (FF_TRACK_UNCHAINED=1) - which must agree before either is compared."""
    def run():
        r = engine.process_head(packed, n, h, w, 12, hp, rate, cal, **kw)
        return r.track.cpu().numpy(), r.stop.cpu().numpy(), r.flags.cpu().numpy()

    for var in ("FF_TRACK_SEQUENTIAL", "FF_TRACK_UNCHAINED"):
        monkeypatch.delenv(var, raising=False)
    spec_out = run()
    monkeypatch.setenv("FF_TRACK_UNCHAINED", "1")
    unchained = run()
    monkeypatch.delenv("FF_TRACK_UNCHAINED", raising=False)
    for a, b in zip(spec_out, unchained):
        assert np.array_equal(a, b), "chained and unchained speculation disagree"
    monkeypatch.setenv("FF_TRACK_SEQUENTIAL", "1")
    seq_out = run()
    monkeypatch.delenv("FF_TRACK_SEQUENTIAL", raising=False)
    return spec_out, seq_out


@pytest.mark.parametrize("case", ["long_flame", "slow_no_exit", "noise", "late_start", "noise_gaps", "two_fronts",
                                  "two_fronts_exit", "exit_then_noise"])
def test_speculative_tracker_equals_sequential_walk(engine, monkeypatch, case):
    """The speculative parallel tracker (32 segments walked at once, then validated in order) must
    reproduce the sequential walk exactly - several batches of 1024 frames, segments that do not
    lock on at once, an exit in the middle of a batch, frames without detections."""
    rng = np.random.default_rng(7)
    if case == "long_flame":        # ~2700 flame frames: 3 batches, exit well inside the last one
        spec = syn.SyntheticSpec(width=1024, height=32, n_frames=3000, style="nova", t_enter=40.0, velocity=0.37,
                                 seed=21)
        frames = syn.render_frames(spec)
    elif case == "slow_no_exit":    # the front never reaches the exit margin
        spec = syn.SyntheticSpec(width=512, height=24, n_frames=1500, style="mini", t_enter=100.0, velocity=0.2,
                                 seed=22)
        frames = syn.render_frames(spec)
    elif case == "late_start":      # long empty lead-in (batches without active frames), then the flame
        spec = syn.SyntheticSpec(width=256, height=16, n_frames=2600, style="mini", t_enter=2300.0, velocity=1.1,
                                 seed=23)
        frames = syn.render_frames(spec)
    elif case == "two_fronts":      # a second, brighter front far ahead appears and vanishes: guesses miss
        spec = syn.SyntheticSpec(width=1024, height=16, n_frames=1200, style="mini", t_enter=30.0, velocity=0.6,
                                 seed=24)
        frames = syn.render_frames(spec)
        for i in range(200, 1100, 37):
            x = 400 + (i * 7) % 500
            frames[i:i + 3, :, x:x + 12] = np.minimum(frames[i:i + 3, :, x:x + 12].astype(np.int64) + 3000, 4095)
    elif case == "two_fronts_exit":  # missed guesses first, then the front leaves the domain
        spec = syn.SyntheticSpec(width=1024, height=16, n_frames=1200, style="mini", t_enter=30.0, velocity=1.0,
                                 seed=25)
        frames = syn.render_frames(spec)
        for i in range(150, 700, 41):
            x = 300 + (i * 11) % 600
            frames[i:i + 3, :, x:x + 12] = np.minimum(frames[i:i + 3, :, x:x + 12].astype(np.int64) + 3000, 4095)
    elif case == "exit_then_noise":  # clean walk to the exit; erratic blobs afterwards must not matter
        spec = syn.SyntheticSpec(width=512, height=16, n_frames=900, style="mini", t_enter=20.0, velocity=1.3,
                                 seed=26)
        frames = syn.render_frames(spec)
        for i in range(spec.exit_frame(15) + 5, 900):
            x = int(rng.integers(0, 480))
            frames[i, :, x:x + int(rng.integers(4, 25))] = 3500
    elif case == "noise_gaps":      # erratic blobs with empty stretches: walks fall back to "nothing" often
        frames = rng.integers(30, 60, size=(1400, 16, 256)).astype(np.uint16)
        for i in range(1, len(frames)):
            if (i // 9) % 3 == 1:
                continue
            x = int(rng.integers(0, 236))
            frames[i, :, x:x + int(rng.integers(3, 20))] += int(rng.integers(200, 3500))
    else:                           # bright random blobs: every frame non-empty, erratic detections
        frames = rng.integers(30, 60, size=(1300, 16, 256)).astype(np.uint16)
        for i in range(1, len(frames)):
            x = int(rng.integers(0, 230))
            frames[i, :, x:x + int(rng.integers(4, 25))] += int(rng.integers(300, 3000))
    n, h, w = frames.shape
    packed = dev(syn.pack_frames(frames, 12), engine)
    hp = HeadParams(exit_margin_px=15 if not case.startswith("noise") else 1)
    (t1, s1, f1), (t2, s2, f2) = _track_both_ways(engine, packed, n, h, w, hp, 160000, 0.000833333, monkeypatch)
    assert np.array_equal(f1, f2)
    assert np.array_equal(s1, s2), (s1, s2)
    assert np.array_equal(t1, t2), np.nonzero((t1 != t2).any(axis=1))[0][:10]
    if case == "long_flame":
        assert (f1 == 1).sum() > 2048 and s1[0] != 2**31 - 1
    if case == "slow_no_exit":
        assert s1[0] == 2**31 - 1 and (t1[:, 0] >= 0).sum() > 100
    if case in ("two_fronts_exit", "exit_then_noise"):
        assert s1[0] != 2**31 - 1 and (t1[s1[0] + 1:] == -1).all()
    # carried-in tracker state (a range that continues an earlier one)
    (t3, s3, _), (t4, s4, _) = _track_both_ways(engine, packed, n, h, w, hp, 160000, 0.000833333, monkeypatch,
                                                tracker_state=(3, 17))
    assert np.array_equal(t3, t4) and np.array_equal(s3, s4)
    # ... also one that lies in the exit zone (detected before this range: it must not end the walk)
    fb = h * w * 3 // 2
    (t5, s5, _), (t6, s6, _) = _track_both_ways(engine, packed[fb:], n - 1, h, w, hp, 160000, 0.000833333, monkeypatch,
                                                frame0=packed[:fb], halo=packed[:fb], first_frame=1000,
                                                tracker_state=(990, w - 2))
    assert np.array_equal(t5, t6) and np.array_equal(s5, s6)


def test_fullsize_c2_head_detector_vs_oracle_on_every_flame_frame(engine):
    """BASELINE config 2 at full size (1024x128 x 20000) through the HEAD detector: the lead-in must
    be skipped as empty, and from twelve frames before the flame enters to the stop every detect()
    output, result row and velocity equals the oracle loop (SciPy images, ~1100 frames)."""
    spec = syn.config_spec("C2")
    n, h, w, fb = spec.n_frames, spec.height, spec.width, spec.frame_bytes
    packed = syn.render_packed_torch(spec, engine.device)
    hp = HeadParams()
    cal, off, rate = 0.000833333, 1.347567, spec.record_rate
    time_of = lambda i: fo.frame_time_absolute(i, spec.start_frame, spec.skip_frame, rate)
    res = engine.process_head(packed, n, h, w, 12, hp, rate, cal)
    flags = res.flags.cpu().numpy()
    got = finish_head_track(res.track.cpu().numpy(), flags, 0, w, rate, cal, off, time_of, hp)
    assert got.stop is not None and got.stop[0] == "exit" and len(got.rows) > 900
    a = int(spec.t_enter) - 12                   # the front's edge (sigma 3 px) is still far outside the frame
    b = min(n, got.stop[1] + 3)
    assert not flags[:a].any(), "the lead-in holds no frame that reaches the detector"
    frame0 = fo.frames_from_bytes(packed[:fb].cpu().numpy(), 1, h, w, 12)
    window = fo.frames_from_bytes(packed[a * fb:b * fb].cpu().numpy(), b - a, h, w, 12)
    shift = a - 1                                   # oracle index i >= 1  <->  clip frame i + shift
    want = ho.run_head(np.concatenate([frame0, window]), rate, cal, off, lambda i: time_of(i + shift))
    assert want.empty >= 1 and want.stop == ("exit", got.stop[1] - shift)
    assert got.per_frame == [dict(p, frame=p["frame"] + shift) for p in want.per_frame]
    assert [list(r) for r in got.rows] == [[r[0] + shift] + r[1:] for r in want.rows]
    assert got.velocity_history == [[e[0] + shift] + e[1:] for e in want.velocity_history]
    assert got.ddt_frame == (None if want.ddt_frame is None else want.ddt_frame + shift)
    assert int(res.stop.cpu()[0]) == got.stop[1]


@pytest.mark.parametrize("h,w,bits,sigma", [(64, 512, 12, 1.5), (9, 264, 16, 2.0), (128, 1024, 12, 1.5), (20, 8, 8, 1.0),
                                            (33, 776, 8, 1.5), (40, 1280, 16, 0.7)])
def test_fast_band_kernel_equals_general_kernel(engine, monkeypatch, h, w, bits, sigma):
    """Rows that start on an 8-pixel boundary take head_band_fast_kernel (grouped loads, separable SIMD
    opening); FF_BAND_GENERAL=1 forces the per-pixel kernel.  Same lines, bit for bit, including
    skip_frames entries and a sub-range with a halo frame."""
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=30, bits=bits, style="nova", t_enter=2.0,
                             velocity=max(1.0, w / 24.0), tail_length=w / 6.0, seed=w + h + bits)
    frames = syn.render_frames(spec)
    frames[11] = np.random.default_rng(w).integers(0, spec.max_value + 1, size=(h, w)).astype(frames.dtype)
    packed = dev(syn.pack_frames(frames, bits), engine)
    fb = spec.frame_bytes
    skip = np.zeros(29, dtype=np.uint8)
    skip[[6, 7, 15]] = 1
    hp = HeadParams(gaussian_sigma=sigma)

    def run():
        r = engine.process_head(packed[fb:], 29, h, w, bits, hp, 160000, 0.000833333, frame0=packed[:fb], first_frame=1,
                                halo=packed[:fb], skip=dev(skip, engine), keep_lines=True)
        fl = r.flags.cpu().numpy()
        return fl, r.lines.cpu().numpy()[fl == 1], r.track.cpu().numpy()

    fast = run()
    monkeypatch.setenv("FF_BAND_GENERAL", "1")
    general = run()
    assert (fast[0] == 1).sum() > 8
    for a, b in zip(fast, general):
        assert np.array_equal(a, b)


def test_chunked_head_walk_with_skip_frames_stops_uploading_at_the_exit(engine, tmp_path, monkeypatch):
    """Chunk boundaries next to skip_frames entries (the halo is the latest non-skipped frame), and the
    upload counter: frames after the chunk that holds the exit never reach the device."""
    from high_speed_image_processing_b200.photron import open_video
    from high_speed_image_processing_b200.process_videos import process_video
    spec = syn.SyntheticSpec(width=256, height=16, n_frames=400, style="mini", t_enter=10.0, velocity=2.0, seed=31)
    frames = syn.render_frames(spec)
    syn.write_clip(tmp_path, "run-1-", spec, frames=frames)
    cfg = VideoSourceConfig(name="t")
    cfg.detection_method = "head"
    cfg.skip_frames = [15, 16, 31, 32, 33, 64]
    uploaded = []
    real_upload = engine.upload
    monkeypatch.setattr(engine, "upload", lambda host: (uploaded.append(np.asarray(host).size), real_upload(host))[1])
    results = {}
    for mb in ("2048", "0"):
        monkeypatch.setenv("FF_HEAD_CHUNK_MB", mb)
        del uploaded[:]
        with open_video(str(tmp_path / "run-1-.cihx")) as video:
            results[mb] = process_video(video, cfg, 0.000833333, 1.347567, engine=engine)
        results[mb + "_bytes"] = sum(uploaded)
    one, many = results["2048"], results["0"]
    assert one.rows == many.rows and one.velocity_history == many.velocity_history and one.stop == many.stop
    assert one.stop[0] == "exit" and len(one.rows) > 50 and np.array_equal(one.pos_px, many.pos_px)
    assert one.empty_frames == many.empty_frames
    fb = spec.frame_bytes
    assert results["2048_bytes"] >= 400 * fb                       # one chunk: the whole clip
    assert results["0_bytes"] < (one.stop[1] + 8) * fb * 1.6       # 2-frame chunks (+ halos): nothing past the exit


@pytest.mark.parametrize("bits", [16, 8])
def test_driver_files_on_16bit_and_8bit_recordings_match_the_reference(tmp_path, golden, bits):
    """The reference's own driver was run on 16-bit and 8-bit recordings (oracle/make_golden.py,
    `driver_other_depths`); the recordings are regenerated here from the same spec (frame sha1
    checked) and every result file must come out byte for byte."""
    import hashlib
    g = golden["driver_other_depths"][str(bits)]
    spec = syn.SyntheticSpec(**g["spec"])
    frames = syn.render_frames(spec)
    assert hashlib.sha1(np.ascontiguousarray(frames).tobytes()).hexdigest() == g["frames_sha1"]
    vdir = tmp_path / "Mini-Video-Files"
    syn.write_clip(vdir, g["stem"], spec, frames=frames)
    cfg = VideoSourceConfig(name="Mini")
    cfg.enabled = True
    cfg.detection_method = "head"
    cfg.calibration = g["calibration"]
    cfg.position_offset = g["position_offset"]
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    process_video_source(cfg, None, verbose=False)
    produced = {p.name: p.read_text() for p in (tmp_path / "out").glob("*.txt")}
    assert produced == g["outputs"]


@pytest.mark.parametrize("chunk_mb", [None, "0"])
def test_driver_files_with_skip_frames_match_the_reference(tmp_path, clip_small, golden, monkeypatch, chunk_mb):
    """skip_frames in the reference's driver (:1443-1445): skipped frames neither reach the detector nor
    become the prior frame.  Same recording, same list, byte-identical result files - also when the
    clip goes through the device in 2-frame chunks (the halo must skip over them)."""
    if chunk_mb is not None:
        monkeypatch.setenv("FF_HEAD_CHUNK_MB", chunk_mb)
    g = golden["driver_outputs_skip"]
    vdir = tmp_path / "Nova-Video-Files"
    vdir.mkdir()
    (vdir / "run-3-.cihx").write_bytes(clip_small["cihx"])
    (vdir / "run-3-.mraw").write_bytes(clip_small["packed"].tobytes())
    cfg = VideoSourceConfig(name="Nova")
    cfg.enabled = True
    cfg.detection_method = "head"
    cfg.calibration = 0.000833333
    cfg.position_offset = 1.347567
    cfg.skip_frames = list(g["skip_frames"])
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    process_video_source(cfg, None, verbose=False)
    produced = {p.name: p.read_text() for p in (tmp_path / "out").glob("*.txt")}
    assert produced == g["outputs"]


@pytest.mark.parametrize("event", ["ddt", "velocity_drop"])
def test_driver_files_on_ddt_and_velocity_drop_recordings_match_the_reference(tmp_path, golden, event):
    """Recordings on which the reference's driver saw a DDT event (velocity jump > 1250 m/s: pre- and
    post-DDT files, :506-516) and a front slowing to under half its speed (the velocity-drop stop,
    :1499-1509): two constant-velocity pieces of the synthetic generator spliced at a frame, rebuilt
    here (frame sha1 checked); every result file byte for byte."""
    import hashlib
    g = golden["driver_events"][event]
    sa, sb = syn.SyntheticSpec(**g["spec_a"]), syn.SyntheticSpec(**g["spec_b"])
    frames = np.concatenate([syn.render_frames(sa, 0, g["splice"]), syn.render_frames(sb, g["splice"], sa.n_frames)])
    assert hashlib.sha1(np.ascontiguousarray(frames).tobytes()).hexdigest() == g["frames_sha1"]
    vdir = tmp_path / "Nova-Video-Files"
    syn.write_clip(vdir, "run-3-", sa, frames=frames)
    cfg = VideoSourceConfig(name="Nova")
    cfg.enabled = True
    cfg.detection_method = "head"
    cfg.calibration = 0.000833333
    cfg.position_offset = 1.347567
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    res = process_video_source(cfg, None, verbose=False)["run-3-.cihx"]
    produced = {p.name: p.read_text() for p in (tmp_path / "out").glob("*.txt")}
    assert produced == g["outputs"]
    if event == "ddt":
        assert res.ddt_frame is not None and any("post-DDT" in k for k in produced) and res.stop[0] == "exit"
    else:
        assert res.stop[0] == "velocity_drop" and res.first_exit is None


def test_driver_files_with_trigger_relative_time_match_the_reference(tmp_path, clip_small, golden):
    """use_absolute_time = False with a trigger frame (:1449-1452): the Time_s column is
    (frame - trigger) / rate, negative before the trigger."""
    g = golden["driver_outputs_trigger"]
    vdir = tmp_path / "Nova-Video-Files"
    vdir.mkdir()
    (vdir / "run-3-.cihx").write_bytes(clip_small["cihx"])
    (vdir / "run-3-.mraw").write_bytes(clip_small["packed"].tobytes())
    cfg = VideoSourceConfig(name="Nova")
    cfg.enabled = True
    cfg.detection_method = "head"
    cfg.calibration = 0.000833333
    cfg.position_offset = 1.347567
    cfg.use_absolute_time = False
    cfg.trigger_frame = g["trigger_frame"]
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    process_video_source(cfg, None, verbose=False)
    produced = {p.name: p.read_text() for p in (tmp_path / "out").glob("*.txt")}
    assert produced == g["outputs"]


@pytest.mark.parametrize("run_name", ["defaults", "k5_sigma1_thr3.5_gaps", "k1_sigma2"])
def test_whole_clip_kernels_reproduce_the_reference_detector_runs(engine, golden, run_name):
    """The reference's FlameDetector runs recorded in the golden file (tests/golden/reference_golden.json,
    ``detector_api``: defaults; 5x5 opening / sigma 1 / threshold 3.5 / frame gaps; 1x1 opening / sigma 2)
    through the WHOLE-CLIP kernels - ff_head_lines (band kernel with a k x k opening) + ff_head_track - instead
    of one detect() call per frame: frames that were not fed to the detector are skip_frames entries, the
    background scalar is the one the reference was called with, no frame counts as empty."""
    from conftest import GOLDEN
    from high_speed_image_processing_b200.head import HeadParams, max_displacement_px
    api = golden["detector_api"]
    frames = np.load(GOLDEN / "detector_frames.npz")["frames"]
    run = next(r for r in api["runs"] if r["name"] == run_name)
    cfg = run["cfg"]
    hp = HeadParams(frame_diff_threshold=cfg.get("frame_diff_threshold", 5.0),
                    morphology_kernel_size=cfg.get("morphology_kernel_size", 3),
                    gaussian_sigma=cfg.get("gaussian_sigma", 1.5),
                    min_gradient_strength=cfg.get("min_gradient_strength", 10.0),
                    edge_margin_px=cfg.get("edge_margin_px", 10),
                    sobel_threshold_fraction=cfg.get("sobel_threshold_fraction", 0.1),
                    search_window_px=cfg.get("search_window_px", 100), min_signal_fraction=0.0)
    n, h, w = frames.shape
    order = run["order"]
    first, last = order[0], order[-1] + 1
    skip = np.ones(n, np.uint8)
    skip[order] = 0
    packed = torch.from_numpy(syn.pack_frames(frames, 16)).to(engine.device)
    fb = h * w * 2
    # ff_background takes the max of "frame 0": a frame filled with the background scalar of the recorded run
    frame0 = torch.from_numpy(np.full((h, w), int(run["background"]), np.uint16).view(np.uint8).reshape(-1)).to(engine.device)
    lines, flags, pending = engine.head_lines(packed[first * fb:last * fb], last - first, h, w, 16, hp, frame0=frame0,
                                              first_frame=first, skip=torch.from_numpy(skip[first:last].copy()).to(engine.device))
    assert engine.head_scalars(pending).background == run["background"]
    maxdisp = max_displacement_px(api["frame_rate"], run["calibration"], hp)
    assert maxdisp == run["max_displacement_px"]
    track, stop = engine.head_track_lines(lines, flags, first, w, hp, maxdisp)
    track, flags = track.cpu().numpy(), flags.cpu().numpy()
    none = lambda v: None if v < 0 else int(v)      # noqa: E731
    for call in run["calls"]:
        i = call["frame"] - first
        got = track[i]
        assert flags[i] == (2 if call["sha1"]["frame_diff"] is None else 1), (run_name, call["frame"])
        if flags[i] == 2:             # the first frame fed: no prior frame, detect() returns no position
            assert call["final"] is None
            continue
        assert (none(got[0]), none(got[1]), none(got[2])) == (call["final"], call["min_gradient"], call["rightmost_sobel"]), \
            (run_name, call["frame"])
        assert [int(got[3]), int(got[4])] == call["search"], (run_name, call["frame"])
    assert (flags[skip[first:last] == 1] == 0).all()
