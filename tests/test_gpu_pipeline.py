"""End-to-end drop-in path (files on disk -> PhotonVideo -> process_video_source) and
full-size property checks on the BASELINE configurations."""
import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED
from high_speed_image_processing_b200.engine import DetectionParams
from high_speed_image_processing_b200.photron import open_video
from high_speed_image_processing_b200.process_videos import (FileCalibration, VideoSourceConfig, process_video,
                                                             process_video_source)
from oracle import flame_oracle as fo

pytestmark = pytest.mark.gpu


def test_12bit_frame_access_uses_the_unpack_kernel(clip_small_on_disk, clip_small, golden):
    c = golden["clip_small"]
    want = fo.frames_from_bytes(clip_small["packed"], c["n_frames"], c["height"], c["width"], 12)
    video = open_video(str(clip_small_on_disk))
    assert np.array_equal(video[0], want[0])
    assert np.array_equal(video[-1], want[-1])
    assert np.array_equal(video[3:9], want[3:9])
    assert np.array_equal(video[::7], want[::7])
    assert video[5].dtype == np.uint16 and video[5].flags["OWNDATA"]


@pytest.mark.parametrize("method,style,residency", [("half_maximum", "nova", "device"), ("threshold", "mini", "host"),
                                                    ("gradient", "mini", "device")])
def test_process_video_source_matches_oracle_rows(tmp_path, method, style, residency):
    spec = syn.config_spec("C1", n_frames=160)
    spec = syn.SyntheticSpec(**{**spec.__dict__, "style": style, "velocity": 4.0, "t_enter": 12.0})
    frames = syn.render_frames(spec)
    vdir = tmp_path / "Nova-Video-Files"
    syn.write_clip(vdir, "run-3-", spec, frames=frames)
    cfg = VideoSourceConfig(name="Nova")
    cfg.enabled = True
    cfg.detection_method = method
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    cfg.file_calibrations = [FileCalibration(calibration=0.000833333, position_offset=1.347567,
                                             files=["run-3-:run-10-"])]
    want = fo.process_clip(frames, fo.ClipParams(method=method))
    if residency == "host":
        video = open_video(str(vdir / "run-3-.cihx"))
        res = process_video(video, cfg, 0.000833333, 1.347567, residency="host")
    else:
        res = process_video_source(cfg, None, verbose=False)["run-3-.cihx"]
    assert res.first_exit == (want.first_exit if want.first_exit < len(frames) else None)
    assert [(r[0], r[2]) for r in res.rows] == want.records                       # Position_px bit-exact
    for f, t, px, pm, _ in res.rows:
        assert t == fo.frame_time_absolute(f, spec.start_frame, spec.skip_frame, spec.record_rate)
        assert pm == fo.position_m(px, 0.000833333, 1.347567)
    assert res.empty_frames == int(want.empty[:want.first_exit].sum())
    if residency != "host":
        lines = (tmp_path / "out" / "run-3--flame-position.txt").read_text().splitlines()
        assert len(lines) == 1 + len(want.records)
        f, px = want.records[0]
        assert lines[1].split()[0] == str(f) and lines[1].split()[2] == str(px)


def _fullsize_checks(engine, name, method, n_frames):
    """Size-independent properties on a BASELINE-shaped clip generated on the GPU."""
    spec = syn.config_spec(name, n_frames=n_frames)
    h, w, fb = spec.height, spec.width, spec.frame_bytes
    packed = syn.render_packed_torch(spec, engine.device)
    scalars, bg_dev = engine.clip_scalars(packed[:fb], h, w, 12)
    params = DetectionParams(method=method)
    full = engine.process_range(packed, n_frames, h, w, 12, params, scalars, bg_dev)
    pos = full.pos.cpu().numpy()
    fe = int(full.first_exit.cpu().item())
    # (1) truncation: nothing survives at/after the exit frame, exit frame is really an exit
    if fe != FF_NO_EXIT:
        assert (pos[fe:] == FF_POS_DROPPED).all() and (pos[:fe] != FF_POS_DROPPED).all()
    # (2) range-split invariance with one-frame halos (the multi-GPU decomposition)
    cuts = [0, n_frames // 3 + 1, (2 * n_frames) // 3 - 1, n_frames]
    pieces, exits = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        halo = None if a == 0 else packed[(a - 1) * fb:a * fb]
        r = engine.process_range(packed[a * fb:b * fb], b - a, h, w, 12, params, scalars, bg_dev, first_frame=a,
                                 halo=halo, truncate=False)
        pieces.append(r.pos.cpu().numpy())
        exits.append(int(r.first_exit.cpu().item()))
    joined = np.concatenate(pieces)
    assert min(exits) == fe
    lim = n_frames if fe == FF_NO_EXIT else fe
    assert np.array_equal(joined[:lim], pos[:lim])
    # (3) host-streamed path == device-resident path
    host = engine.process_host(packed.cpu().numpy(), n_frames, h, w, 12, params, scalars)
    assert host.first_exit == fe and np.array_equal(host.pos, pos)
    # (4) sampled frames against the oracle (frames around the flame and in the lead-in)
    rng = np.random.default_rng(0)
    t_in = int(spec.t_enter)
    sample = sorted(set([1, 2, n_frames - 1, max(1, t_in - 1), t_in + 5, t_in + 50, min(lim - 1, t_in + 400)]
                        + rng.integers(1, n_frames, size=24).tolist()))
    frame0 = fo.frames_from_bytes(packed[:fb].cpu().numpy(), 1, h, w, 12)[0]
    unt = engine.process_range(packed, n_frames, h, w, 12, params, scalars, bg_dev, truncate=False)
    upos, ucnt = unt.pos.cpu().numpy(), unt.counts.cpu().numpy()
    for f in sample:
        two = fo.frames_from_bytes(packed[(f - 1) * fb:(f + 1) * fb].cpu().numpy(), 2, h, w, 12)
        o = fo.process_clip(two[1:], fo.ClipParams(method=method), frame0=frame0, first_index=f, prior_frame=two[0])
        assert upos[f] == o.pos_px[0], (name, f)
        assert ucnt[f] == o.nonempty[0], (name, f)
    # (5) physics: detected positions follow the synthetic front
    det = np.nonzero(pos >= 0)[0]
    assert det.size > 50
    ideal = np.array([spec.front_position(float(f)) for f in det])
    assert np.abs(pos[det] - ideal).max() < 20
    return pos, fe


def test_fullsize_c2_properties(engine):
    _fullsize_checks(engine, "C2", "half_maximum", 20000)         # BASELINE size: 1024x128 x 20000


def test_fullsize_c3_exit(engine):
    pos, fe = _fullsize_checks(engine, "C3", "threshold", 20000)  # BASELINE size: 1024x256 x 20000
    spec = syn.config_spec("C3")
    assert fe != FF_NO_EXIT and abs(fe - spec.exit_frame(10)) < 12 and abs(fe - 15000) < 40


def _torch_diff_reference(dec, prev, bg, thr):
    """Test-only integer restatement of subtract_scalar_background + frame difference
    (scripts/process_videos.py:670-674, :397-399) in torch, on decoded frames."""
    sub = torch.clamp(dec.to(torch.int32) - bg, min=0)
    prior = torch.cat([prev[None], sub[:-1]])
    d = sub - prior
    d[d < thr] = 0
    return d, sub[-1]


def test_fullsize_c4_gradient_with_retained_uint16_difference(engine):
    """BASELINE config 4 at full size (1024x1024 x 5000, gradient, diff retained): every frame of the
    10.5 GB difference image is checked against an independent torch restatement computed from
    the unpack kernel's output, and sampled frames against the oracle."""
    spec = syn.config_spec("C4")
    n, h, w, fb = spec.n_frames, spec.height, spec.width, spec.frame_bytes
    packed = syn.render_packed_torch(spec, engine.device)
    scalars, bg_dev = engine.clip_scalars(packed[:fb], h, w, 12)
    params = DetectionParams(method="gradient")
    res = engine.process_range(packed, n, h, w, 12, params, scalars, bg_dev, diff_dtype="uint16")
    bg, thr = int(scalars.background), 5
    assert res.diff.dtype == torch.uint16 and tuple(res.diff.shape) == (n, h, w)
    assert int(res.diff[0].to(torch.int32).abs().sum()) == 0          # no prior for the first frame
    prev = None
    step = 250
    for a in range(0, n, step):
        dec = engine.unpack(packed[a * fb:(a + step) * fb], step, h, w, 12)
        if prev is None:
            prev = torch.clamp(dec[0].to(torch.int32) - bg, min=0)
        want, prev = _torch_diff_reference(dec, prev, bg, thr)
        got = res.diff[a:a + step].to(torch.int32)
        if a == 0:
            want[0] = 0
        assert torch.equal(got, want), f"difference image differs in frames [{a}, {a + step})"
        del dec, want, got
    # the count-only and the difference kernels must agree on everything else
    plain = engine.process_range(packed, n, h, w, 12, params, scalars, bg_dev)
    assert torch.equal(plain.pos, res.pos) and torch.equal(plain.counts, res.counts)
    assert int(plain.first_exit.item()) == int(res.first_exit.item())
    # sampled frames against the oracle (positions, counts and the centre row of the difference)
    frame0 = fo.frames_from_bytes(packed[:fb].cpu().numpy(), 1, h, w, 12)[0]
    unt = engine.process_range(packed, n, h, w, 12, params, scalars, bg_dev, truncate=False)
    upos, ucnt = unt.pos.cpu().numpy(), unt.counts.cpu().numpy()
    t_in = int(spec.t_enter)
    for f in (1, t_in - 1, t_in + 3, t_in + 200, n - 1):
        two = fo.frames_from_bytes(packed[(f - 1) * fb:(f + 1) * fb].cpu().numpy(), 2, h, w, 12)
        o = fo.process_clip(two[1:], fo.ClipParams(method="gradient"), frame0=frame0, first_index=f,
                            prior_frame=two[0])
        assert upos[f] == o.pos_px[0] and ucnt[f] == o.nonempty[0], f
        sub = np.maximum(two.astype(np.int64) - bg, 0)
        d = sub[1] - sub[0]
        d[d < thr] = 0
        assert np.array_equal(res.diff[f].cpu().numpy().astype(np.int64), d), f
    fe = int(res.first_exit.item())
    assert fe != FF_NO_EXIT and abs(fe - spec.exit_frame(10)) < 40


def test_fullsize_unpack_roundtrip(engine):
    """Stage 1 at BASELINE size: repacking the unpack kernel's output reproduces the .mraw bytes."""
    spec = syn.config_spec("C3")
    n, h, w, fb = spec.n_frames, spec.height, spec.width, spec.frame_bytes
    packed = syn.render_packed_torch(spec, engine.device)
    step = 2500
    for a in range(0, n, step):
        src = packed[a * fb:(a + step) * fb]
        px = engine.unpack(src, step, h, w, 12).reshape(-1).to(torch.int32)
        assert int(px.max()) <= 4095
        p0, p1 = px[0::2], px[1::2]
        trip = torch.stack((p0 >> 4, ((p0 & 15) << 4) | (p1 >> 8), p1 & 255), dim=1).to(torch.uint8).reshape(-1)
        assert torch.equal(trip, src)
        del px, p0, p1, trip


def test_collection_of_pinned_recordings_matches_oracle(tmp_path, engine):
    """Config 5 in miniature: files on disk -> open_collection -> pin_memory -> process_collection
    (mixed methods / calibrations); every row against the oracle."""
    from high_speed_image_processing_b200.photron import open_collection
    from high_speed_image_processing_b200.process_videos import process_collection
    methods = ("half_maximum", "threshold", "gradient", "half_maximum")
    vdir = tmp_path / "videos"
    specs, cfgs = [], []
    for i, m in enumerate(methods):
        base = syn.config_spec("C2" if i % 2 == 0 else "C1", n_frames=220 + 17 * i, seed=900 + i)
        spec = syn.SyntheticSpec(**{**base.__dict__, "velocity": 6.0 if i != 3 else 1.0, "t_enter": 9.0 + i,
                                    "style": "mini" if m != "half_maximum" else "nova"})
        specs.append(spec)
        syn.write_clip(vdir, f"run-{i}-", spec)
        cfg = VideoSourceConfig(name=f"v{i}")
        cfg.detection_method = m
        cfg.file_calibrations = [FileCalibration(calibration=0.001 * (i + 1), position_offset=0.25 * i,
                                                 files=[f"run-{i}-"])]
        cfgs.append(cfg)
    coll = open_collection(str(vdir))
    for v in coll:
        v.pin_memory()
        assert v.frame_store.is_pinned
    results = process_collection(coll, cfgs, engine=engine)
    assert list(results) == [0, 1, 2, 3]
    for i, spec in enumerate(specs):
        frames = syn.render_frames(spec)
        want = fo.process_clip(frames, fo.ClipParams(method=methods[i]))
        res = results[i]
        assert [(r[0], r[2]) for r in res.rows] == want.records, i
        assert res.first_exit == (want.first_exit if want.first_exit < spec.n_frames else None)
        for f, t, px, pm, _ in res.rows:
            assert pm == fo.position_m(px, 0.001 * (i + 1), 0.25 * i)
            assert t == fo.frame_time_absolute(f, spec.start_frame, spec.skip_frame, spec.record_rate)
    coll.close_all()


def test_fullsize_c1(engine):
    spec = syn.config_spec("C1")
    frames = syn.render_frames(spec)
    packed = torch.from_numpy(syn.pack_frames(frames, 12)).to(engine.device)
    scalars, bg_dev = engine.clip_scalars(packed[:spec.frame_bytes], spec.height, spec.width, 12)
    res = engine.process_range(packed, spec.n_frames, spec.height, spec.width, 12,
                               DetectionParams(method="half_maximum"), scalars, bg_dev)
    want = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    exp = want.pos_px.copy()
    exp[want.first_exit:] = FF_POS_DROPPED
    assert np.array_equal(res.pos.cpu().numpy(), exp)
    assert int(res.first_exit.cpu().item()) == want.first_exit


@pytest.mark.parametrize("bits,header", [(16, "cihx"), (8, "cih"), (16, "cih")])
def test_file_level_driver_on_16bit_and_8bit_recordings(tmp_path, bits, header):
    """Decode row a1 end to end for the other storage depths: .cihx / legacy .cih + .mraw on disk ->
    process_video_source (memory-mapped streaming) -> rows identical to the oracle's."""
    spec = syn.SyntheticSpec(width=256, height=48, n_frames=140, bits=bits, style="mini", t_enter=10.0,
                             velocity=3.0, seed=160 + bits)
    frames = syn.render_frames(spec)
    vdir = tmp_path / "Mini-Video-Files"
    syn.write_clip(vdir, "run-2-", spec, frames=frames, header=header)
    cfg = VideoSourceConfig(name="Mini")
    cfg.enabled = True
    cfg.detection_method = "threshold"
    cfg.video_path = str(vdir)
    cfg.output_dir = str(tmp_path / "out")
    cfg.file_calibrations = [FileCalibration(calibration=0.000869565, position_offset=0.050237,
                                             files=["run-1-:run-10-"])]
    want = fo.process_clip(frames, fo.ClipParams(method="threshold"))
    if header == "cihx":
        res = process_video_source(cfg, None, verbose=False)["run-2-.cihx"]
    else:                                            # process_video_source globs *.cihx like the reference (:1300)
        with open_video(str(vdir / "run-2-.cih")) as video:
            assert video.storage_bits == bits and video[3].dtype == (np.uint8 if bits == 8 else np.uint16)
            assert np.array_equal(video[3], frames[3])
            res = process_video(video, cfg, 0.000869565, 0.050237)
    assert [(r[0], r[2]) for r in res.rows] == want.records
    assert res.first_exit == (want.first_exit if want.first_exit < len(frames) else None)
    for f, t, px, pm, _ in res.rows:
        assert pm == fo.position_m(px, 0.000869565, 0.050237)
