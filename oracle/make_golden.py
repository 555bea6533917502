"""Generate tests/golden/* by EXECUTING THE REFERENCE in this container.

TEST INFRASTRUCTURE ONLY.  Run here (``python oracle/make_golden.py``); the GPU box has no
``/root/reference``, so the vectors are committed under ``tests/golden/``.

The reference imports two packages that are not installed (pyMRAW, matplotlib).  Both are
stubbed: ``pyMRAW.load_video`` is served by the oracle's restated decoder (so the decode layout
itself stays *unpinned*, see oracle/flame_oracle.py), matplotlib by an empty module (no plot
is ever drawn).  Everything else below is the reference's own code:

  * PhotonVideo / TimingInfo / SpatialCalibration / VideoCollection / MPIVideoProcessor
    (src/photron/*.py)
  * subtract_scalar_background, is_empty_frame, subtract_prior_frame, FlameDetector.detect,
    FileCalibration, VideoSourceConfig (scripts/process_videos.py)

While generating, the script asserts that oracle/flame_oracle.py reproduces each pinned
quantity bit-for-bit.
"""
from __future__ import annotations

import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLD = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))

from oracle import flame_oracle as fo  # noqa: E402
from high_speed_image_processing_b200 import mraw as our_mraw  # noqa: E402  (header parsing only)
from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402


def _install_stubs() -> None:
    pm = types.ModuleType("pyMRAW")

    def load_video(path):
        info = our_mraw.get_cih(path)
        raw = np.fromfile(str(Path(path).with_suffix(".mraw")), dtype=np.uint8)
        images = fo.frames_from_bytes(raw, int(info["Total Frame"]), int(info["Image Height"]),
                                      int(info["Image Width"]), int(info["Color Bit"]))
        return images, info

    pm.load_video = load_video
    sys.modules["pyMRAW"] = pm
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def _jsonable(o):
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    if isinstance(o, np.bool_):
        return bool(o)
    raise TypeError(f"not JSON serialisable: {type(o).__name__}")


def _sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


class FakeComm:
    def __init__(self, rank, size):
        self._r, self._s = rank, size

    def Get_rank(self):
        return self._r

    def Get_size(self):
        return self._s


def replay_head_loop(pv, video, cal, off, exit_margin=15):
    """The reference loop scripts/process_videos.py:1441-1516 with plotting removed, calling the
    reference's own functions.  Returns per-frame detector outputs and the result tuples."""
    background_scalar = float(np.max(video[0]))                          # :1357-1358
    cfg = pv.FlameDetectorConfig(gaussian_sigma=1.5, morphology_kernel_size=3, max_velocity_change_m_s=200.0)
    det = pv.FlameDetector(config=cfg, frame_rate=video.frame_rate, calibration_m_per_px=cal)
    per_frame, results, empty = [], [], 0
    stop_reason = None
    for frame_idx in range(len(video)):
        frame = video[frame_idx]
        time_s = video.get_absolute_time(frame_idx)
        sub = pv.subtract_scalar_background(frame, background_scalar)
        noise_thresh = max(10.0, background_scalar * 0.5)
        if pv.is_empty_frame(sub, noise_threshold=noise_thresh, min_signal_fraction=0.0005):
            empty += 1
            det._prior_frame = sub.copy()
            continue
        r = det.detect(frame=frame, frame_idx=frame_idx, background_scalar=background_scalar)
        per_frame.append({
            "frame": frame_idx,
            "final": r.final_position, "min_gradient": r.pos_min_gradient,
            "rightmost_sobel": r.pos_rightmost_sobel, "search": list(r.search_bounds),
            "diff_sha1": None if r.frame_diff is None else _sha(r.frame_diff),
        })
        pos = r.final_position
        velocity = det.last_velocity
        if pos is not None and pos >= video.width - exit_margin:
            det.clear_last_central_difference()
            stop_reason = ("exit", frame_idx)
            break
        vh = det.get_velocity_history()
        if velocity is not None and len(vh) >= 2:
            prev_v1 = vh[-2][1]
            if prev_v1 is not None and prev_v1 > 100 and (prev_v1 - velocity) / prev_v1 > 0.5:
                det.clear_last_central_difference()
                stop_reason = ("velocity_drop", frame_idx)
                break
        if pos is not None:
            post = det.ddt_detected and frame_idx >= det.ddt_frame
            results.append([frame_idx, time_s, int(pos), pos * cal + off, bool(post)])
    vel = [[e[0], e[1], e[2], e[3]] for e in det.get_velocity_history()]
    return {"background": background_scalar, "empty": empty, "per_frame": per_frame, "results": results,
            "velocity_history": vel, "ddt_frame": det.ddt_frame, "stop": stop_reason}


def detector_api_golden(pv, ho) -> dict:
    """The reference's frame-level API driven directly (no driver loop, no empty-frame skip):
    FlameDetector.detect with every intermediate image hashed, two configurations, on frames
    committed as tests/golden/detector_frames.npz; plus the element-wise frame functions."""
    spec = syn.SyntheticSpec(width=136, height=40, n_frames=20, bits=16, style="nova", t_enter=2.0,
                             velocity=7.0, tail_length=30.0, curvature_px=3.0, seed=4242, record_rate=100000)
    frames = syn.render_frames(spec)
    np.savez_compressed(GOLD / "detector_frames.npz", frames=frames)
    images = ("frame_subtracted", "frame_diff", "noise_removed", "blurred", "sobel_output", "gradient_output")
    out = {"frames_sha1": _sha(frames), "shape": list(frames.shape), "frame_rate": spec.record_rate, "runs": []}
    variants = [
        {"name": "defaults", "cfg": {}, "calibration": 0.000833333, "order": list(range(20)), "background": None},
        {"name": "k5_sigma1_thr3.5_gaps", "calibration": 0.002, "background": 37,
         "cfg": {"morphology_kernel_size": 5, "gaussian_sigma": 1.0, "frame_diff_threshold": 3.5,
                 "min_gradient_strength": 4.0, "edge_margin_px": 3, "search_window_px": 12,
                 "sobel_threshold_fraction": 0.25, "use_spline_estimator": False},
         "order": [0, 2, 3, 5, 6, 7, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19]},
        {"name": "k1_sigma2", "calibration": 0.000833333, "background": None,
         "cfg": {"morphology_kernel_size": 1, "gaussian_sigma": 2.0}, "order": list(range(1, 12))},
    ]
    for var in variants:
        cfg = pv.FlameDetectorConfig(**var["cfg"])
        det = pv.FlameDetector(config=cfg, frame_rate=spec.record_rate, calibration_m_per_px=var["calibration"])
        bg = float(np.max(frames[0])) if var["background"] is None else float(var["background"])
        hcfg = ho.HeadConfig(**{k: v for k, v in var["cfg"].items() if k in ho.HeadConfig.__dataclass_fields__})
        orc = ho.FrameDetectorOracle(spec.record_rate, var["calibration"], hcfg,
                                     kernel_size=var["cfg"].get("morphology_kernel_size", 3))
        calls = []
        for idx in var["order"]:
            r = det.detect(frame=frames[idx], frame_idx=idx, background_scalar=bg)
            o = orc.detect(frames[idx], idx, bg)
            for name in images:
                a, b = getattr(r, name), o["images"][name]
                assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), (var["name"], idx, name)
            assert (r.final_position, r.pos_min_gradient, r.pos_rightmost_sobel, tuple(r.search_bounds)) == \
                (o["final"], o["min_gradient"], o["rightmost_sobel"], tuple(o["search"])), (var["name"], idx)
            calls.append({"frame": idx, "time_s": r.time_s, "final": r.final_position,
                          "min_gradient": r.pos_min_gradient, "rightmost_sobel": r.pos_rightmost_sobel,
                          "spline": r.pos_spline_predicted, "search": list(r.search_bounds),
                          "sha1": {n: (None if getattr(r, n) is None else _sha(getattr(r, n))) for n in images}})
        vel = [list(e) for e in det.get_velocity_history()]
        assert vel == orc.velocities and det.ddt_frame == orc.ddt_frame
        out["runs"].append({"name": var["name"], "cfg": var["cfg"], "calibration": var["calibration"],
                            "background": bg, "order": var["order"], "calls": calls, "velocity_history": vel,
                            "ddt_frame": det.ddt_frame, "last_position": det.last_position,
                            "last_velocities": list(det.last_velocities),
                            "pre_ddt": len(det.get_pre_ddt_velocities()), "post_ddt": len(det.get_post_ddt_velocities()),
                            "max_displacement_px": det._max_displacement_px,
                            "prior_sha1": _sha(det._prior_frame)})
    # ---- element-wise frame functions ------------------------------------------------------------
    f5, f6, f7 = frames[5], frames[6], frames[7]
    sub6 = pv.subtract_scalar_background(f6, 41.5)
    ops = {
        "sub_bg_41.5": _sha(sub6),
        "sub_bg_max0": _sha(pv.subtract_scalar_background(f6, float(np.max(frames[0])))),
        "prior_raw_thr0": _sha(pv.subtract_prior_frame(f6, f5, threshold=0.0)),
        "prior_raw_thr7.5": _sha(pv.subtract_prior_frame(f6, f5, threshold=7.5)),
        "prior_f64": _sha(pv.subtract_prior_frame(sub6, pv.subtract_scalar_background(f5, 41.5), threshold=5.0)),
        "three_thr0": _sha(pv.three_frame_difference(f5, f6, f7)),
        "three_thr3": _sha(pv.three_frame_difference(f5, f6, f7, threshold=3.0)),
        "empty": {f"{thr}/{frac}": bool(pv.is_empty_frame(sub6, noise_threshold=thr, min_signal_fraction=frac))
                  for thr in (10.0, 50.0, 20.5, 2000.0) for frac in (0.001, 0.0005, 0.05, 0.5)},
        "empty_raw_defaults": bool(pv.is_empty_frame(f6)),
        "count_above": {str(thr): int(np.sum(sub6 > thr)) for thr in (10.0, 50.0, 20.5, 2000.0)},
    }
    assert ops["sub_bg_41.5"] == _sha(fo.subtract_scalar_background(f6, 41.5))
    assert ops["prior_raw_thr7.5"] == _sha(fo.frame_difference(f6, f5, 7.5))
    assert ops["three_thr3"] == _sha(fo.three_frame_difference(f5, f6, f7, 3.0))
    work = GOLD / "_work"
    work.mkdir(parents=True, exist_ok=True)
    table = {"Frame": [39, 40], "Time_s": ["0.003368750", "0.003375000"], "Position_px": [6, 14],
             "Note": ["a b", 'q"x']}
    pv.write_results(table, str(work / "wr.txt"))
    ops["write_results"] = {"columns": [[k, v] for k, v in table.items()],   # JSON objects are key-sorted
                             "bytes": (work / "wr.txt").read_bytes().decode("latin-1")}
    out["ops"] = ops
    out["config_defaults"] = {k: getattr(pv.FlameDetectorConfig(), k) for k in pv.FlameDetectorConfig.__dataclass_fields__}
    out["result_fields"] = list(pv.FlameDetectionResult.__dataclass_fields__)
    return out


def main() -> None:
    _install_stubs()
    sys.path.insert(0, str(REF / "scripts"))
    sys.path.insert(0, str(REF))
    import process_videos as pv            # the reference script
    from src import photron as rp          # the reference package

    GOLD.mkdir(parents=True, exist_ok=True)
    work = GOLD / "_work"
    gold: dict = {"generated_by": "oracle/make_golden.py", "reference": "Nadexterbrown/High-Speed-Image-Processing"}

    # ---- a small synthetic recording written to disk ---------------------------------------
    spec = syn.SyntheticSpec(width=128, height=16, n_frames=72, bits=12, style="nova", t_enter=8.0,
                             velocity=2.0, tail_length=40.0, curvature_px=2.0, seed=77,
                             record_rate=160000, start_frame=500)
    frames = syn.render_frames(spec)
    cihx = syn.write_clip(work, "run-3-_C001H001S0001", spec, frames=frames)
    packed = np.fromfile(str(cihx.with_suffix(".mraw")), dtype=np.uint8)
    np.savez_compressed(GOLD / "clip_small.npz", packed=packed,
                        cihx=np.frombuffer(cihx.read_bytes(), dtype=np.uint8))
    gold["clip_small"] = {"width": spec.width, "height": spec.height, "n_frames": spec.n_frames, "bits": 12,
                          "record_rate": spec.record_rate, "start_frame": spec.start_frame,
                          "stem": "run-3-_C001H001S0001"}

    # ---- PhotonVideo through the reference ------------------------------------------------
    video = rp.open_video(str(cihx), calibration=rp.SpatialCalibration(scale=0.000833333, units="m"))
    assert len(video) == spec.n_frames and video.frame_shape == (spec.height, spec.width)
    assert np.array_equal(video[5], frames[5])
    gold["video"] = {
        "len": len(video), "frame_shape": list(video.frame_shape), "frame_rate": video.frame_rate,
        "dtype": str(video.dtype), "duration": video.duration, "has_absolute_timing": video.has_absolute_timing,
        "absolute_time": [video.get_absolute_time(i) for i in range(len(video))],
        "time_trigger10": [rp.TimingInfo(frame_rate=160000, trigger_frame=10).frame_to_time(i) for i in range(20)],
        "cihx_metadata": {k: (str(v) if k == "recording_datetime" else v) for k, v in video.cihx_metadata.items()},
        "datetime_5": str(video.get_datetime(5)),
        "time_to_frame": [video.timing.time_to_frame(t) for t in (0.0, 1e-4, 3.3e-4)],
    }
    for i, t in enumerate(gold["video"]["absolute_time"]):
        assert t == fo.frame_time_absolute(i, spec.start_frame, spec.skip_frame, spec.record_rate)

    # ---- per-frame primitives through the reference ---------------------------------------------
    f0 = video[0]
    bg = float(np.max(f0))                                                   # :1357-1358
    row = f0.shape[0] // 2
    line = f0[row, :].astype(np.float64)                                     # :1361-1362
    c_mean, c_std, c_max = np.mean(line), np.std(line), np.max(line)         # :1363-1365
    c_thr = max(c_mean + 5 * c_std, c_max * 2.0)                             # :1367-1370
    assert fo.background_scalar(f0) == bg
    assert fo.centerline_stats(f0) == (float(c_mean), float(c_std), float(c_max), float(c_thr))
    noise_thr = max(10.0, bg * 0.5)                                          # :1458
    prim = {"background": bg, "centerline_mean": float(c_mean), "centerline_std": float(c_std),
            "centerline_max": float(c_max), "flame_threshold": float(c_thr), "noise_threshold": noise_thr,
            "sub_sha1": [], "empty": [], "nonempty_count": [], "diff_sha1": []}
    prior = None
    profiles = np.zeros((spec.n_frames, spec.width), dtype=np.float64)
    for i in range(spec.n_frames):
        sub = pv.subtract_scalar_background(video[i], bg)                    # :670-674
        assert np.array_equal(sub, fo.subtract_scalar_background(frames[i], bg))
        is_empty = bool(pv.is_empty_frame(sub, noise_threshold=noise_thr, min_signal_fraction=0.0005))
        assert is_empty == fo.is_empty_frame(sub, noise_thr, 0.0005)
        prim["sub_sha1"].append(_sha(sub))
        prim["empty"].append(is_empty)
        prim["nonempty_count"].append(int(np.sum(sub > noise_thr)))
        if prior is not None:
            d = pv.subtract_prior_frame(sub, prior, threshold=5.0)           # :677-701 (== :397-399)
            assert np.array_equal(d, fo.frame_difference(sub, prior, 5.0))
            prim["diff_sha1"].append(_sha(d))
            profiles[i] = d[row, :]
        else:
            prim["diff_sha1"].append(None)
        prior = sub
    gold["primitives"] = prim
    np.savez_compressed(GOLD / "clip_small_profiles.npz", diff_profiles=profiles)

    # the oracle loop must agree with those primitives
    oc = fo.process_clip(frames, fo.ClipParams(method="gradient", keep_profiles=True))
    assert oc.background == bg and oc.flame_threshold == float(c_thr)
    assert list(oc.empty) == prim["empty"] and list(oc.nonempty) == prim["nonempty_count"]
    assert np.array_equal(oc.profiles[1:], profiles[1:])

    # ---- gradient method == HEAD Method A primitives on the raw profile (:413,427-430) -----------
    grad = []
    for i in range(1, spec.n_frames):
        g = np.gradient(profiles[i])
        mv = np.min(g)
        ref_pos = int(np.argmin(g)) if mv < -10.0 else None
        assert ref_pos == fo.detect_gradient(profiles[i], 10.0)
        grad.append(ref_pos)
    gold["gradient_on_profiles"] = [None] + grad

    # ---- HEAD detector replay (f1/f2 rows of SURVEY 8f) ----------------------------------------
    gold["head_replay"] = replay_head_loop(pv, video, 0.000833333, 1.347567)

    # ---- the frame-level API (seam B3) driven directly -------------------------------------------
    from oracle import head_oracle as ho
    gold["detector_api"] = detector_api_golden(pv, ho)

    # ---- the reference's own driver, end to end (f1 + f2 + f3) -----------------------------------
    # process_video_source (:1277-1629) with only the matplotlib renderers replaced by no-ops;
    # the text files it writes are the golden outputs of the HEAD path.
    pv.save_frame_image = lambda *a, **k: None
    pv.generate_stacked_sequence = lambda *a, **k: None
    pv.generate_stacked_sequence_single_column = lambda *a, **k: None
    ddir = work / "driver" / "Nova-Video-Files"
    syn.write_clip(ddir, "run-3-", spec, frames=frames)                       # matches "run-3-:run-10-"
    syn.write_clip(ddir, "run-5-_C001H001S0001", spec, frames=frames)         # the last-integer bug: default cal
    dcfg = pv.VideoSourceConfig(name="Nova")
    dcfg.enabled = True
    dcfg.calibration = 0.001
    dcfg.position_offset = 0.25
    dcfg.video_path = str(ddir)
    dcfg.output_dir = str(work / "driver" / "out")
    dcfg.file_calibrations = [pv.FileCalibration(calibration=0.000833333, position_offset=1.347567,
                                                 files=["run-3-:run-10-"])]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        pv.process_video_source(dcfg, None)
    outs = {}
    for path in sorted((work / "driver" / "out").glob("*.txt")):
        outs[path.name] = path.read_text()
    assert "run-3--flame-position.txt" in outs and "run-5-_C001H001S0001-flame-position.txt" in outs
    gold["driver_outputs"] = outs
    gold["driver_config"] = {"calibration": 0.001, "position_offset": 0.25}

    # ---- ... and with skip_frames (:1443-1445: skipped frames are not processed and do not become the
    # prior frame), on the small recording ---------------------------------------------------------
    sdir = work / "driver_skip" / "Nova-Video-Files"
    syn.write_clip(sdir, "run-3-", spec, frames=frames)
    scfg = pv.VideoSourceConfig(name="Nova")
    scfg.enabled = True
    scfg.calibration = 0.000833333
    scfg.position_offset = 1.347567
    scfg.skip_frames = [21, 22, 30, 41, 42, 43]
    scfg.video_path = str(sdir)
    scfg.output_dir = str(work / "driver_skip" / "out")
    with contextlib.redirect_stdout(io.StringIO()):
        pv.process_video_source(scfg, None)
    gold["driver_outputs_skip"] = {"skip_frames": scfg.skip_frames,
                                   "outputs": {q.name: q.read_text()
                                               for q in sorted((work / "driver_skip" / "out").glob("*.txt"))}}
    assert gold["driver_outputs_skip"]["outputs"] and \
        gold["driver_outputs_skip"]["outputs"]["run-3--flame-position.txt"] != outs["run-3--flame-position.txt"]

    # ---- ... with trigger-relative time (use_absolute_time = False, trigger_frame set: :1449-1452) --------
    tdir = work / "driver_trigger" / "Nova-Video-Files"
    syn.write_clip(tdir, "run-3-", spec, frames=frames)
    tcfg = pv.VideoSourceConfig(name="Nova")
    tcfg.enabled = True
    tcfg.calibration = 0.000833333
    tcfg.position_offset = 1.347567
    tcfg.use_absolute_time = False
    tcfg.trigger_frame = 12
    tcfg.video_path = str(tdir)
    tcfg.output_dir = str(work / "driver_trigger" / "out")
    with contextlib.redirect_stdout(io.StringIO()):
        pv.process_video_source(tcfg, None)
    gold["driver_outputs_trigger"] = {"trigger_frame": 12, "outputs": {
        q.name: q.read_text() for q in sorted((work / "driver_trigger" / "out").glob("*.txt"))}}
    assert gold["driver_outputs_trigger"]["outputs"]["run-3--flame-position.txt"] != outs["run-3--flame-position.txt"]

    # ---- ... on recordings with a DDT event (velocity jump > 1250 m/s: pre-/post-DDT files, :506-516) and
    # with a front that slows to under half its speed (the velocity-drop stop, :1499-1509).  A recording
    # is two constant-velocity pieces of the synthetic generator spliced at a frame; the test rebuilds it.
    def spliced(kind, va, vb, splice, n_frames_e):
        base = dict(width=256, height=16, bits=12, style="mini", curvature_px=1.0, seed=700, record_rate=160000,
                    start_frame=0, n_frames=n_frames_e)
        sa = syn.SyntheticSpec(t_enter=5.0, velocity=va, **base)
        xb = sa.front_position(float(splice))
        sb = syn.SyntheticSpec(t_enter=splice - xb / vb, velocity=vb, **base)
        fr = np.concatenate([syn.render_frames(sa, 0, splice), syn.render_frames(sb, splice, n_frames_e)])
        edir = work / f"event_{kind}" / "Nova-Video-Files"
        syn.write_clip(edir, "run-3-", sa, frames=fr)
        ec = pv.VideoSourceConfig(name="Nova")
        ec.enabled = True
        ec.calibration = 0.000833333
        ec.position_offset = 1.347567
        ec.video_path = str(edir)
        ec.output_dir = str(work / f"event_{kind}" / "out")
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            pv.process_video_source(ec, None)
        files = {q.name: q.read_text() for q in sorted((work / f"event_{kind}" / "out").glob("*.txt"))}
        spec_keys = ("width", "height", "n_frames", "bits", "style", "t_enter", "velocity", "curvature_px", "seed",
                     "record_rate", "start_frame")
        return {"spec_a": {k: getattr(sa, k) for k in spec_keys}, "spec_b": {k: getattr(sb, k) for k in spec_keys},
                "splice": splice, "frames_sha1": _sha(fr), "outputs": files,
                "log_mentions": [w for w in ("DDT", "velocity", "exited") if w.lower() in log.getvalue().lower()]}
    events = {"ddt": spliced("ddt", 2.0, 20.0, 40, 70), "velocity_drop": spliced("velocity_drop", 10.0, 1.0, 20, 70)}
    assert any("post-DDT" in k for k in events["ddt"]["outputs"]), list(events["ddt"]["outputs"])
    vd_rows = [ln for ln in events["velocity_drop"]["outputs"]["run-3--flame-position.txt"].splitlines()
               if ln and not ln.startswith("#")]
    assert int(vd_rows[-1].split()[0]) < 30, "the velocity-drop recording must stop at the deceleration"
    gold["driver_events"] = events

    # ---- the reference's driver on the other storage depths (16-bit little-endian, 8-bit) -----------
    # The recordings are regenerated by the test from the same SyntheticSpec (sha1 of the frames kept
    # here), so only the reference's output files are committed.
    depth_runs = {}
    for bits_d in (16, 8):
        dspec = syn.SyntheticSpec(width=192, height=24, n_frames=90, bits=bits_d, style="mini", t_enter=6.0,
                                  velocity=2.5, curvature_px=2.0, seed=500 + bits_d, record_rate=100000,
                                  start_frame=0)
        dframes = syn.render_frames(dspec)
        ddir = work / f"depth{bits_d}" / "Mini-Video-Files"
        syn.write_clip(ddir, f"run-2-_{bits_d}bit", dspec, frames=dframes)
        dc = pv.VideoSourceConfig(name="Mini")
        dc.enabled = True
        dc.calibration = 0.000869565
        dc.position_offset = 0.050237
        dc.video_path = str(ddir)
        dc.output_dir = str(work / f"depth{bits_d}" / "out")
        with contextlib.redirect_stdout(io.StringIO()):
            pv.process_video_source(dc, None)
        files = {q.name: q.read_text() for q in sorted((work / f"depth{bits_d}" / "out").glob("*.txt"))}
        assert files, f"{bits_d}-bit: the reference wrote no result file"
        depth_runs[str(bits_d)] = {"spec": {k: getattr(dspec, k) for k in ("width", "height", "n_frames", "bits", "style",
                                                                          "t_enter", "velocity", "curvature_px", "seed",
                                                                          "record_rate", "start_frame")},
                                   "frames_sha1": _sha(dframes), "stem": f"run-2-_{bits_d}bit",
                                   "calibration": 0.000869565, "position_offset": 0.050237, "outputs": files}
    gold["driver_other_depths"] = depth_runs

    # ---- FileCalibration / VideoSourceConfig ----------------------------------------------------
    rules = [
        pv.FileCalibration(calibration=0.000833333, position_offset=1.0159, files=["run-1-"]),
        pv.FileCalibration(calibration=0.000833333, position_offset=1.197565, files=["run-2-"]),
        pv.FileCalibration(calibration=0.000833333, position_offset=1.347567, files=["run-3-:run-10-"]),
    ]
    cfg = pv.VideoSourceConfig(name="Nova", calibration=1.0, position_offset=0.0, file_calibrations=rules)
    names = ["run-1-.cihx", "run-2-_C001H001S0001.cihx", "run-3-.cihx", "run-5-_C001H001S0001.cihx", "run-7-.cihx",
             "run-10-.cihx", "run-11-.cihx", "Run-001.cihx", "other.cihx", "run-12-x3.cihx", "noint.cihx"]
    gold["calibration_lookup"] = {n: list(cfg.get_calibration_for_file(n)) for n in names}
    fc = pv.FileCalibration(calibration=1.0, files=["Run-001:Run-005", "special", "A:B"])
    gold["file_calibration_matches"] = {n: fc.matches(n) for n in
                                        ["Run-003.cihx", "Run-006.cihx", "xx-special-yy.cihx", "Run-000.cihx",
                                         "Run-005_v2.cihx", "abc.cihx", "7.cihx"]}
    cfg2 = pv.VideoSourceConfig(name="x")
    cfg2.video_path = "/abs/path"
    gold["abs_path_kept"] = cfg2.video_path

    # ---- distribute_indices -----------------------------------------------------------------------
    dist = {}
    for total in (0, 1, 7, 10, 23):
        for size in (1, 2, 3, 4, 8):
            for strategy in ("round_robin", "contiguous"):
                dist[f"{total}/{size}/{strategy}"] = [
                    rp.MPIVideoProcessor(FakeComm(r, size)).distribute_indices(total, strategy) for r in range(size)]
    gold["distribute_indices"] = dist
    serial = rp.MPIVideoProcessor(None)
    gold["serial_processor"] = {"rank": serial.rank, "size": serial.size, "is_root": serial.is_root,
                                "is_parallel": serial.is_parallel, "gather": serial.gather([1, 2]),
                                "indices": serial.distribute_indices(5)}

    # ---- VideoCollection global index ----------------------------------------------------------------
    spec_b = syn.SyntheticSpec(width=64, height=8, n_frames=5, bits=16, seed=5)
    spec_c = syn.SyntheticSpec(width=64, height=8, n_frames=9, bits=8, seed=6)
    cdir = work / "coll"
    syn.write_clip(cdir, "a_first", spec_b)
    syn.write_clip(cdir, "b_second", spec_c)
    coll = rp.open_collection(str(cdir))
    gold["collection"] = {
        "len": len(coll), "total_frames": coll.total_frames,
        "resolve": {str(g): list(coll.global_to_local(g)) for g in (0, 4, 5, 13, -1, -14)},
        "local_to_global": coll.local_to_global(1, 3),
        "names": [p.name for p in coll.filepaths],
        "frame_4_sha1": _sha(coll.get_global_frame(4)), "frame_5_sha1": _sha(coll.get_global_frame(5)),
        "dtypes": [str(v.dtype) for v in coll],
    }

    # ---- README sample rows (README.md:93-96) --------------------------------------------------------
    gold["readme_rows"] = [
        {"frame": 39, "time_s": "0.003368750", "px": 6, "pos_m": "1.352566998"},
        {"frame": 40, "time_s": "0.003375000", "px": 14, "pos_m": "1.359233662"},
    ]
    t39 = rp.TimingInfo(frame_rate=160000, start_frame=500, skip_frame=1).frame_to_absolute_time(39)
    assert f"{t39:.9f}" == "0.003368750" and f"{6 * 0.000833333 + 1.347567:.9f}" == "1.352566998"

    (GOLD / "reference_golden.json").write_text(json.dumps(gold, indent=1, sort_keys=True, default=_jsonable))
    import shutil
    shutil.rmtree(work)
    print("wrote", GOLD / "reference_golden.json", (GOLD / "reference_golden.json").stat().st_size, "bytes")
    hr = gold["head_replay"]
    print("HEAD replay:", len(hr["per_frame"]), "detect calls,", len(hr["results"]), "rows, stop =", hr["stop"])


if __name__ == "__main__":
    main()
