"""CUDA kernels against the oracle, through the C-ABI (FlameFrontEngine -> libflamefront.so).
Bit-exact comparisons: everything on this path is integer work."""
import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED
from high_speed_image_processing_b200.engine import ClipScalars, DetectionParams
from oracle import flame_oracle as fo

pytestmark = pytest.mark.gpu


def dev(a, engine):
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


def oracle_pos(res):
    """Oracle positions with the exit truncation applied the way ff_truncate does."""
    want = res.pos_px.copy()
    want[res.first_exit:] = FF_POS_DROPPED
    return want


# ---------------------------------------------------------------------------------- stage 1
@pytest.mark.parametrize("n,h,w", [
    (3, 64, 512),     # P=32768: K=1 tiles, 16 tiles/frame
    (2, 128, 1024),   # P=131072: K=4 tiles, 16 tiles/frame
    (5, 8, 100),      # P=800 -> P%32==0, single ragged tile
    (4, 36, 1000),    # P=36000: ragged last tile (K=1)
    (2, 130, 1024),   # K=4 with a ragged last tile
    (3, 5, 6),        # P=30: generic kernel
    (1, 1, 2),        # minimum
])
def test_unpack12_matches_oracle(engine, n, h, w):
    rng = np.random.default_rng(n * 1000 + h + w)
    packed = rng.integers(0, 256, size=n * h * w * 3 // 2, dtype=np.uint8)
    got = engine.unpack(dev(packed, engine), n, h, w, 12).cpu().numpy()
    assert got.dtype == np.uint16
    assert np.array_equal(got, fo.frames_from_bytes(packed, n, h, w, 12))


def test_unpack_passthrough_and_errors(engine):
    rng = np.random.default_rng(5)
    raw16 = rng.integers(0, 65536, size=(3, 8, 64), dtype=np.uint16)
    got = engine.unpack(dev(raw16.view(np.uint8).reshape(-1), engine), 3, 8, 64, 16).cpu().numpy()
    assert np.array_equal(got, raw16)
    raw8 = rng.integers(0, 256, size=(3, 8, 64), dtype=np.uint8)
    assert np.array_equal(engine.unpack(dev(raw8.reshape(-1), engine), 3, 8, 64, 8).cpu().numpy(), raw8)
    with pytest.raises(ValueError):
        engine.unpack(dev(np.zeros(10, np.uint8), engine), 1, 3, 3, 12)       # odd pixel count
    with pytest.raises(ValueError):
        engine.unpack(dev(np.zeros(10, np.uint8), engine), 2, 8, 64, 12)      # buffer too small
    with pytest.raises(ValueError):
        engine.unpack(torch.zeros(64, dtype=torch.uint8), 1, 4, 8, 12)        # not on the device


# --------------------------------------------------------------------------------- stage 2a
@pytest.mark.parametrize("bits", [8, 12, 16])
@pytest.mark.parametrize("h,w", [(16, 128), (7, 50), (1, 64), (128, 1024)])
def test_background_and_scalars(engine, bits, h, w):
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=1, bits=bits, seed=h * w + bits)
    frame = syn.render_frames(spec)
    if bits == 16:
        frame = (frame.astype(np.uint32) * 13 % 65536).astype(np.uint16)      # use the full 16-bit range
    packed = syn.pack_frames(frame, bits)
    scalars, bg_dev = engine.clip_scalars(dev(packed, engine), h, w, bits)
    assert scalars.background == fo.background_scalar(frame[0])
    assert (scalars.centerline_mean, scalars.centerline_std, scalars.centerline_max,
            scalars.flame_threshold) == fo.centerline_stats(frame[0])
    assert scalars.noise_threshold == fo.empty_noise_threshold(scalars.background)
    assert int(bg_dev.cpu().item()) == int(frame.max())


# ------------------------------------------------------------------------------ stages 2b-4
def run_range(engine, frames, bits, params, **kw):
    n, h, w = frames.shape
    packed = dev(syn.pack_frames(frames, bits), engine)
    frame0 = kw.pop("frame0", frames[0])
    scalars, bg_dev = engine.clip_scalars(dev(syn.pack_frames(frame0[None], bits), engine), h, w, bits)
    return engine.process_range(packed, n, h, w, bits, params, scalars, bg_dev, **kw), scalars


def small_clip(bits=12, style="nova", w=128, h=16, n=72, seed=77, **kw):
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=n, bits=bits, style=style, t_enter=8.0, velocity=2.0,
                             tail_length=40.0, curvature_px=2.0, seed=seed, **kw)
    return syn.render_frames(spec)


@pytest.mark.parametrize("method", ["threshold", "gradient", "half_maximum"])
@pytest.mark.parametrize("use_diff", [True, False])
@pytest.mark.parametrize("bits,style", [(12, "nova"), (12, "mini"), (16, "mini"), (8, "nova")])
def test_methods_match_oracle(engine, method, use_diff, bits, style):
    frames = small_clip(bits=bits, style=style)
    params = DetectionParams(method=method, use_frame_diff=use_diff)
    res, scalars = run_range(engine, frames, bits, params, keep_profiles=True)
    want = fo.process_clip(frames, fo.ClipParams(method=method, use_frame_diff=use_diff, keep_profiles=True))
    assert scalars.background == want.background and scalars.flame_threshold == want.flame_threshold
    assert np.array_equal(res.counts.cpu().numpy(), want.nonempty.astype(np.int32))
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    fe = int(res.first_exit.cpu().item())
    assert fe == (want.first_exit if want.first_exit < len(frames) else FF_NO_EXIT)
    # profiles (int32) equal the reference's float64 profiles exactly
    prof = res.profiles.cpu().numpy()
    for i in range(len(frames)):
        if use_diff and i == 0:
            continue
        assert np.array_equal(prof[i].astype(np.float64), want.profiles[i]), i


def test_gradient_matches_reference_golden(engine, clip_small, golden):
    """`gradient` on the committed clip equals what the reference's own primitives
    (np.gradient / np.argmin on the reference's frame-difference profile) produced."""
    c = golden["clip_small"]
    n, h, w = c["n_frames"], c["height"], c["width"]
    packed = dev(clip_small["packed"], engine)
    scalars, bg_dev = engine.clip_scalars(packed[: h * w * 3 // 2], h, w, 12)
    g = golden["primitives"]
    assert (scalars.background, scalars.flame_threshold, scalars.noise_threshold) == (
        g["background"], g["flame_threshold"], g["noise_threshold"])
    res = engine.process_range(packed, n, h, w, 12, DetectionParams(method="gradient", exit_margin_px=0),
                               scalars, bg_dev, truncate=False)
    assert res.counts.cpu().numpy().tolist() == g["nonempty_count"]
    pos = res.pos.cpu().numpy()
    for i, want in enumerate(golden["gradient_on_profiles"]):
        if g["empty"][i]:
            assert pos[i] == -1
        else:
            assert pos[i] == (-1 if want is None else want), i


@pytest.mark.parametrize("dtype,np_dtype", [("uint16", np.uint16), ("float32", np.float32), ("float64", np.float64)])
@pytest.mark.parametrize("shape", [(16, 128, 40), (64, 512, 12), (128, 1024, 5), (36, 1000, 9)])
def test_retained_difference_images(engine, dtype, np_dtype, shape):
    h, w, n = shape
    frames = small_clip(w=w, h=h, n=n, seed=h + n)
    res, _ = run_range(engine, frames, 12, DetectionParams(method="gradient"), diff_dtype=dtype, keep_decoded=True)
    want = fo.process_clip(frames, fo.ClipParams(method="gradient", keep_diffs=True))
    got = res.diff.cpu().numpy()
    assert got.dtype == np_dtype
    assert np.array_equal(got.astype(np.float64), want.diffs)
    assert np.array_equal(res.decoded.cpu().numpy(), frames)
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))


def test_generic_shapes(engine):
    """Frame sizes the TMA path cannot take (H*W % 32 != 0) go through the generic kernels."""
    for (h, w, bits) in [(5, 70, 12), (3, 33, 16), (9, 45, 8), (6, 31, 12)]:
        frames = small_clip(bits=bits, w=w, h=h, n=30, seed=w)
        for method in ("threshold", "gradient", "half_maximum"):
            res, _ = run_range(engine, frames, bits, DetectionParams(method=method), diff_dtype="float64")
            want = fo.process_clip(frames, fo.ClipParams(method=method, keep_diffs=True))
            assert np.array_equal(res.counts.cpu().numpy(), want.nonempty.astype(np.int32)), (h, w, bits)
            assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want)), (h, w, bits, method)
            assert np.array_equal(res.diff.cpu().numpy(), want.diffs)


def fuzz_frames(rng, n, h, w, maxval):
    """Piecewise-constant random rows: many ties, plateaus and runs for the tie-break rules."""
    frames = np.zeros((n, h, w), dtype=np.uint16)
    frames[0] = rng.integers(0, 40, size=(h, w))
    for i in range(1, n):
        row = np.zeros(w, dtype=np.int64)
        x = 0
        while x < w:
            seg = int(rng.integers(1, 9))
            row[x:x + seg] = rng.choice([0, 0, 30, 60, 200, 200, 1000, maxval])
            x += seg
        frames[i] = np.clip(row[None, :] + rng.integers(0, 3, size=(h, w)), 0, maxval)
    return frames


@pytest.mark.parametrize("w", [32, 33, 64, 95, 128, 257, 1024])
def test_detection_tie_breaks_on_random_profiles(engine, w):
    rng = np.random.default_rng(w)
    h = 4 if (4 * w) % 2 == 0 else 2
    frames = fuzz_frames(rng, 48, h, w, 4095)
    for method in ("threshold", "gradient", "half_maximum"):
        for use_diff in (True, False):
            for min_run in (1, 2, 5):
                if method != "threshold" and min_run != 1:
                    continue
                params = DetectionParams(method=method, use_frame_diff=use_diff, min_run_px=min_run, exit_margin_px=0)
                res, _ = run_range(engine, frames, 12, params, truncate=False)
                want = fo.process_clip(frames, fo.ClipParams(method=method, use_frame_diff=use_diff,
                                                             min_run_px=min_run, exit_margin_px=0))
                assert np.array_equal(res.pos.cpu().numpy(), want.pos_px), (w, method, use_diff, min_run)


def test_parameter_variants(engine):
    frames = small_clip(style="mini")
    for kw in [dict(frame_diff_threshold=0.0), dict(frame_diff_threshold=12.5), dict(min_gradient_strength=0.25),
               dict(min_gradient_strength=400.0), dict(exit_margin_px=15), dict(exit_margin_px=40),
               dict(min_signal_fraction=0.05), dict(min_signal_fraction=0.0)]:
        for method in ("gradient", "threshold"):
            res, _ = run_range(engine, frames, 12, DetectionParams(method=method, **kw))
            want = fo.process_clip(frames, fo.ClipParams(method=method, **kw))
            assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want)), (kw, method)


def test_skip_frames(engine):
    frames = small_clip()
    skip = [0, 11, 12, 30, 71]
    mask = np.zeros(len(frames), dtype=np.uint8)
    mask[skip] = 1
    for method in ("half_maximum", "threshold"):
        res, _ = run_range(engine, frames, 12, DetectionParams(method=method), skip=dev(mask, engine),
                           diff_dtype="uint16")
        want = fo.process_clip(frames, fo.ClipParams(method=method, skip_frames=skip, keep_diffs=True))
        assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
        assert np.array_equal(res.diff.cpu().numpy().astype(np.float64), want.diffs)


def test_subrange_with_halo_equals_serial(engine):
    """A contiguous range + one-frame halo reproduces the serial run (multi-GPU sharding rule)."""
    frames = small_clip(n=96, w=256, h=32, seed=5)
    n = len(frames)
    params = DetectionParams(method="half_maximum")
    serial = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    pieces, exits = [], []
    for a, b in [(0, 31), (31, 64), (64, 96)]:
        halo = None if a == 0 else dev(syn.pack_frames(frames[a - 1:a], 12), engine)
        res, _ = run_range(engine, frames[a:b], 12, params, frame0=frames[0], first_frame=a, halo=halo,
                           truncate=False, diff_dtype="uint16")
        pieces.append(res.pos.cpu().numpy())
        exits.append(int(res.first_exit.cpu().item()))
        d = fo.process_clip(frames[a:b], fo.ClipParams(method="half_maximum", keep_diffs=True), frame0=frames[0],
                            first_index=a, prior_frame=frames[a - 1] if a else None).diffs
        assert np.array_equal(res.diff.cpu().numpy().astype(np.float64), d)
    assert np.array_equal(np.concatenate(pieces), serial.pos_px)
    assert min(exits) == (serial.first_exit if serial.first_exit < n else FF_NO_EXIT)


def test_big_tile_kernels_with_skips_halo_and_ragged_tiles(engine):
    """Frames large enough for the 8192-pixel tiles (count12_kernel / stream12_kernel): ragged last
    tile (130 rows -> 16.25 tiles), CTAs whose run of items crosses tile boundaries, skip_frames,
    a sub-range with a halo frame, uint16 difference + decoded output, all against the oracle."""
    frames = small_clip(w=1024, h=130, n=23, seed=41, style="mini")
    n = len(frames)
    skip = [3, 4, 9, 22]
    mask = np.zeros(n, dtype=np.uint8)
    mask[skip] = 1
    want = fo.process_clip(frames, fo.ClipParams(method="threshold", skip_frames=skip, keep_diffs=True))
    res, _ = run_range(engine, frames, 12, DetectionParams(method="threshold"), skip=dev(mask, engine),
                       diff_dtype="uint16", keep_decoded=True)
    live = mask == 0          # the reference never looks at skip_frames entries (:1443-1445): counts undefined there
    assert np.array_equal(res.counts.cpu().numpy()[live], want.nonempty.astype(np.int32)[live])
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    assert np.array_equal(res.diff.cpu().numpy().astype(np.float64), want.diffs)
    assert np.array_equal(res.decoded.cpu().numpy(), frames)
    # counts only (count12_kernel) and difference only (no decoded output)
    plain, _ = run_range(engine, frames, 12, DetectionParams(method="threshold"), skip=dev(mask, engine))
    assert np.array_equal(plain.counts.cpu().numpy()[live], want.nonempty.astype(np.int32)[live])
    assert np.array_equal(plain.pos.cpu().numpy(), oracle_pos(want))
    only, _ = run_range(engine, frames, 12, DetectionParams(method="threshold"), skip=dev(mask, engine),
                        diff_dtype="uint16")
    assert np.array_equal(only.diff.cpu().numpy().astype(np.float64), want.diffs)
    # sub-range [10, 23) whose prior frame (9) is skipped: the halo is frame 8
    a = 10
    halo = dev(syn.pack_frames(frames[8:9], 12), engine)
    sub, _ = run_range(engine, frames[a:], 12, DetectionParams(method="threshold"), frame0=frames[0], first_frame=a,
                       halo=halo, skip=dev(mask[a:], engine), diff_dtype="uint16", truncate=False)
    assert np.array_equal(sub.diff.cpu().numpy().astype(np.float64), want.diffs[a:])
    assert np.array_equal(sub.pos.cpu().numpy(), want.pos_px[a:])
    # zero difference threshold keeps every non-negative difference
    z, _ = run_range(engine, frames[:6], 12, DetectionParams(method="gradient", frame_diff_threshold=0.0),
                     diff_dtype="uint16")
    wz = fo.process_clip(frames[:6], fo.ClipParams(method="gradient", frame_diff_threshold=0.0, keep_diffs=True))
    assert np.array_equal(z.diff.cpu().numpy().astype(np.float64), wz.diffs)
    # 16-bit and 8-bit storage of large frames (general template, 8192-pixel tiles)
    for bits in (16, 8):
        fr = small_clip(bits=bits, w=1024, h=130, n=7, seed=bits, style="mini")
        r, _ = run_range(engine, fr, bits, DetectionParams(method="gradient"), diff_dtype="uint16")
        w_ = fo.process_clip(fr, fo.ClipParams(method="gradient", keep_diffs=True))
        assert np.array_equal(r.counts.cpu().numpy(), w_.nonempty.astype(np.int32)), bits
        assert np.array_equal(r.pos.cpu().numpy(), oracle_pos(w_)), bits
        assert np.array_equal(r.diff.cpu().numpy().astype(np.float64), w_.diffs), bits


@pytest.mark.parametrize("dtype,np_dtype", [("float32", np.float32), ("float64", np.float64)])
def test_big_tile_float_difference(engine, dtype, np_dtype):
    """float32 / float64 difference image at the 8192-pixel tile size without a decoded output (streamx_kernel's float path:
    16x2 arithmetic, lanes widened on the way out, lane pairs swapping halves before the stores): ragged last tile,
    skip_frames, a sub-range whose halo is an earlier frame, zero threshold, 12-, 16- and 8-bit storage."""
    frames = small_clip(w=1024, h=130, n=23, seed=43, style="mini")
    n = len(frames)
    skip = [3, 4, 9, 22]
    mask = np.zeros(n, dtype=np.uint8)
    mask[skip] = 1
    live = mask == 0
    want = fo.process_clip(frames, fo.ClipParams(method="threshold", skip_frames=skip, keep_diffs=True))
    res, _ = run_range(engine, frames, 12, DetectionParams(method="threshold"), skip=dev(mask, engine),
                       diff_dtype=dtype)
    got = res.diff.cpu().numpy()
    assert got.dtype == np_dtype and np.array_equal(got.astype(np.float64), want.diffs)
    assert np.array_equal(res.counts.cpu().numpy()[live], want.nonempty.astype(np.int32)[live])
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    a = 10
    halo = dev(syn.pack_frames(frames[8:9], 12), engine)
    sub, _ = run_range(engine, frames[a:], 12, DetectionParams(method="threshold"), frame0=frames[0], first_frame=a,
                       halo=halo, skip=dev(mask[a:], engine), diff_dtype=dtype, truncate=False)
    assert np.array_equal(sub.diff.cpu().numpy().astype(np.float64), want.diffs[a:])
    z, _ = run_range(engine, frames[:6], 12, DetectionParams(method="gradient", frame_diff_threshold=0.0),
                     diff_dtype=dtype)
    wz = fo.process_clip(frames[:6], fo.ClipParams(method="gradient", frame_diff_threshold=0.0, keep_diffs=True))
    assert np.array_equal(z.diff.cpu().numpy().astype(np.float64), wz.diffs)
    for bits in (16, 8):
        fr = small_clip(bits=bits, w=1024, h=130, n=7, seed=bits + 1, style="mini")
        r, _ = run_range(engine, fr, bits, DetectionParams(method="gradient"), diff_dtype=dtype)
        w_ = fo.process_clip(fr, fo.ClipParams(method="gradient", keep_diffs=True))
        assert np.array_equal(r.counts.cpu().numpy(), w_.nonempty.astype(np.int32)), bits
        assert np.array_equal(r.pos.cpu().numpy(), oracle_pos(w_)), bits
        assert np.array_equal(r.diff.cpu().numpy().astype(np.float64), w_.diffs), bits
    # a long clip (frame-synchronous units with their halo items; >= 96 frames) of full tiles
    long = small_clip(w=1024, h=128, n=130, seed=7, style="nova")
    wl = fo.process_clip(long, fo.ClipParams(method="half_maximum", keep_diffs=True))
    for dt in (dtype, "uint16"):
        rl, _ = run_range(engine, long, 12, DetectionParams(method="half_maximum"), diff_dtype=dt)
        assert np.array_equal(rl.diff.cpu().numpy().astype(np.float64), wl.diffs), dt
        assert np.array_equal(rl.pos.cpu().numpy(), oracle_pos(wl)), dt


def test_empty_and_degenerate_inputs(engine):
    # all-dark clip: every frame empty, nothing detected, no exit
    dark = np.full((10, 8, 64), 40, dtype=np.uint16)
    res, _ = run_range(engine, dark, 12, DetectionParams(method="threshold"))
    assert res.pos.cpu().numpy().tolist() == [-1] * 10 and int(res.first_exit.cpu().item()) == FF_NO_EXIT
    assert res.counts.cpu().numpy().tolist() == [0] * 10
    # a single frame: no prior -> no detection in difference mode
    one = small_clip(n=1)
    res, _ = run_range(engine, one, 12, DetectionParams(method="gradient"))
    assert res.pos.cpu().numpy().tolist() == [-1]
    # saturated 12-bit frames (maximum values)
    sat = np.full((4, 8, 64), 4095, dtype=np.uint16)
    sat[0] = 10
    res, _ = run_range(engine, sat, 12, DetectionParams(method="threshold", use_frame_diff=False))
    want = fo.process_clip(sat, fo.ClipParams(method="threshold", use_frame_diff=False))
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    with pytest.raises(ValueError):
        run_range(engine, small_clip(n=2, w=1, h=2), 16, DetectionParams(method="gradient"))   # np.gradient needs 2
    with pytest.raises(ValueError):
        run_range(engine, one, 12, DetectionParams(method="gradient", frame_diff_threshold=-3.0), diff_dtype="uint16")


# ------------------------------------------------------------------- host-resident streaming
@pytest.mark.parametrize("method", ["threshold", "half_maximum"])
def test_process_host_equals_device_path_and_oracle(method):
    from high_speed_image_processing_b200.engine import FlameFrontEngine
    frames = small_clip(n=200, w=128, h=16, style="mini", seed=3)
    n, h, w = frames.shape
    packed = syn.pack_frames(frames, 12)
    fb = h * w * 3 // 2
    eng = FlameFrontEngine(0, host_chunk_bytes=fb * 7)        # 7 frames per chunk: many chunks + early exit
    scalars, _ = eng.clip_scalars(dev(packed[:fb], eng), h, w, 12)
    params = DetectionParams(method=method)
    want = fo.process_clip(frames, fo.ClipParams(method=method))
    assert want.first_exit < n
    for src in (packed, torch.from_numpy(packed.copy()).pin_memory()):
        got = eng.process_host(src, n, h, w, 12, params, scalars)
        assert got.first_exit == want.first_exit
        assert np.array_equal(got.pos, oracle_pos(want))
        assert got.frames_done < n, "copy should stop after the exit frame was seen"
        assert np.array_equal(got.counts[:got.frames_done], want.nonempty[:got.frames_done].astype(np.int32))
    # profile = bg-subtracted centre row instead of the frame difference (use_frame_diff=False)
    nd = DetectionParams(method=method, use_frame_diff=False)
    want_nd = fo.process_clip(frames, fo.ClipParams(method=method, use_frame_diff=False))
    got = eng.process_host(packed, n, h, w, 12, nd, scalars)
    assert got.first_exit == (want_nd.first_exit if want_nd.first_exit < n else FF_NO_EXIT)
    assert np.array_equal(got.pos, oracle_pos(want_nd))
    # with skip frames and a sub-range halo
    skip = np.zeros(n, dtype=np.uint8)
    skip[[6, 7, 13, 14, 20]] = 1
    a = 14
    got = eng.process_host(packed[a * fb:], n - a, h, w, 12, DetectionParams(method=method, exit_margin_px=0), scalars,
                           first_frame=a, halo=packed[12 * fb:13 * fb], skip=skip[a:])
    full = fo.process_clip(frames, fo.ClipParams(method=method, exit_margin_px=0,
                                                 skip_frames=[6, 7, 13, 14, 20]))
    assert np.array_equal(got.pos, full.pos_px[a:])
    eng.close()


# ------------------------------------------------------------------------ output bounds (canaries)
@pytest.mark.parametrize("h,w,n", [(130, 1024, 5), (36, 1000, 7), (5, 70, 9), (128, 1024, 3)])
@pytest.mark.parametrize("bits", [8, 12, 16])
def test_kernels_stay_inside_their_output_buffers(engine, h, w, n, bits):
    """compute-sanitizer is not available on the GPU pool, so every output of the streaming and
    detection entry points is placed between canary regions and the canaries are checked."""
    import ctypes as C
    from high_speed_image_processing_b200._cabi import FF_DIFF_F32, FF_DIFF_F64, FF_DIFF_NONE, FF_DIFF_U16
    lib, dev_ = engine._lib, engine.device
    frames = small_clip(bits=bits, w=w, h=h, n=n, seed=h + w + bits, style="mini")
    packed = dev(syn.pack_frames(frames, bits), engine)
    bg = torch.tensor([int(frames[0].max())], dtype=torch.int32, device=dev_)
    n_elems, per_frame = C.c_int64(0), C.c_int(0)
    assert lib.ff_partial_len(n, h, w, bits, C.byref(n_elems), C.byref(per_frame)) == 0
    guard = 1024                                            # elements of canary on each side

    def padded(n_items, dtype):
        buf = torch.full((n_items + 2 * guard,), 0x5A, dtype=torch.uint8, device=dev_).repeat_interleave(
            torch.tensor([], dtype=dtype).element_size()).view(dtype)
        return buf, buf[guard:guard + n_items]

    def intact(buf, n_items):
        raw = buf.view(torch.uint8)
        es = buf.element_size()
        return bool((raw[:guard * es] == 0x5A).all()) and bool((raw[(guard + n_items) * es:] == 0x5A).all())

    st = torch.cuda.current_stream().cuda_stream
    px = n * h * w
    cases = [(FF_DIFF_NONE, None, False), (FF_DIFF_U16, torch.uint16, False), (FF_DIFF_F32, torch.float32, False),
             (FF_DIFF_F64, torch.float64, False)]
    if bits == 12:
        cases += [(FF_DIFF_NONE, None, True), (FF_DIFF_U16, torch.uint16, True)]
    for code, dt, decoded in cases:
        pbuf, pview = padded(n_elems.value, torch.int32)
        dbuf, dview = padded(px, dt) if dt is not None else (None, None)
        ebuf, eview = padded(px, torch.uint16) if decoded else (None, None)
        rc = lib.ff_stream_frames(packed.data_ptr(), None, n, h, w, bits, bg.data_ptr(), -1, 5, None,
                                  pview.data_ptr(), None if dview is None else dview.data_ptr(), code,
                                  None if eview is None else eview.data_ptr(), st)
        assert rc == 0, (code, decoded, rc)
        torch.cuda.synchronize()
        assert intact(pbuf, n_elems.value), ("partial", code, decoded)
        if dbuf is not None:
            assert intact(dbuf, px), ("diff", code)
        if ebuf is not None:
            assert intact(ebuf, px), ("decoded", code)
        # detection outputs
        posb, posv = padded(n, torch.int32)
        cntb, cntv = padded(n, torch.int32)
        prob, prov = padded(n * w, torch.int32)
        fe = torch.full((1,), FF_NO_EXIT, dtype=torch.int32, device=dev_)
        rc = lib.ff_detect(packed.data_ptr(), None, n, 0, h, w, bits, bg.data_ptr(), pview.data_ptr(), 1, 1, 1, 5,
                           100, -20, 1, 10, None, posv.data_ptr(), cntv.data_ptr(), fe.data_ptr(), prov.data_ptr(), st)
        assert rc == 0
        torch.cuda.synchronize()
        assert intact(posb, n) and intact(cntb, n) and intact(prob, n * w)
    if bits == 12:
        ubuf, uview = padded(px, torch.uint16)
        assert lib.ff_unpack(packed.data_ptr(), uview.data_ptr(), n, h, w, 12, st) == 0
        torch.cuda.synchronize()
        assert intact(ubuf, px) and np.array_equal(uview.cpu().numpy().reshape(n, h, w), frames)
