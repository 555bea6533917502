"""ff_process_range (prep kernel + fused range kernel / stream + detect) and the single-rank form of the
peer exchange, through the C-ABI, against the oracle.  Bit-exact comparisons."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED, RangeHooks
from high_speed_image_processing_b200.engine import ClipScalars, DetectionParams
from oracle import flame_oracle as fo

pytestmark = pytest.mark.gpu


def dev(a, engine):
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


def oracle_pos(res):
    want = res.pos_px.copy()
    want[res.first_exit:] = FF_POS_DROPPED
    return want


def clip(w, h, n, bits=12, style="nova", t_enter=4.0, velocity=None, seed=7):
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=n, bits=bits, style=style, t_enter=t_enter,
                             velocity=velocity if velocity is not None else w / (0.6 * n), seed=seed)
    return spec, syn.render_frames(spec)


# ------------------------------------------------------------------- prep kernel: scalars on the device
@pytest.mark.parametrize("bits", [8, 12, 16])
@pytest.mark.parametrize("h,w", [(16, 128), (7, 50), (1, 64), (128, 1024), (33, 640), (4, 3072), (9, 130), (3, 6),
                                 (64, 1281 + 1), (2, 7 * 2)])
def test_device_threshold_is_numpys_float64(engine, bits, h, w):
    """mean / std / max of the centre row and max(mean + 5 std, 2 max) from prep_kernel (NumPy's pairwise
    add.reduce order) equal NumPy's own float64 bit for bit, for widths on both sides of every leaf boundary."""
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=3, bits=bits, seed=h * w + bits, noise_std=9.0)
    frames = syn.render_frames(spec)
    if bits == 16:
        frames = (frames.astype(np.uint32) * 97 % 65536).astype(np.uint16)
    packed = dev(syn.pack_frames(frames, bits), engine)
    res = engine.process_range(packed, 3, h, w, bits, DetectionParams(method="threshold"))
    want = ClipScalars.from_frame0_stats(int(frames[0].max()), frames[0][h // 2])
    pending = res._scalars
    got = res.scalars                       # NumPy on the copied-back row + cross-check of the device's floor
    assert got == want
    pend_block = pending.block              # the clip-scalar block the kernels used
    stats = pend_block[4:12].view(np.float64)
    assert int(pend_block[0]) == int(frames[0].max())
    assert (stats[0], stats[1], stats[2], stats[3]) == (want.centerline_mean, want.centerline_std, want.centerline_max,
                                                        want.flame_threshold)
    assert int(pend_block[1]) == int(np.floor(want.flame_threshold))


# ------------------------------------------------------------------- fused range kernel vs oracle
@pytest.mark.parametrize("method", ["half_maximum", "threshold", "gradient"])
@pytest.mark.parametrize("bits,h,w,n", [(12, 128, 1024, 150), (12, 130, 1024, 90), (16, 128, 1024, 60), (8, 256, 1024, 60),
                                        (12, 256, 512, 333)])
def test_fused_range_equals_oracle_and_unfused(engine, method, bits, h, w, n):
    spec, frames = clip(w, h, n, bits=bits, style="mini" if method == "threshold" else "nova")
    packed = dev(syn.pack_frames(frames, bits), engine)
    want = fo.process_clip(frames, fo.ClipParams(method=method))
    params = DetectionParams(method=method)
    plan = [C.c_int64(0), C.c_int64(0), C.c_int(0)]
    engine._lib.ff_process_range_plan(n, h, w, bits, 0, 0, 0, *(C.byref(x) for x in plan))
    assert plan[2].value == 1 and plan[1].value == 0, "this shape must run as ONE kernel"
    before = engine.launches
    res = engine.process_range(packed, n, h, w, bits, params)
    assert engine.launches == before + 2                       # prep + range kernel
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    assert np.array_equal(res.counts.cpu().numpy(), want.nonempty.astype(np.int32))
    assert int(res.first_exit.item()) == (want.first_exit if want.first_exit < n else FF_NO_EXIT)
    assert res.scalars.flame_threshold == want.flame_threshold
    unt = engine.process_range(packed, n, h, w, bits, params, truncate=False)
    assert np.array_equal(unt.pos.cpu().numpy(), want.pos_px)
    # the workspace's tickets and per-frame arrival words are left zero
    assert int(engine._ws[:16].view(torch.int64).abs().sum().item()) == 0
    assert int(engine._ws[512:].view(torch.int64).abs().sum().item()) == 0
    # the three-kernel form gives the same answers
    os.environ["FF_RANGE_UNFUSED"] = "1"
    try:
        res2 = engine.process_range(packed, n, h, w, bits, params)
    finally:
        del os.environ["FF_RANGE_UNFUSED"]
    assert torch.equal(res2.pos, res.pos) and torch.equal(res2.counts, res.counts) and torch.equal(res2.first_exit, res.first_exit)


@pytest.mark.parametrize("method", ["half_maximum", "threshold", "gradient"])
@pytest.mark.parametrize("bits,h,w,n", [(12, 64, 512, 140), (12, 66, 512, 77), (16, 64, 512, 60), (8, 64, 512, 60),
                                        (12, 32, 512, 50), (12, 16, 1024, 64), (8, 16, 1024, 7)])
def test_range_kernel_on_small_frames(engine, method, bits, h, w, n):
    """Frames of BASELINE config 1's size (a few 12-KiB items per frame, ragged last item, frames smaller than one
    item for 8-bit pixels) through the fused range kernel (FF_RANGE_FUSE_SMALL): the oracle's answers, the
    three-kernel form's answers, skip_frames and a sub-range with a halo frame."""
    spec, frames = clip(w, h, n, bits=bits, style="mini" if method == "threshold" else "nova")
    packed = dev(syn.pack_frames(frames, bits), engine)
    params = DetectionParams(method=method)
    skip = [0, 5, 6, n - 1]
    skip_np = np.zeros(n, np.uint8)
    skip_np[skip] = 1
    want = fo.process_clip(frames, fo.ClipParams(method=method))
    want_s = fo.process_clip(frames, fo.ClipParams(method=method, skip_frames=skip))
    fb = spec.frame_bytes
    os.environ["FF_RANGE_FUSE_SMALL"] = "1"
    try:
        plan = [C.c_int64(0), C.c_int64(0), C.c_int(0)]
        engine._lib.ff_process_range_plan(n, h, w, bits, 0, 0, 0, *(C.byref(x) for x in plan))
        assert plan[2].value == 1 and plan[1].value == 0, "this shape must run as ONE kernel"
        before = engine.launches
        res = engine.process_range(packed, n, h, w, bits, params)
        assert engine.launches == before + 2                   # prep + range kernel
        res_s = engine.process_range(packed, n, h, w, bits, params, skip=dev(skip_np, engine))
        a = n // 2
        sub = engine.process_range(packed[a * fb:].clone(), n - a, h, w, bits, params, frame0=packed[:fb], first_frame=a,
                                   halo=packed[(a - 1) * fb:a * fb].clone(), truncate=False)
        assert int(engine._ws[:16].view(torch.int64).abs().sum().item()) == 0
        assert int(engine._ws[512:].view(torch.int64).abs().sum().item()) == 0
    finally:
        del os.environ["FF_RANGE_FUSE_SMALL"]
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    assert np.array_equal(res.counts.cpu().numpy(), want.nonempty.astype(np.int32))
    assert int(res.first_exit.item()) == (want.first_exit if want.first_exit < n else FF_NO_EXIT)
    assert res.scalars.flame_threshold == want.flame_threshold
    assert np.array_equal(res_s.pos.cpu().numpy(), oracle_pos(want_s))
    keep = skip_np == 0
    assert np.array_equal(res_s.counts.cpu().numpy()[keep], want_s.nonempty.astype(np.int32)[keep])
    assert np.array_equal(sub.pos.cpu().numpy(), want.pos_px[a:])
    os.environ["FF_RANGE_UNFUSED"] = "1"
    try:
        res2 = engine.process_range(packed, n, h, w, bits, params)
    finally:
        del os.environ["FF_RANGE_UNFUSED"]
    assert torch.equal(res2.pos, res.pos) and torch.equal(res2.counts, res.counts) and torch.equal(res2.first_exit, res.first_exit)


def test_fused_range_skip_frames_halo_and_subranges(engine):
    h, w, n = 128, 1024, 120
    spec, frames = clip(w, h, n, t_enter=10.0, velocity=11.0)
    skip = [0, 30, 31, 32, 60, 119]
    want = fo.process_clip(frames, fo.ClipParams(method="half_maximum", skip_frames=skip))
    packed = dev(syn.pack_frames(frames, 12), engine)
    fb = spec.frame_bytes
    skip_np = np.zeros(n, np.uint8)
    skip_np[skip] = 1
    params = DetectionParams(method="half_maximum")
    res = engine.process_range(packed, n, h, w, 12, params, skip=dev(skip_np, engine))
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    keep = skip_np == 0                  # (skipped frames still get their pixel count; the oracle leaves 0 there)
    assert np.array_equal(res.counts.cpu().numpy()[keep], want.nonempty.astype(np.int32)[keep])
    # any split into sub-ranges with a halo frame (the latest non-skipped frame before the range) agrees
    pos = np.full(n, -9, np.int32)
    fe = FF_NO_EXIT
    for a, b in ((0, 31), (31, 33), (33, 61), (61, n)):
        hidx = a - 1
        while hidx >= 0 and skip_np[hidx]:
            hidx -= 1
        halo = packed[hidx * fb:(hidx + 1) * fb].clone() if hidx >= 0 else None
        r = engine.process_range(packed[a * fb:b * fb].clone(), b - a, h, w, 12, params, frame0=packed[:fb], first_frame=a,
                                 halo=halo, skip=dev(skip_np[a:b], engine), truncate=False)
        pos[a:b] = r.pos.cpu().numpy()
        fe = min(fe, int(r.first_exit.item()))
    assert np.array_equal(pos, want.pos_px)
    assert fe == (want.first_exit if want.first_exit < n else FF_NO_EXIT)


def test_every_frame_holds_a_flame(engine):
    """All frames reach the detector warp (queue back-pressure) - and none when the clip is empty."""
    h, w, n = 128, 1024, 700
    spec = syn.SyntheticSpec(width=w, height=h, n_frames=n, style="mini", t_enter=-3000.0, velocity=0.5, x_enter=0.0, seed=3)
    frames = syn.render_frames(spec)
    want = fo.process_clip(frames, fo.ClipParams(method="threshold", exit_margin_px=0))
    assert (want.nonempty[1:] > 0.0005 * h * w).all()
    packed = dev(syn.pack_frames(frames, 12), engine)
    res = engine.process_range(packed, n, h, w, 12, DetectionParams(method="threshold", exit_margin_px=0))
    assert np.array_equal(res.pos.cpu().numpy(), oracle_pos(want))
    assert np.array_equal(res.counts.cpu().numpy(), want.nonempty.astype(np.int32))
    dark = np.full((40, h, w), 37, np.uint16)
    res = engine.process_range(dev(syn.pack_frames(dark, 12), engine), 40, h, w, 12, DetectionParams())
    assert (res.pos.cpu().numpy() == -1).all() and int(res.first_exit.item()) == FF_NO_EXIT
    assert (res.counts.cpu().numpy() == 0).all()


# ------------------------------------------------------------------- exchange protocol on one rank
def _single_rank_exchange(engine, cap):
    lib = engine._lib
    x = C.c_void_p()
    assert lib.ff_exchange_create(engine.device.index, 0, 1, cap, C.byref(x)) == 0
    return x


def _block_views(engine, x, cap):
    lib = engine._lib
    pos_p, cnt_p, fe_p, hooks = C.c_void_p(), C.c_void_p(), C.c_void_p(), RangeHooks()
    assert lib.ff_exchange_begin(x, C.byref(pos_p), C.byref(cnt_p), C.byref(fe_p), C.byref(hooks)) == 0
    from high_speed_image_processing_b200.sharding import _DevicePointerView

    def view(p, n):
        return torch.as_tensor(_DevicePointerView(p.value, n), device=engine.device)
    return view(pos_p, cap), view(cnt_p, cap), view(fe_p, 1), hooks


@pytest.mark.parametrize("fused", [True, False])
def test_range_kernel_publishes_and_merge_acknowledges(engine, fused):
    """world = 1 exercises the whole protocol on one GPU: prep waits for the acks of epoch-2, the last CTA of
    the range kernel (or of ff_detect) publishes, the merge kernel waits for the flag, truncates against
    the global exit, resets the exit word and acknowledges - five epochs over the two block slots."""
    h, w, n = 128, 1024, 96
    spec, frames = clip(w, h, n, t_enter=6.0, velocity=14.0)
    want = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    assert want.first_exit < n
    packed = dev(syn.pack_frames(frames, 12), engine)
    lib = engine._lib
    x = _single_rank_exchange(engine, n)
    st = torch.cuda.current_stream(engine.device).cuda_stream
    try:
        for epoch in range(5):
            pos_v, cnt_v, fe_v, hooks = _block_views(engine, x, n)
            assert hooks.epoch == epoch + 1 and hooks.world == 1
            engine.process_range(packed, n, h, w, 12, DetectionParams(method="half_maximum"), truncate=False,
                                 pos_out=pos_v, counts_out=cnt_v, first_exit=fe_v, hooks=hooks, init_first_exit=True,
                                 diff_dtype=None if fused else "uint16")
            pos = torch.full((n,), 77, dtype=torch.int32, device=engine.device)
            cnt = torch.full((n,), 77, dtype=torch.int32, device=engine.device)
            fe = torch.zeros(1, dtype=torch.int32, device=engine.device)
            assert lib.ff_exchange_finish(x, n, pos.data_ptr(), cnt.data_ptr(), fe.data_ptr(), st) == 0
            assert int(fe.item()) == want.first_exit
            assert np.array_equal(pos.cpu().numpy(), oracle_pos(want))
            assert np.array_equal(cnt.cpu().numpy(), want.nonempty.astype(np.int32))
            exit_word = torch.as_tensor(
                __import__("high_speed_image_processing_b200.sharding", fromlist=["x"])._DevicePointerView(hooks.exit_word_dev, 1),
                device=engine.device)
            assert int(exit_word.item()) == FF_NO_EXIT          # reset by the merge for epoch + 2
            status = C.c_int32(-1)
            assert lib.ff_exchange_status(x, C.byref(status), st) == 0 and status.value == 0
    finally:
        lib.ff_exchange_destroy(x)


def test_host_streamed_range_into_a_block_stops_at_the_global_exit(engine):
    """ff_process_host_range with a block + hooks: results stay on the device, the block is published, and the
    upload stops at the first chunk behind the smallest exit frame in the rank's exit word - also when that
    exit was found by ANOTHER rank (simulated by writing the word before the call)."""
    from high_speed_image_processing_b200.engine import FlameFrontEngine
    from high_speed_image_processing_b200.sharding import RangeBlock, _DevicePointerView
    h, w, n = 128, 1024, 200
    spec, frames = clip(w, h, n, t_enter=5.0, velocity=9.0)
    want = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    assert 100 < want.first_exit < 140
    packed_np = syn.pack_frames(frames, 12)
    fb = spec.frame_bytes
    eng = FlameFrontEngine(engine.device.index, host_chunk_bytes=8 * fb)      # 8-frame chunks
    scalars, _ = eng.clip_scalars(dev(packed_np[:fb], eng), h, w, 12)
    lib = eng._lib
    x = _single_rank_exchange(eng, n)
    st = torch.cuda.current_stream(eng.device).cuda_stream
    try:
        for case in ("own_exit", "peer_exit_before_range", "peer_exit_inside_range"):
            pos_v, cnt_v, fe_v, hooks = _block_views(eng, x, n)
            blk = RangeBlock(n, n, pos_v, cnt_v, fe_v, None, hooks, True)
            first, a = 0, 0
            if case != "own_exit":      # this rank owns frames [40, 100): no exit of its own
                first, a = 40, 40
                word = torch.as_tensor(_DevicePointerView(hooks.exit_word_dev, 1), device=eng.device)
                # (the real writer is a peer's detecting warp; the prep kernel does not touch the word)
                word.fill_(12 if case == "peer_exit_before_range" else 70)
                torch.cuda.synchronize()
            b = n if case == "own_exit" else 100
            halo = packed_np[(a - 1) * fb:a * fb] if a else None
            res = eng.process_host(packed_np[a * fb:b * fb], b - a, h, w, 12, DetectionParams(method="half_maximum"), scalars,
                                   first_frame=first, halo=halo, block=blk, hooks=hooks, to_host=False)
            pos = torch.full((n,), 77, dtype=torch.int32, device=eng.device)
            fe = torch.zeros(1, dtype=torch.int32, device=eng.device)
            assert lib.ff_exchange_finish(x, b - a, pos.data_ptr(), None, fe.data_ptr(), st) == 0
            torch.cuda.synchronize()
            if case == "own_exit":
                assert res.first_exit == want.first_exit and int(fe.item()) == want.first_exit
                assert res.frames_done < n and res.frames_done >= want.first_exit       # stopped soon after the exit
                assert res.frames_done <= want.first_exit + 3 * 8
                assert np.array_equal(pos.cpu().numpy()[:n], oracle_pos(want))
            elif case == "peer_exit_before_range":
                assert res.frames_done == 0 and res.bytes_uploaded == 0                 # nothing uploaded at all
                assert res.first_exit == FF_NO_EXIT
            else:
                assert 70 - 40 <= res.frames_done <= 70 - 40 + 8                         # the chunk that reaches frame 70
                got = pos.cpu().numpy()[:res.frames_done]
                assert np.array_equal(got[:31], want.pos_px[40:71])                      # frames up to the peer's exit frame
                # frames behind it will be dropped by the merge: the detecting warps may skip them (position "none")
                assert all(g in (-1, w) for g, w in zip(got[31:], want.pos_px[71:40 + res.frames_done]))
    finally:
        lib.ff_exchange_destroy(x)
        eng.close()
