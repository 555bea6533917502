"""Multi-rank host logic on CPU: range partition, exit-min all-reduce, result all-gather,
global truncation.  world_size 2 and 3 over gloo; the per-rank 'kernel output' is the oracle
evaluated on the rank's contiguous range with its one-frame halo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from high_speed_image_processing_b200 import synthetic as syn
from high_speed_image_processing_b200.sharding import (BLOCK_HEADER, RangeExchange, assign_videos,
                                                       bind_to_gpu_numa_node, contiguous_range)
from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED
from oracle import flame_oracle as fo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torch_merge(gathered, world, cap, total, pos_out, counts_out, first_exit_out):
    """CPU stand-in for the ff_merge_ranges kernel (csrc/ff_exchange.cu) - same block layout."""
    blocks = gathered.view(world, BLOCK_HEADER + 2 * cap)
    fe = int(blocks[:, 0].min())
    first_exit_out[0] = fe
    for r in range(world):
        a, b = contiguous_range(total, r, world)
        pos_out[a:b] = blocks[r, BLOCK_HEADER:BLOCK_HEADER + (b - a)]
        if counts_out is not None:
            counts_out[a:b] = blocks[r, BLOCK_HEADER + cap:BLOCK_HEADER + cap + (b - a)]
    pos_out[min(fe, total):] = FF_POS_DROPPED


def _worker(rank, size, port, n_frames, margin, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        spec = syn.SyntheticSpec(width=96, height=8, n_frames=n_frames, style="mini", t_enter=6.0, velocity=2.0,
                                 curvature_px=1.0, seed=11)
        frames = syn.render_frames(spec)
        ex = RangeExchange()
        a, b = ex.my_range(n_frames)
        params = fo.ClipParams(method="threshold", exit_margin_px=margin)
        part = fo.process_clip(frames[a:b], params, frame0=frames[0], first_index=a,
                               prior_frame=frames[a - 1] if a > 0 else None)
        fe = FF_NO_EXIT if part.first_exit == b - a else a + part.first_exit
        if rank % 2 == 0:      # in place, the way ff_detect fills a block
            blk = ex.begin(n_frames, torch.device("cpu"))
            assert int(blk.first_exit[0]) == FF_NO_EXIT and blk.pos.numel() == ex.block_cap(n_frames)
            blk.pos[:b - a] = torch.from_numpy(part.pos_px.copy())
            blk.counts[:b - a] = torch.from_numpy(part.nonempty.astype(np.int32))
            blk.first_exit[0] = fe
            g = ex.finish(blk, _torch_merge)
        else:                  # results that came from elsewhere (the host-streamed path)
            g = ex.finish_arrays(torch.from_numpy(part.pos_px.copy()), torch.tensor([fe], dtype=torch.int32),
                                 n_frames, _torch_merge, counts_local=torch.from_numpy(part.nonempty.astype(np.int32)))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), pos=g.pos.numpy(), counts=g.counts.numpy(),
                 first_exit=g.first_exit)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,n_frames", [(2, 61), (3, 64)])
def test_range_sharded_run_equals_serial(tmp_path, size, n_frames):
    margin = 10
    mp.spawn(_worker, args=(size, _free_port(), n_frames, margin, str(tmp_path)), nprocs=size, join=True)
    spec = syn.SyntheticSpec(width=96, height=8, n_frames=n_frames, style="mini", t_enter=6.0, velocity=2.0,
                             curvature_px=1.0, seed=11)
    serial = fo.process_clip(syn.render_frames(spec), fo.ClipParams(method="threshold", exit_margin_px=margin))
    assert serial.first_exit < n_frames, "fixture must contain an exit"
    want = serial.pos_px.copy()
    want[serial.first_exit:] = FF_POS_DROPPED
    for r in range(size):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["first_exit"]) == serial.first_exit
        assert np.array_equal(z["pos"], want)
        assert np.array_equal(z["counts"], serial.nonempty.astype(np.int32))


def test_contiguous_range_is_the_reference_partition(golden):
    for key, want in golden["distribute_indices"].items():
        total, size, strategy = key.split("/")
        if strategy != "contiguous":
            continue
        for r in range(int(size)):
            a, b = contiguous_range(int(total), r, int(size))
            assert list(range(a, b)) == want[r]
    with pytest.raises(ValueError):
        contiguous_range(5, 3, 3)


def test_assign_videos():
    assert [assign_videos(10, r, 4) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    w = [5, 1, 1, 1, 4, 4]
    parts = [assign_videos(6, r, 2, weights=w) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == list(range(6))
    assert abs(sum(w[i] for i in parts[0]) - sum(w[i] for i in parts[1])) <= 1
    with pytest.raises(ValueError):
        assign_videos(3, 0, 2, weights=[1])


def test_single_process_exchange_is_identity():
    ex = RangeExchange()
    assert (ex.rank, ex.size) == (0, 1) and ex.my_range(9) == (0, 9)
    pos = torch.tensor([1, 2, 90, 3], dtype=torch.int32)
    g = ex.finish_arrays(pos, torch.tensor([2], dtype=torch.int32), 4, _torch_merge)
    assert g.first_exit == 2 and g.pos.tolist() == [1, 2, FF_POS_DROPPED, FF_POS_DROPPED] and g.counts is None
    assert ex.transport == "gathered"
    with pytest.raises(ValueError):
        ex.finish(ex.begin(4, torch.device("cpu")))          # no engine and no merge callable
    with pytest.raises(ValueError):
        RangeExchange(transport="carrier-pigeon")


def test_steps_reuse_the_block_and_reset_the_header():
    ex = RangeExchange()
    for step, fe in enumerate((3, FF_NO_EXIT, 1)):
        blk = ex.begin(5, torch.device("cpu"))
        assert int(blk.first_exit[0]) == FF_NO_EXIT          # reset every step
        blk.pos[:5] = torch.arange(5, dtype=torch.int32) + 10 * step
        blk.counts[:5] = 7
        blk.first_exit[0] = fe
        g = ex.finish(blk, _torch_merge)
        want = [(10 * step + i) if i < fe else FF_POS_DROPPED for i in range(5)]
        assert g.pos.tolist() == want and g.counts.tolist() == [7] * 5 and g.first_exit == fe


def test_numa_binding_is_harmless_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


# ---- config 5: whole videos sharded across ranks -------------------------------------------
def _collection_worker(rank, size, port, vdir, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from high_speed_image_processing_b200.photron import open_collection
        from high_speed_image_processing_b200.process_videos import (FileCalibration, VideoSourceConfig,
                                                                     process_collection)
        coll = open_collection(vdir)
        cfgs = []
        for i in range(len(coll)):
            c = VideoSourceConfig(name=f"v{i}")
            c.detection_method = ("half_maximum", "threshold", "gradient")[i % 3]
            c.file_calibrations = [FileCalibration(calibration=0.5, position_offset=float(i), files=[f"run-{i}-"])]
            cfgs.append(c)
        seen = []

        def stub(video, cfg, cal, off):            # stands in for the GPU path: host logic only
            seen.append(video.filepath.name)
            return (video.filepath.name, cfg.detection_method, cal, off, len(video), rank)

        res = process_collection(coll, cfgs, exchange=RangeExchange(), per_video=stub)
        import json
        with open(os.path.join(out_dir, f"c{rank}.json"), "w") as fh:
            json.dump({"res": {str(k): list(v) for k, v in res.items()}, "seen": seen}, fh)
    finally:
        dist.destroy_process_group()


def test_collection_sharded_by_video(tmp_path):
    import json
    vdir = tmp_path / "videos"
    lengths = [4, 9, 3, 7, 5]
    for i, n in enumerate(lengths):
        syn.write_clip(vdir, f"run-{i}-", syn.SyntheticSpec(width=64, height=8, n_frames=n, bits=16, seed=i))
    out = tmp_path / "out"
    out.mkdir()
    mp.spawn(_collection_worker, args=(2, _free_port(), str(vdir), str(out)), nprocs=2, join=True)
    r0 = json.load(open(out / "c0.json"))
    r1 = json.load(open(out / "c1.json"))
    assert r0["res"] == r1["res"]                                   # every rank gets the full map
    assert sorted(r0["seen"] + r1["seen"]) == sorted(f"run-{i}-.cihx" for i in range(5))
    assert not set(r0["seen"]) & set(r1["seen"])                    # each video processed exactly once
    for i, n in enumerate(lengths):
        name, method, cal, off, length, _ = r0["res"][str(i)]
        assert (name, method, cal, off, length) == (f"run-{i}-.cihx", ("half_maximum", "threshold", "gradient")[i % 3],
                                                    0.5, float(i), n)
    load = [sum(lengths[int(k)] for k, v in r0["res"].items() if v[5] == r) for r in (0, 1)]
    assert abs(load[0] - load[1]) <= max(lengths)                   # size-balanced


# ---- the HEAD detector, range-sharded: image work per rank, tracker state handed rank to rank ----------
class _OracleHeadEngine:
    """CPU stand-in for the three engine calls of ``process_videos._process_video_head`` (the CUDA
    kernels need a GPU): the oracle's SciPy lines and its candidate selection, same tensor layouts."""
    device = torch.device("cpu")

    def upload(self, raw):
        return torch.from_numpy(np.ascontiguousarray(raw).reshape(-1).copy())

    def head_lines(self, frames, n, h, w, bits, hp, *, frame0=None, first_frame=0, halo=None, skip=None):
        from oracle import head_oracle as ho
        dec = fo.frames_from_bytes(frames.numpy(), n, h, w, bits)
        f0 = dec[0] if frame0 is None else fo.frames_from_bytes(frame0.numpy(), 1, h, w, bits)[0]
        bg = fo.background_scalar(f0)
        prior = None if halo is None else fo.subtract_scalar_background(fo.frames_from_bytes(halo.numpy(), 1, h, w, bits)[0], bg)
        lines = torch.zeros((n, 2, w), dtype=torch.float64)
        flags = torch.zeros(n, dtype=torch.uint8)
        for i in range(n):
            if skip is not None and int(skip[i]):      # :1443-1445: neither processed nor kept as the prior frame
                continue
            sub = fo.subtract_scalar_background(dec[i], bg)
            if not fo.is_empty_frame(sub, fo.empty_noise_threshold(bg), hp.min_signal_fraction):
                if prior is None:
                    flags[i] = 2
                else:
                    s, g = ho.detect_lines_scipy(fo.frame_difference(sub, prior, hp.frame_diff_threshold))
                    lines[i, 0], lines[i, 1], flags[i] = torch.from_numpy(s.copy()), torch.from_numpy(g.copy()), 1
            prior = sub
        return lines, flags, None

    @staticmethod
    def head_scalars(pending):
        return None

    def head_track_lines(self, lines, flags, first_frame, width, hp, max_displacement, tracker_state=(-1, -1)):
        from oracle import head_oracle as ho
        cfg = ho.HeadConfig()
        last_f, last_p = tracker_state
        track = torch.full((flags.numel(), 5), -1, dtype=torch.int32)
        stop = torch.tensor([FF_NO_EXIT, last_f, last_p], dtype=torch.int32)
        for i in range(flags.numel()):
            if flags[i] == 0:
                continue
            gf = first_frame + i
            if last_p < 0:
                s0, s1 = cfg.edge_margin_px, width - cfg.edge_margin_px
            else:
                s0, s1 = last_p, min(width - cfg.edge_margin_px, last_p + max_displacement * max(1, gf - last_f) + cfg.search_window_px)
            pa = pb = final = None
            if flags[i] == 1:
                pa, pb, final = ho.select_position(lines[i, 0].numpy(), lines[i, 1].numpy(), s0, s1, cfg)
            track[i] = torch.tensor([-1 if v is None else v for v in (final, pa, pb)] + [s0, s1], dtype=torch.int32)
            if final is not None:
                last_f, last_p = gf, final
                if final >= width - cfg.exit_margin_px:
                    stop[0] = gf
                    break
        stop[1], stop[2] = last_f, last_p
        return track, stop


def _head_worker(rank, size, port, n_frames, velocity, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from high_speed_image_processing_b200.photron import open_video
        from high_speed_image_processing_b200.process_videos import VideoSourceConfig, _process_video_head
        spec = syn.SyntheticSpec(width=160, height=12, n_frames=n_frames, style="nova", t_enter=5.0, velocity=velocity,
                                 tail_length=30.0, curvature_px=1.0, seed=13, bits=16)
        clip_dir = os.path.join(out_dir, f"clip{rank}")
        syn.write_clip(clip_dir, "run-1-", spec)
        cfg = VideoSourceConfig(name="t")
        cfg.detection_method = "head"
        with open_video(os.path.join(clip_dir, "run-1-.cihx")) as video:
            res = _process_video_head(video, cfg, 0.000833333, 1.347567, _OracleHeadEngine(), RangeExchange())
        import pickle
        with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
            pickle.dump({"rows": [list(r) for r in res.rows], "vel": res.velocity_history, "stop": res.stop,
                         "ddt": res.ddt_frame, "first_exit": res.first_exit}, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,n_frames,velocity", [(2, 61, 3.5), (3, 64, 3.5), (3, 50, 1.2)])
def test_head_detector_range_sharded_equals_oracle_loop(tmp_path, size, n_frames, velocity):
    """State hand-over between ranks (send/recv), an exit inside an early rank's range (later ranks must
    not track), no exit at all, and the all-gather / de-padding of the result rows."""
    import pickle
    from oracle import head_oracle as ho
    mp.spawn(_head_worker, args=(size, _free_port(), n_frames, velocity, str(tmp_path)), nprocs=size, join=True)
    spec = syn.SyntheticSpec(width=160, height=12, n_frames=n_frames, style="nova", t_enter=5.0, velocity=velocity,
                             tail_length=30.0, curvature_px=1.0, seed=13, bits=16)
    time_of = lambda i: fo.frame_time_absolute(i, spec.start_frame, spec.skip_frame, spec.record_rate)  # noqa: E731
    want = ho.run_head(syn.render_frames(spec), spec.record_rate, 0.000833333, 1.347567, time_of)
    assert len(want.rows) > 15 and (want.stop is not None) == (velocity > 2)
    for r in range(size):
        got = pickle.load(open(tmp_path / f"r{r}.pkl", "rb"))
        assert got["rows"] == want.rows and got["vel"] == want.velocity_history
        assert got["stop"] == want.stop and got["ddt"] == want.ddt_frame
        assert got["first_exit"] == (want.stop[1] if want.stop and want.stop[0] == "exit" else None)
