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
