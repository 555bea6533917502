#!/usr/bin/env python
"""BASELINE config C5: a VideoCollection of 64 synthetic Nova+Mini recordings, mixed detection
methods and calibrations, sharded BY VIDEO across the GPUs (src/photron/parallel.py:173-208).

    python tools/bench_collection.py [--clips 64] [--frames 2000] [--dir /tmp/ff_c5]
    torchrun --nproc-per-node N tools/bench_collection.py ...

Clips alternate C2 (1024x128, Nova-style) and C3 (1024x256, Mini-style) shapes, methods cycle
half_maximum / threshold / gradient, every clip has its own FileCalibration rule.  The files are
real .cihx/.mraw pairs on disk opened through ``open_collection``; each rank stages ITS videos in
pinned host memory (``video.pin_memory()``, untimed I/O), then the timed region runs
``process_collection``: per video frame 0 -> scalars, chunked H2D streaming + kernels, result
rows on the host; one all_gather_object of the per-video results at the end.  Prints one JSON line
on rank 0.  (Parity of this path against the oracle: tests/test_gpu_pipeline.py.)
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200.engine import FlameFrontEngine  # noqa: E402
from high_speed_image_processing_b200.photron import open_collection  # noqa: E402
from high_speed_image_processing_b200.process_videos import (FileCalibration, VideoSourceConfig,  # noqa: E402
                                                             process_collection)
from high_speed_image_processing_b200.sharding import RangeExchange, assign_videos, bind_to_gpu_numa_node  # noqa: E402

METHODS = ("half_maximum", "threshold", "gradient")


def clip_spec(i: int, frames: int) -> syn.SyntheticSpec:
    base = syn.config_spec("C2" if i % 2 == 0 else "C3", n_frames=frames, seed=5000 + i)
    # the front enters in the last ~55 % of the clip and leaves before the end on most clips
    return syn.SyntheticSpec(**{**base.__dict__, "t_enter": float(max(2, frames - 1100 - 37 * (i % 5)))})


def clip_config(i: int) -> VideoSourceConfig:
    cfg = VideoSourceConfig(name=f"clip{i:02d}")
    cfg.detection_method = METHODS[i % 3]
    cfg.file_calibrations = [FileCalibration(calibration=0.0008 + 1e-6 * i, position_offset=0.05 * i,
                                             files=[f"run-{i:02d}-"])]
    return cfg


def run_collection(eng, exchange, device, rank: int, world: int, clips: int, frames: int, vdir, reps: int = 3,
                   residency: str = "auto", pin: bool = True) -> dict:
    """Write `clips` recordings under `vdir` (each rank writes the ones it will process), open them as a
    VideoCollection, stage this rank's videos in pinned memory and time ``process_collection``.
    Returns {"summary": JSON-able figures, "results", "specs", "configs", "mine"}."""
    vdir = Path(vdir)
    specs = [clip_spec(i, frames) for i in range(clips)]
    weights = [s.n_frames * s.height * s.width for s in specs]
    mine = assign_videos(clips, rank, world, weights)
    t0 = time.perf_counter()
    vdir.mkdir(parents=True, exist_ok=True)
    for i in mine:
        spec = specs[i]
        packed = syn.render_packed_torch(spec, device).cpu().numpy()
        (vdir / f"run-{i:02d}-.mraw").write_bytes(packed.tobytes())
        (vdir / f"run-{i:02d}-.cihx").write_bytes(syn.cihx_bytes(spec, spec.n_frames))
        del packed
    write_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()

    coll = open_collection(str(vdir))
    assert len(coll) == clips, f"expected {clips} recordings, found {len(coll)}"
    cfgs = [clip_config(i) for i in range(clips)]
    t0 = time.perf_counter()
    if pin:
        for i in mine:
            coll[i].pin_memory()
    stage_s = time.perf_counter() - t0
    my_bytes = sum(coll[i].frame_store.nbytes_raw for i in mine)
    total_frames = sum(len(v) for v in coll)
    total_bytes = sum(v.frame_store.nbytes_raw for v in coll)

    from high_speed_image_processing_b200.process_videos import process_video

    def per_video(video, cfg, cal, off):
        return process_video(video, cfg, cal, off, engine=eng, exchange=None, residency=residency)

    def run():
        return process_collection(coll, cfgs, engine=eng, exchange=exchange if world > 1 else None,
                                  per_video=per_video)

    results = run()                                   # warm-up (allocations, contexts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        results = run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt)
    sec = min(times)
    n_rows = sum(len(r.rows) for r in results.values())
    exits = sum(1 for r in results.values() if r.first_exit is not None)
    summary = {
        "config": f"BASELINE config 5: VideoCollection of {clips} synthetic Nova+Mini recordings x {frames} frames, mixed "
                  f"methods / calibrations, sharded by whole videos over {world} GPU(s)",
        "scaling": "weak", "pinned": pin, "residency": residency, "clips": clips, "frames_per_clip": frames,
        "n_gpus": world, "total_frames": total_frames, "total_gb": total_bytes / 1e9,
        "sharding": "whole videos, size-balanced (assign_videos)", "seconds": sec,
        "value": total_frames / sec, "unit": "frames/s", "gbs_aggregate": total_bytes / sec / 1e9,
        "gbs_per_gpu": total_bytes / sec / 1e9 / world, "ms_per_video": sec / (clips / world) * 1e3,
        "all_times_s": times, "detections": n_rows, "clips_with_exit": exits,
        "untimed": {"write_files_s": write_s, "pin_stage_s": stage_s, "rank0_bytes": my_bytes}}
    coll.close_all()
    return {"summary": summary, "results": results, "specs": specs, "configs": cfgs, "mine": mine}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--dir", default="/tmp/ff_c5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--no-pin", action="store_true", help="leave the recordings memory-mapped (pageable) instead of "
                    "staging them in pinned memory: the page-cache -> GPU path a plain script would take")
    ap.add_argument("--residency", default="auto", choices=["auto", "device", "host"])
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    bind_to_gpu_numa_node(local)
    eng = FlameFrontEngine(local)
    exchange = RangeExchange(engine=eng)
    rep = run_collection(eng, exchange, device, rank, world, args.clips, args.frames, args.dir, reps=args.reps,
                         residency=args.residency, pin=not args.no_pin)
    # (parity of the collection path against the oracle: tests/test_gpu_pipeline.py and bench.py's c5 leg;
    # here only the calibration arithmetic of the rows is re-checked, which needs no oracle)
    for i in rep["mine"][:3]:
        cal, off = rep["configs"][i].get_calibration_for_file(f"run-{i:02d}-.cihx")
        for frame_idx, t_s, px, p_m, _ in rep["results"][i].rows[:50]:
            assert p_m == px * cal + off
    if rank == 0:
        print(json.dumps(rep["summary"]), flush=True)
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.keep:
        shutil.rmtree(args.dir, ignore_errors=True)
    exchange.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
