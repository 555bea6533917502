#!/usr/bin/env python
"""TEST INFRASTRUCTURE (it checks against the oracle; run by tests/test_gpu_exchange.py or by hand).
Range-sharded run under torchrun: both block transports of the exchange step against the
oracle's serial answer, plus the latency of the step itself.

    torchrun --nproc-per-node N tests/check_exchange.py [--out report.json] [--steps 20]

Rank 0 writes a JSON report; every rank exits non-zero on a mismatch.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from high_speed_image_processing_b200 import synthetic as syn  # noqa: E402
from high_speed_image_processing_b200._cabi import FF_NO_EXIT, FF_POS_DROPPED  # noqa: E402
from high_speed_image_processing_b200.engine import DetectionParams, FlameFrontEngine  # noqa: E402
from high_speed_image_processing_b200.sharding import RangeExchange  # noqa: E402
from oracle import flame_oracle as fo  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    eng = FlameFrontEngine(local)
    report = {"world": world, "transports": {}, "ok": True}

    cases = [("threshold", 403, 9.0), ("half_maximum", 400, 12.0), ("gradient", 397, 300.0)]   # last: no exit
    for transport in ("peer", "gathered"):
        ex = RangeExchange(engine=eng, transport=transport)
        lat = []
        for method, n_frames, t_enter in cases:
            spec = syn.SyntheticSpec(width=256, height=32, n_frames=n_frames, style="mini" if method != "half_maximum"
                                     else "nova", t_enter=t_enter, velocity=1.5, seed=77)
            frames = syn.render_frames(spec)
            want = fo.process_clip(frames, fo.ClipParams(method=method))
            exp = want.pos_px.copy()
            exp[want.first_exit:] = FF_POS_DROPPED
            exp_fe = want.first_exit if want.first_exit < n_frames else FF_NO_EXIT
            packed = torch.from_numpy(syn.pack_frames(frames, 12)).to(device)
            fb = spec.frame_bytes
            a, b = ex.my_range(n_frames)
            mine = packed[a * fb:b * fb].clone()
            halo = packed[(a - 1) * fb:a * fb].clone() if a > 0 else None
            frame0 = packed[:fb].clone()
            params = DetectionParams(method=method)
            for step in range(args.steps):
                blk = ex.begin(n_frames)
                eng.process_range(mine, b - a, spec.height, spec.width, 12, params, frame0=frame0, first_frame=a,
                                  halo=halo, **ex.range_kwargs(blk))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g = ex.finish(blk).wait()        # (the merge runs on a side stream; wait() joins it)
                e1.record()
                torch.cuda.synchronize()
                if step >= 3:
                    lat.append(e0.elapsed_time(e1) * 1e3)
                ok = (np.array_equal(g.pos.cpu().numpy(), exp) and g.first_exit == exp_fe and
                      np.array_equal(g.counts.cpu().numpy(), want.nonempty.astype(np.int32)))
                if not ok:
                    report["ok"] = False
                    print(f"[rank {rank}] MISMATCH transport={transport} method={method} step={step}", flush=True)
            ex.check()
        assert ex.transport == transport, f"asked for {transport}, got {ex.transport}"
        lat.sort()
        report["transports"][transport] = {"finish_us_median": lat[len(lat) // 2] if lat else None,
                                           "finish_us_min": lat[0] if lat else None, "samples": len(lat)}
        ex.close()

    # ---- the drop-in driver on a file, range-sharded, both residencies ------------------------------
    import tempfile
    from high_speed_image_processing_b200.photron import open_video
    from high_speed_image_processing_b200.process_videos import VideoSourceConfig, process_video
    spec = syn.SyntheticSpec(width=512, height=64, n_frames=301, style="nova", t_enter=15.0, velocity=2.5, seed=5)
    frames = syn.render_frames(spec)
    want = fo.process_clip(frames, fo.ClipParams(method="half_maximum"))
    root = [tempfile.mkdtemp(prefix="ff_xchg_") if rank == 0 else None]
    dist.broadcast_object_list(root, src=0)
    if rank == 0:
        syn.write_clip(root[0], "run-1-", spec, frames=frames)
    dist.barrier()
    cfg = VideoSourceConfig(name="t")
    cfg.detection_method = "half_maximum"
    cfg.skip_frames = [40, 41]
    want_skip = fo.process_clip(frames, fo.ClipParams(method="half_maximum", skip_frames=[40, 41]))
    ex = RangeExchange(engine=eng)
    for residency in ("host", "device"):
        with open_video(f"{root[0]}/run-1-.cihx") as video:
            res = process_video(video, cfg, 0.001, 0.5, engine=eng, exchange=ex, residency=residency)
        same = ([(r[0], r[2]) for r in res.rows] == want_skip.records and
                res.first_exit == (want_skip.first_exit if want_skip.first_exit < spec.n_frames else None))
        if not same:
            report["ok"] = False
            print(f"[rank {rank}] MISMATCH process_video residency={residency}", flush=True)
    report["process_video_sharded"] = "host+device residency vs oracle rows"
    # ---- the HEAD detector, range-sharded: image pipeline per rank, tracker state handed rank to rank ----
    from oracle import head_oracle as ho
    hcfg = VideoSourceConfig(name="t")
    hcfg.detection_method = "head"
    for label, hspec in (("exit", syn.SyntheticSpec(width=512, height=64, n_frames=301, style="nova", t_enter=15.0,
                                                    velocity=2.5, seed=5)),
                         ("no_exit", syn.SyntheticSpec(width=512, height=32, n_frames=157, style="mini", t_enter=90.0,
                                                       velocity=1.5, seed=6))):
        hframes = syn.render_frames(hspec)
        if rank == 0:
            syn.write_clip(root[0], f"head-{label}", hspec, frames=hframes)
        dist.barrier()
        with open_video(f"{root[0]}/head-{label}.cihx") as video:
            res = process_video(video, hcfg, 0.000833333, 1.347567, engine=eng, exchange=ex)
            want_h = ho.run_head(hframes, video.frame_rate, 0.000833333, 1.347567, video.get_absolute_time)
        same = ([list(r) for r in res.rows] == want_h.rows and res.velocity_history == want_h.velocity_history
                and res.ddt_frame == want_h.ddt_frame and res.stop == want_h.stop and len(want_h.rows) > 20)
        if not same:
            report["ok"] = False
            print(f"[rank {rank}] MISMATCH head sharded {label}: {res.stop} vs {want_h.stop}, "
                  f"{len(res.rows)} vs {len(want_h.rows)} rows", flush=True)
    report["head_sharded"] = "rows, velocities, DDT frame and stop vs the oracle loop (exit / no exit)"
    ex.close()
    dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(root[0], ignore_errors=True)

    flag = torch.tensor([1 if report["ok"] else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["ok"] = bool(flag.item())
    if rank == 0:
        text = json.dumps(report)
        print(text, flush=True)
        if args.out:
            Path(args.out).write_text(text)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()
